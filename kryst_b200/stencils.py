"""Synthetic stencil operators of BASELINE.json's configs, as host CSR arrays.

Deterministic integer / IEEE-exact arithmetic (no RNG library): ordering row = i + N j (+ N^2 k),
Dirichlet truncation, ascending columns, usize (uint64) indices like CsrMatrix::from_csr.
Every generator can emit a row block [lo, hi) (global column indices) so each rank of a
row-partitioned run builds only its own shard.  Vectorised numpy; an independent C++ twin lives in
the oracle and tests check the two bit-for-bit.
"""
import numpy as np

KINDS = ("poisson2d", "convdiff2d", "varcoef27", "poisson3d", "convdiff3d")
_MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def dim(kind, N):
    return N * N if kind in ("poisson2d", "convdiff2d") else N * N * N


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _MASK64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK64
    return z ^ (z >> np.uint64(31))


def kappa(r):
    """kappa(r) = 0.1 + 1.9 * ((splitmix64(0x5EEDB200 + r) >> 11) * 2^-53)  in [0.1, 2.0)."""
    with np.errstate(over="ignore"):
        z = _splitmix64(np.asarray(r, dtype=np.uint64) + np.uint64(0x5EEDB200))
    return 0.1 + 1.9 * ((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0))


def _compress(rows_n, cols, vals, mask):
    counts = mask.sum(axis=1, dtype=np.int64)
    row_ptr = np.zeros(rows_n + 1, dtype=np.uint64)
    np.cumsum(counts, out=row_ptr[1:].view(np.int64))
    return row_ptr, cols[mask].astype(np.uint64), vals[mask]


def stencil(kind, N, lo=None, hi=None, pe=(0.4, 0.2, 0.1)):
    """-> (n_global, row_ptr[hi-lo+1], col_idx, vals) for rows [lo, hi)."""
    n = dim(kind, N)
    lo = 0 if lo is None else int(lo)
    hi = n if hi is None else int(hi)
    row = np.arange(lo, hi, dtype=np.int64)
    m = row.size
    i = row % N
    j = (row // N) % N
    k = row // (N * N)
    px, py, pz = pe
    if kind in ("poisson2d", "convdiff2d"):
        offs = np.array([-N, -1, 0, 1, N], dtype=np.int64)
        if kind == "poisson2d":
            v = np.array([-1.0, -1.0, 4.0, -1.0, -1.0])
        else:
            v = np.array([-(1.0 + py), -(1.0 + px), (4.0 + px) + py, -1.0, -1.0])
        mask = np.stack([j > 0, i > 0, np.ones(m, bool), i < N - 1, j < N - 1], axis=1)
    elif kind in ("poisson3d", "convdiff3d"):
        offs = np.array([-N * N, -N, -1, 0, 1, N, N * N], dtype=np.int64)
        if kind == "poisson3d":
            v = np.array([-1.0, -1.0, -1.0, 6.0, -1.0, -1.0, -1.0])
        else:
            v = np.array([-(1.0 + pz), -(1.0 + py), -(1.0 + px), ((6.0 + px) + py) + pz, -1.0, -1.0, -1.0])
        mask = np.stack([k > 0, j > 0, i > 0, np.ones(m, bool), i < N - 1, j < N - 1, k < N - 1], axis=1)
    elif kind == "varcoef27":
        if N < 3:
            raise ValueError("varcoef27 needs N >= 3")
        kr = kappa(row)
        cols = np.empty((m, 27), dtype=np.int64)
        vals = np.zeros((m, 27))
        mask = np.zeros((m, 27), dtype=bool)
        diag = np.zeros(m)
        t = 0
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    off = dx + N * dy + N * N * dz
                    inside = ((i + dx >= 0) & (i + dx < N) & (j + dy >= 0) & (j + dy < N) & (k + dz >= 0) & (k + dz < N))
                    cols[:, t] = row + off
                    mask[:, t] = inside
                    if not (dx == 0 and dy == 0 and dz == 0):
                        kc = kappa(np.where(inside, row + off, 0))
                        w = 0.5 * (kr + kc)
                        diag = diag + np.where(inside, w, kr)
                        vals[:, t] = -w
                    t += 1
        vals[:, 13] = diag
        rp, ci, vv = _compress(m, cols, vals, mask)
        return n, rp, ci, vv
    else:
        raise ValueError("unknown stencil kind %r" % (kind,))
    cols = row[:, None] + offs[None, :]
    vals = np.broadcast_to(v[None, :], cols.shape)
    rp, ci, vv = _compress(m, cols, vals, mask)
    return n, rp, ci, vv


# BASELINE.json configs
CONFIGS = {
    "C1": dict(kind="poisson2d", N=512, solver="pcg", pc="jacobi"),
    "C2": dict(kind="convdiff2d", N=1024, solver="gmres", restart=30, pc="ilu0"),
    "C3": dict(kind="varcoef27", N=128, solver="bicgstab", pc="jacobi"),
    "C4": dict(kind="poisson3d", N=256, solver="pcg", pc="jacobi"),
    "C4g": dict(kind="poisson3d", N=256, solver="gmres", restart=30, pc="ilu0"),
    "C5": dict(kind="convdiff3d", N=384, solver="gmres", restart=50, pc="ilu0"),
}


def spmv_bytes(n, nnz):
    """Algorithmic bytes of one SpMV: 12 nnz + 4 (n+1) + 16 n."""
    return 12 * nnz + 4 * (n + 1) + 16 * n
