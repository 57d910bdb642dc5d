"""CPU model of the x-staging tables of kb_spmv_xtile.cuh (chunking, column intervals, 16-bit local ids, tail slot).

It restates, step for step, what kb_xt_chunk_build / kb_xt_build compute on the device and what the producer warp
copies, and checks the invariants the kernel relies on: every copy is 16-byte aligned and sized, windows stay inside
the padded allocations, a stage never overflows, and xs[lcol[q]] is x[col[q]] for every stored entry.  The GPU parity
of the kernel itself is tests/test_gpu_xtile.py."""
import numpy as np
import pytest

from kryst_b200 import stencils

KB_TILE, KMAX, GAP = 512, 16, 4
CFGS = {0: dict(cap=3072, xcap=2048, maxrows=512), 1: dict(cap=3584, xcap=1280, maxrows=256), 2: dict(cap=1792, xcap=1296, maxrows=256)}


def chunk_build(rp, n, cap, maxrows):
    tile_chunk, chunk_row, chunk_nz = [0], [], []
    for r0 in range(0, n, KB_TILE):
        r1 = min(n, r0 + KB_TILE)
        tnz = int(rp[r1] - rp[r0])
        k = max(1, (tnz + cap - 1) // cap)
        rt = max(1, min(maxrows, (r1 - r0 + k - 1) // k))
        start, base = r0, int(rp[r0])
        chunk_row.append(r0); chunk_nz.append(base)
        for r in range(r0, r1):
            e = int(rp[r + 1])
            if r > start and (e - base > cap or r - start >= rt):
                start, base = r, int(rp[r])
                chunk_row.append(r); chunk_nz.append(base)
        tile_chunk.append(len(chunk_row))
    chunk_row.append(n); chunk_nz.append(int(rp[n]))
    return tile_chunk, chunk_row, chunk_nz


def xt_build(col, chunk_nz, c, ncols, xcap):
    nz0, nz1 = chunk_nz[c], chunk_nz[c + 1]
    key = np.sort(col[nz0:nz1])
    nv = key.size
    starts = [q for q in range(nv) if q == 0 or (key[q] >> 1) - (key[q - 1] >> 1) > GAP]
    if len(starts) > KMAX:
        return None
    ne = ncols & ~1
    lo, ln, off_k, off, tl = [], [], [], 0, -1
    for k, q0 in enumerate(starts):
        q1 = starts[k + 1] if k + 1 < len(starts) else nv
        l, h = (int(key[q0]) >> 1) << 1, ((int(key[q1 - 1]) >> 1) + 1) << 1
        if h > ne:
            h, tl = ne, 0
        l = min(l, h)
        lo.append(l); ln.append(h - l); off_k.append(off)
        off += h - l
    if tl == 0:
        tl = off
        off += 1
    if off > xcap:
        return None
    lcol = np.zeros(nz1 - nz0, dtype=np.uint16)
    for i, cc in enumerate(col[nz0:nz1]):
        if tl >= 0 and cc == ncols - 1:
            lcol[i] = tl
            continue
        for k in range(len(lo)):
            if lo[k] <= cc < lo[k] + ln[k]:
                lcol[i] = off_k[k] + cc - lo[k]
                break
        else:
            raise AssertionError("column %d not covered" % cc)
    return lo, ln, tl, lcol


def model(n, ncols, rp, ci, cfg):
    g = CFGS[cfg]
    rp = np.asarray(rp, dtype=np.int64); ci = np.asarray(ci, dtype=np.int64)
    nnz = int(rp[n])
    if n == 0 or nnz == 0 or np.diff(rp).max() > g["cap"]:
        return None
    tile_chunk, chunk_row, chunk_nz = chunk_build(rp, n, g["cap"], g["maxrows"])
    assert tile_chunk[-1] == len(chunk_row) - 1
    x = np.random.default_rng(0).standard_normal(ncols)
    for c in range(len(chunk_row) - 1):
        ra, rb, nz0, nz1 = chunk_row[c], chunk_row[c + 1], chunk_nz[c], chunk_nz[c + 1]
        assert ra // KB_TILE == (rb - 1) // KB_TILE and 0 < rb - ra <= g["maxrows"] and nz1 - nz0 <= g["cap"]
        assert nz0 == rp[ra] and nz1 == rp[rb]
        t = xt_build(ci, chunk_nz, c, ncols, g["xcap"])
        if t is None:
            return None
        lo, ln, tl, lcol = t
        # what the producer copies
        b0, b1 = nz0 & ~7, (nz1 + 7) & ~7
        assert b1 - b0 <= g["cap"] + 16 and b1 <= nnz + 8            # vals / col allocations carry 8 spare entries, lcol 16
        r_al = ra & ~3
        nrp = ((rb + 1 - r_al) + 3) & ~3
        assert nrp <= g["maxrows"] + 8 and r_al + nrp <= n + 1 + 8   # row_ptr allocation carries 8 spare entries
        xs = np.full(g["xcap"] + 8, np.nan)
        off = 0
        for l, m in zip(lo, ln):
            assert l % 2 == 0 and m % 2 == 0 and off % 2 == 0 and 0 <= l and l + m <= ncols
            xs[off:off + m] = x[l:l + m]
            off += m
        if tl >= 0:
            assert tl == off and ncols % 2 == 1
            xs[tl] = x[ncols - 1]
        assert np.array_equal(xs[lcol], x[ci[nz0:nz1]])
    return len(chunk_row) - 1


@pytest.mark.parametrize("cfg", [0, 1, 2])
@pytest.mark.parametrize("kind,N", [("varcoef27", 9), ("varcoef27", 24), ("poisson3d", 17), ("convdiff3d", 12), ("poisson2d", 33), ("convdiff2d", 70)])
def test_stencils_fit_and_cover(cfg, kind, N):
    n, rp, ci, v = stencils.stencil(kind, N)
    assert model(n, n, rp, ci, cfg) is not None


@pytest.mark.parametrize("cfg", [0, 1])
def test_27pt_128_chunk_geometry(cfg):
    """The C3 operator: interior chunks have 9 intervals (3 planes x 3 lines) and fit a stage with room to spare."""
    N = 128
    g = CFGS[cfg]
    # one interior tile's worth of rows of the 27-point pattern, built directly
    r = np.arange(40 * N * N + 7 * N, 40 * N * N + 7 * N + KB_TILE)
    offs = np.array([dz * N * N + dy * N + dx for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)])
    cols = (r[:, None] + offs[None, :])
    rp = np.arange(0, 27 * KB_TILE + 1, 27)
    tile_chunk, chunk_row, chunk_nz = chunk_build(rp, KB_TILE, g["cap"], g["maxrows"])
    assert tile_chunk == [0, 5 if cfg == 0 else 4]
    t = xt_build(cols.reshape(-1), chunk_nz, 1, N ** 3, g["xcap"])
    assert t is not None
    lo, ln, tl, lcol = t
    # geometry 0: 103-row chunks, 9 separate lines; geometry 1: 128-row chunks = whole grid lines, the three lines of a plane merge
    assert len(lo) == (9 if cfg == 0 else 3) and tl == -1 and sum(ln) <= g["xcap"]


@pytest.mark.parametrize("cfg", [0, 1, 2])
@pytest.mark.parametrize("n,m", [(1, 1), (5, 5), (513, 513), (700, 707), (1500, 1501)])
def test_ragged_and_odd(cfg, n, m):
    rng = np.random.default_rng(n)
    rp, ci = [0], []
    for i in range(n):
        k = int(rng.integers(0, 12)) if i % 5 else 0
        centre = min(m - 1, i)
        cand = np.unique(np.clip(centre + rng.integers(-6, 7, size=k), 0, m - 1))
        if i == n - 1:
            cand = np.unique(np.append(cand, m - 1))          # reference the last column (tail slot when m is odd)
        ci.extend(cand.tolist()); rp.append(len(ci))
    if rp[-1] == 0:
        pytest.skip("empty")
    assert model(n, m, rp, ci, cfg) is not None


def test_scattered_columns_do_not_fit():
    rng = np.random.default_rng(3)
    n = 600
    rp, ci = [0], []
    for i in range(n):
        ci.extend(np.sort(rng.choice(200000, size=9, replace=False)).tolist()); rp.append(len(ci))
    assert model(n, 200000, rp, ci, 0) is None
