"""Matrix Market coordinate I/O for the arrays `CsrMatrix::from_csr` takes.

The reference has no on-disk format at all (its examples build matrices in code); this is the host-side step
before the path so that user matrices beyond the synthetic stencils can reach `DeviceCsr.from_csr`.
Supported: `matrix coordinate {real|integer|pattern} {general|symmetric|skew-symmetric}`; entries are sorted by
(row, column); repeated entries are summed in file order (the usual assembly convention); explicit zeros are kept
(`from_csr` keeps them too, sparse.rs:26-47).  Pure numpy host code: nothing here touches the GPU.
"""
import numpy as np


class MatrixMarketError(ValueError):
    pass


def read_matrix_market(path):
    """-> (nrows, ncols, row_ptr[uint64], col_idx[uint64], values[float64]) with strictly ascending columns per row."""
    with open(path, "r") as f:
        header = f.readline().split()
        if len(header) < 5 or header[0] != "%%MatrixMarket" or header[1].lower() != "matrix":
            raise MatrixMarketError("not a Matrix Market matrix file")
        fmt, field, sym = header[2].lower(), header[3].lower(), header[4].lower()
        if fmt != "coordinate":
            raise MatrixMarketError("only the coordinate format is supported, got %r" % fmt)
        if field not in ("real", "integer", "pattern"):
            raise MatrixMarketError("unsupported field %r" % field)
        if sym not in ("general", "symmetric", "skew-symmetric"):
            raise MatrixMarketError("unsupported symmetry %r" % sym)
        line = f.readline()
        while line and (line.startswith("%") or not line.strip()):
            line = f.readline()
        try:
            nrows, ncols, nent = (int(t) for t in line.split())
        except Exception:
            raise MatrixMarketError("bad size line %r" % line)
        body = np.loadtxt(f, ndmin=2, dtype=np.float64) if nent else np.zeros((0, 3))
    want = 2 if field == "pattern" else 3
    if body.shape[0] != nent or (nent and body.shape[1] < want):
        raise MatrixMarketError("expected %d entries with %d columns, got array of shape %r" % (nent, want, body.shape))
    i = body[:, 0].astype(np.int64) - 1
    j = body[:, 1].astype(np.int64) - 1
    v = np.ones(nent) if field == "pattern" else body[:, 2].astype(np.float64)
    if nent and (i.min() < 0 or j.min() < 0 or i.max() >= nrows or j.max() >= ncols):
        raise MatrixMarketError("entry index out of range")
    if sym != "general":
        if nrows != ncols:
            raise MatrixMarketError("symmetric storage needs a square matrix")
        off = i != j
        if sym == "skew-symmetric" and np.any(~off):
            raise MatrixMarketError("skew-symmetric storage cannot hold diagonal entries")
        sign = -1.0 if sym == "skew-symmetric" else 1.0
        i, j, v = np.concatenate([i, j[off]]), np.concatenate([j, i[off]]), np.concatenate([v, sign * v[off]])
    return (nrows, ncols) + coo_to_csr(nrows, ncols, i, j, v)


def coo_to_csr(nrows, ncols, i, j, v):
    """Sort by (row, column) (stable), sum repeats in input order -> (row_ptr, col_idx, values)."""
    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    v = np.asarray(v, dtype=np.float64)
    order = np.lexsort((j, i))            # stable: ties keep input order
    i, j, v = i[order], j[order], v[order]
    if i.size:
        first = np.ones(i.size, dtype=bool)
        first[1:] = (i[1:] != i[:-1]) | (j[1:] != j[:-1])
        starts = np.flatnonzero(first)
        if starts.size != i.size:
            # sequential left-to-right sums per group, in file order (np.add.reduceat adds left to right)
            v = np.add.reduceat(v, starts)
            i, j = i[starts], j[starts]
    row_ptr = np.zeros(nrows + 1, dtype=np.uint64)
    np.add.at(row_ptr, i + 1, 1)
    row_ptr = np.cumsum(row_ptr, dtype=np.uint64)
    return row_ptr, j.astype(np.uint64), v


def write_matrix_market(path, nrows, ncols, row_ptr, col_idx, values, comment=None):
    """Write `matrix coordinate real general` with 17 significant digits (round-trips f64 exactly)."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    col_idx = np.asarray(col_idx, dtype=np.int64)
    values = np.asarray(values, dtype=np.float64)
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        if comment:
            for ln in str(comment).splitlines():
                f.write("% " + ln + "\n")
        f.write("%d %d %d\n" % (nrows, ncols, col_idx.size))
        rows = np.repeat(np.arange(nrows, dtype=np.int64), np.diff(row_ptr))
        for r, c, x in zip(rows + 1, col_idx + 1, values):
            f.write("%d %d %s\n" % (r, c, repr(float(x))))
