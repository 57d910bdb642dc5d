"""ctypes binding to the CPU oracle (oracle/libkryst_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (kryst_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(_ORACLE_DIR, "libkryst_oracle.so")

u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class KoStats(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("final_residual", C.c_double),
                ("converged", C.c_int32), ("breakdown", C.c_int32)]


class KoCsr(C.Structure):
    _fields_ = [("n", C.c_uint64), ("ncols", C.c_uint64), ("row_ptr", u64p), ("col_idx", u64p), ("vals", f64p)]


def build_oracle(force=False):
    src = os.path.join(_ORACLE_DIR, "kryst_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_SO)
        L.ko_dot.restype = C.c_double
        L.ko_dot.argtypes = [C.c_uint64, f64p, f64p]
        L.ko_dot_sharded.restype = C.c_double
        L.ko_dot_sharded.argtypes = [C.c_uint64, f64p, f64p, C.c_uint64]
        L.ko_norm.restype = C.c_double
        L.ko_norm.argtypes = [C.c_uint64, f64p]
        L.ko_sum.restype = C.c_double
        L.ko_sum.argtypes = [C.c_uint64, f64p]
        L.ko_spmv.argtypes = [C.POINTER(KoCsr), f64p, f64p]
        L.ko_csr_validate.restype = C.c_uint64
        L.ko_csr_validate.argtypes = [C.POINTER(KoCsr)]
        L.ko_jacobi_setup.argtypes = [C.POINTER(KoCsr), f64p]
        L.ko_jacobi_apply.argtypes = [C.c_uint64, f64p, f64p, f64p]
        L.ko_ilu0_factor.restype = C.c_int
        L.ko_ilu0_factor.argtypes = [C.POINTER(KoCsr), f64p, u64p, f64p, u64p]
        L.ko_ilu0_apply.argtypes = [C.POINTER(KoCsr), f64p, u64p, f64p, f64p, f64p]
        L.ko_levels.restype = C.c_uint64
        L.ko_levels.argtypes = [C.POINTER(KoCsr), C.c_int, u64p, u64p, u64p]
        L.ko_ilu_literal_setup.argtypes = [C.c_uint64, f64p, f64p, f64p]
        L.ko_ilu_literal_apply.argtypes = [C.c_uint64, f64p, f64p, f64p, f64p]
        for f in ("ko_pc_create_jacobi",):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.POINTER(KoCsr)]
        L.ko_pc_create_ilu0.restype = C.c_void_p
        L.ko_pc_create_ilu0.argtypes = [C.POINTER(KoCsr), C.c_uint64]
        L.ko_pc_create_ilu_literal.restype = C.c_void_p
        L.ko_pc_create_ilu_literal.argtypes = [C.c_uint64, f64p]
        L.ko_pc_create_asm.restype = C.c_void_p
        L.ko_pc_create_asm.argtypes = [C.POINTER(KoCsr), C.c_uint64, C.c_uint64, u64p, u64p, C.c_int]
        L.ko_pc_asm_block_size.restype = C.c_uint64
        L.ko_pc_asm_block_size.argtypes = [C.c_void_p, C.c_uint64]
        L.ko_pc_asm_block_indices.argtypes = [C.c_void_p, C.c_uint64, u64p]
        L.ko_pc_status.restype = C.c_int
        L.ko_pc_status.argtypes = [C.c_void_p, u64p]
        L.ko_pc_apply.argtypes = [C.c_void_p, f64p, f64p]
        L.ko_pc_destroy.argtypes = [C.c_void_p]
        L.ko_pc_block_nnz.restype = C.c_uint64
        L.ko_pc_block_nnz.argtypes = [C.c_void_p, C.c_uint64]
        L.ko_pc_block_get.argtypes = [C.c_void_p, C.c_uint64, f64p, u64p, f64p, u64p, u64p]
        L.ko_pcg.restype = C.c_int
        L.ko_pcg.argtypes = [C.POINTER(KoCsr), C.c_void_p, f64p, f64p, C.c_double, C.c_uint64, C.c_int,
                             C.c_uint64, f64p, C.c_uint64, u64p, C.POINTER(KoStats)]
        L.ko_pcg_pipe.restype = C.c_int
        L.ko_pcg_pipe.argtypes = [C.POINTER(KoCsr), C.c_void_p, f64p, f64p, C.c_double, C.c_uint64, C.c_int,
                                  C.c_uint64, f64p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(KoStats)]
        L.ko_pcg_sr.restype = C.c_int
        L.ko_pcg_sr.argtypes = [C.POINTER(KoCsr), C.c_void_p, f64p, f64p, C.c_double, C.c_uint64, C.c_int,
                             C.c_uint64, f64p, C.c_uint64, u64p, C.POINTER(KoStats)]
        L.ko_gmres.restype = C.c_int
        L.ko_gmres.argtypes = [C.POINTER(KoCsr), C.c_void_p, f64p, f64p, C.c_uint64, C.c_double, C.c_uint64,
                               C.c_int, C.c_int, C.c_uint64, C.POINTER(KoStats)]
        L.ko_fgmres.restype = C.c_int
        L.ko_fgmres.argtypes = [C.POINTER(KoCsr), C.c_void_p, f64p, f64p, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64,
                                C.POINTER(KoStats)]
        L.ko_bicgstab.restype = C.c_int
        L.ko_bicgstab.argtypes = [C.POINTER(KoCsr), C.c_void_p, f64p, f64p, C.c_double, C.c_uint64,
                                  C.c_int, C.c_uint64, C.POINTER(KoStats)]
        L.ko_submatrix.restype = C.c_uint64
        L.ko_submatrix.argtypes = [C.POINTER(KoCsr), u64p, C.c_uint64, u64p, u64p, f64p]
        L.ko_partition_range.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p]
        L.ko_ghost_list.restype = C.c_uint64
        L.ko_ghost_list.argtypes = [C.POINTER(KoCsr), C.c_uint64, C.c_uint64, u64p]
        L.ko_stencil_dim.restype = C.c_uint64
        L.ko_stencil_dim.argtypes = [C.c_int, C.c_uint64]
        L.ko_stencil.restype = C.c_uint64
        L.ko_stencil.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double,
                                 u64p, u64p, f64p]
        L.ko_num_threads.restype = C.c_int
        L.ko_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(f64p)


def _u(a):
    return a.ctypes.data_as(u64p)


def _vec(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OCsr:
    """Host CSR (usize indices as in CsrMatrix::from_csr, src/matrix/sparse.rs:26-34)."""

    def __init__(self, n, ncols, row_ptr, col_idx, vals):
        self.n, self.ncols = int(n), int(ncols)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        self.col_idx = np.ascontiguousarray(col_idx, dtype=np.uint64)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.c = KoCsr(self.n, self.ncols, _u(self.row_ptr), _u(self.col_idx), _f(self.vals))

    @property
    def nnz(self):
        return int(self.row_ptr[-1])

    @classmethod
    def from_dense(cls, a):
        a = np.asarray(a, dtype=np.float64)
        n, m = a.shape
        rp = [0]
        ci, v = [], []
        for i in range(n):
            for j in range(m):
                if a[i, j] != 0.0:
                    ci.append(j)
                    v.append(a[i, j])
            rp.append(len(ci))
        return cls(n, m, rp, ci, v)

    def to_dense(self):
        a = np.zeros((self.n, self.ncols))
        for i in range(self.n):
            for p in range(int(self.row_ptr[i]), int(self.row_ptr[i + 1])):
                a[i, int(self.col_idx[p])] += self.vals[p]
        return a

    def ptr(self):
        return C.byref(self.c)

    def rows(self, lo, hi):
        """Row slice [lo,hi) keeping global column indices (a shard)."""
        a, b = int(self.row_ptr[lo]), int(self.row_ptr[hi])
        return OCsr(hi - lo, self.ncols, self.row_ptr[lo:hi + 1] - self.row_ptr[lo], self.col_idx[a:b], self.vals[a:b])


STENCIL_KINDS = {"poisson2d": 0, "convdiff2d": 1, "varcoef27": 2, "poisson3d": 3, "convdiff3d": 4}


def stencil(kind, N, lo=None, hi=None, pe=(0.4, 0.2, 0.1)):
    L = lib()
    k = STENCIL_KINDS[kind] if isinstance(kind, str) else int(kind)
    n = int(L.ko_stencil_dim(k, N))
    lo = 0 if lo is None else int(lo)
    hi = n if hi is None else int(hi)
    rp = np.zeros(hi - lo + 1, dtype=np.uint64)
    nnz = int(L.ko_stencil(k, N, lo, hi, pe[0], pe[1], pe[2], _u(rp), None, None))
    ci = np.zeros(nnz, dtype=np.uint64)
    v = np.zeros(nnz, dtype=np.float64)
    L.ko_stencil(k, N, lo, hi, pe[0], pe[1], pe[2], _u(rp), _u(ci), _f(v))
    return OCsr(hi - lo, n, rp, ci, v)


def dot(x, y, nshards=1):
    x, y = _vec(x), _vec(y)
    return float(lib().ko_dot_sharded(x.size, _f(x), _f(y), nshards))


def norm(x):
    x = _vec(x)
    return float(lib().ko_norm(x.size, _f(x)))


def csum(v):
    v = _vec(v)
    return float(lib().ko_sum(v.size, _f(v)))


def spmv(A, x):
    x = _vec(x)
    y = np.zeros(A.n)
    lib().ko_spmv(A.ptr(), _f(x), _f(y))
    return y


def jacobi_inv_diag(A):
    d = np.zeros(A.n)
    lib().ko_jacobi_setup(A.ptr(), _f(d))
    return d


def ilu0_factor(A):
    lu = np.zeros(A.nnz)
    dp = np.zeros(A.n, dtype=np.uint64)
    iud = np.zeros(A.n)
    bad = C.c_uint64(0)
    st = lib().ko_ilu0_factor(A.ptr(), _f(lu), _u(dp), _f(iud), C.byref(bad))
    return st, lu, dp, iud, int(bad.value)


def ilu0_apply(A, lu, dp, iud, r):
    r = _vec(r)
    z = np.zeros(A.n)
    lib().ko_ilu0_apply(A.ptr(), _f(lu), _u(dp), _f(iud), _f(r), _f(z))
    return z


def levels(A, upper=False):
    lev = np.zeros(A.n, dtype=np.uint64)
    order = np.zeros(A.n, dtype=np.uint64)
    lp = np.zeros(A.n + 2, dtype=np.uint64)
    nl = int(lib().ko_levels(A.ptr(), 1 if upper else 0, _u(lev), _u(order), _u(lp)))
    return nl, lev, order, lp[:nl + 1].copy()


def partition_range(n, p, r):
    lo, hi = C.c_uint64(0), C.c_uint64(0)
    lib().ko_partition_range(n, p, r, C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


def ghost_list(A, lo, hi):
    cnt = int(lib().ko_ghost_list(A.ptr(), lo, hi, None))
    g = np.zeros(cnt, dtype=np.uint64)
    if cnt:
        lib().ko_ghost_list(A.ptr(), lo, hi, _u(g))
    return g


class OPc:
    def __init__(self, handle, n):
        self.h, self.n = handle, n

    @classmethod
    def jacobi(cls, A):
        return cls(lib().ko_pc_create_jacobi(A.ptr()), A.n)

    @classmethod
    def ilu0(cls, A, nblocks=1):
        return cls(lib().ko_pc_create_ilu0(A.ptr(), nblocks), A.n)

    @classmethod
    def asm(cls, A, overlap=0, subdomains=1, inner="ilu0"):
        """AdditiveSchwarz (asm.rs:34-116): subdomains = int p (uniform chunks) or a list of index lists."""
        if isinstance(subdomains, int):
            nsub, ptr, idx = max(subdomains, 1), None, None
        else:
            nsub = len(subdomains)
            ptr = np.zeros(nsub + 1, dtype=np.uint64)
            ptr[1:] = np.cumsum([len(s) for s in subdomains])
            idx = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.uint64) for s in subdomains]) if nsub and int(ptr[-1])
                                       else np.zeros(0, dtype=np.uint64))
        pc = cls(lib().ko_pc_create_asm(A.ptr(), overlap, nsub, _u(ptr) if ptr is not None else None,
                                        _u(idx) if idx is not None else None, {"ilu0": 0, "jacobi": 1}[inner]), A.n)
        pc.nsub = nsub
        return pc

    def asm_blocks(self):
        out = []
        for b in range(self.nsub):
            k = int(lib().ko_pc_asm_block_size(self.h, b))
            idx = np.zeros(max(k, 1), dtype=np.uint64)
            lib().ko_pc_asm_block_indices(self.h, b, _u(idx))
            out.append(idx[:k].copy())
        return out

    @classmethod
    def ilu_literal(cls, dense):
        d = np.ascontiguousarray(dense, dtype=np.float64)
        return cls(lib().ko_pc_create_ilu_literal(d.shape[0], _f(d)), d.shape[0])

    def status(self):
        bad = C.c_uint64(0)
        st = lib().ko_pc_status(self.h, C.byref(bad))
        return st, int(bad.value)

    def apply(self, r):
        r = _vec(r)
        z = np.zeros(self.n)
        lib().ko_pc_apply(self.h, _f(r), _f(z))
        return z

    def block(self, b, m):
        nnz = int(lib().ko_pc_block_nnz(self.h, b))
        lu = np.zeros(nnz)
        dp = np.zeros(m, dtype=np.uint64)
        iud = np.zeros(m)
        rp = np.zeros(m + 1, dtype=np.uint64)
        ci = np.zeros(nnz, dtype=np.uint64)
        lib().ko_pc_block_get(self.h, b, _f(lu), _u(dp), _f(iud), _u(rp), _u(ci))
        return lu, dp, iud, rp, ci

    def __del__(self):
        try:
            if self.h:
                lib().ko_pc_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _h(pc):
    return pc.h if pc is not None else None


def pcg(A, pc, b, x0, tol, max_iters, norm_type=1, nshards=1, hist_cap=0):
    b = _vec(b)
    x = _vec(x0).copy()
    st = KoStats()
    hist = np.zeros(max(hist_cap, 1))
    hl = C.c_uint64(0)
    rc = lib().ko_pcg(A.ptr(), _h(pc), _f(b), _f(x), tol, max_iters, norm_type, nshards,
                      _f(hist), hist_cap, C.byref(hl), C.byref(st))
    return rc, x, st, hist[:min(int(hl.value), hist_cap)]


def submatrix(A, indices):
    """SubmatrixExtract::submatrix (sparse.rs:72-93) -> OCsr; raises IndexError when an index is out of range."""
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    rp = np.zeros(idx.size + 1, dtype=np.uint64)
    nnz = int(lib().ko_submatrix(A.ptr(), _u(idx), idx.size, _u(rp), None, None))
    if nnz == 2 ** 64 - 1:
        raise IndexError("submatrix index out of range")
    ci = np.zeros(nnz, dtype=np.uint64)
    v = np.zeros(nnz, dtype=np.float64)
    lib().ko_submatrix(A.ptr(), _u(idx), idx.size, _u(rp), _u(ci), _f(v))
    return OCsr(idx.size, idx.size, rp, ci, v)


def pcg_sr(A, pc, b, x0, tol, max_iters, norm_type=1, nshards=1, hist_cap=0):
    """Single-reduction (Chronopoulos-Gear) PCG, SURVEY 8(f3)."""
    b = _vec(b)
    x = _vec(x0).copy()
    st = KoStats()
    hist = np.zeros(max(hist_cap, 1))
    hl = C.c_uint64(0)
    rc = lib().ko_pcg_sr(A.ptr(), _h(pc), _f(b), _f(x), tol, max_iters, norm_type, nshards,
                         _f(hist), hist_cap, C.byref(hl), C.byref(st))
    return rc, x, st, hist[:min(int(hl.value), hist_cap)]


def pcg_pipe(A, pc, b, x0, tol, max_iters, norm_type=1, nshards=1, hist_cap=0):
    """Pipelined (Ghysels-Vanroose) PCG, SURVEY 8(f3)."""
    b = _vec(b)
    x = _vec(x0).copy()
    st = KoStats()
    hist = np.zeros(max(hist_cap, 1))
    hl = C.c_uint64(0)
    rc = lib().ko_pcg_pipe(A.ptr(), _h(pc), _f(b), _f(x), tol, max_iters, norm_type, nshards,
                           _f(hist), hist_cap, C.byref(hl), C.byref(st))
    return rc, x, st, hist[:min(int(hl.value), hist_cap)]


GMRES_LITERAL, GMRES_CGS2, GMRES_MGS2, GMRES_BLOCK = 0, 1, 2, 3
MODE_NONE, MODE_LEFT, MODE_RIGHT = 0, 1, 2


def gmres(A, pc, b, x0, restart, tol, max_iters, mode=MODE_LEFT, variant=GMRES_LITERAL, nshards=1):
    b = _vec(b)
    x = _vec(x0).copy()
    st = KoStats()
    rc = lib().ko_gmres(A.ptr(), _h(pc), _f(b), _f(x), restart, tol, max_iters, mode, variant, nshards, C.byref(st))
    return rc, x, st


def fgmres(A, pc, b, x0, restart, tol, max_iters, nshards=1):
    b = _vec(b)
    x = _vec(x0).copy()
    st = KoStats()
    rc = lib().ko_fgmres(A.ptr(), _h(pc), _f(b), _f(x), restart, tol, max_iters, nshards, C.byref(st))
    return rc, x, st


BICG_LITERAL, BICG_TEXTBOOK = 0, 1


def bicgstab(A, pc, b, x0, tol, max_iters, variant=BICG_LITERAL, nshards=1):
    b = _vec(b)
    x = _vec(x0).copy()
    st = KoStats()
    rc = lib().ko_bicgstab(A.ptr(), _h(pc), _f(b), _f(x), tol, max_iters, variant, nshards, C.byref(st))
    return rc, x, st


def num_threads():
    return int(lib().ko_num_threads())


def use_all_cores():
    """OpenMP threads = the cores this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would silently turn the CPU baseline into a single-thread run."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().ko_set_num_threads(max(1, n))
    return num_threads()
