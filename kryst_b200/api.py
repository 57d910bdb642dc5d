"""Host-side mirror of kryst's trait API for the Krylov hot path, backed by the CUDA library.

Same names, argument meaning and error behaviour as the reference (paths relative to the kryst
crate root):
  * ``DeviceCsr.from_csr`` / ``matvec`` / ``nrows`` / ``ncols``  <- CsrMatrix::from_csr (src/matrix/sparse.rs:26-47),
    MatVec::matvec (src/core/traits.rs:4-7), Indexing / MatShape (traits.rs:26-35)
  * ``Jacobi`` / ``Ilu0`` with ``setup(a)`` and ``apply(r, z)``    <- Preconditioner (src/preconditioner/mod.rs:8-13)
  * ``PcgSolver(tol, max_iters)``, ``GmresSolver(restart, tol, max_iters)``, ``BiCgStabSolver(tol, max_iters)``
    with ``solve(a, pc, b, x) -> SolveStats``                     <- LinearSolver::solve (src/solver/mod.rs:43-49)
  * ``KError`` subclasses                                          <- src/error.rs:6-19
Vectors are host float64 numpy arrays (the reference's ``V: AsRef<[f64]> + AsMut<[f64]>``); they are
borrowed for the call, ``x`` is in/out and is written only when ``solve`` returns Ok.  CUDA torch
tensors are also accepted and then stay resident in HBM (no host copies).
"""
import ctypes as C
import enum

import numpy as np

from . import _ffi
from ._ffi import KbStats, KbProfile, f64p, u64p

KB_FLAG_DEVICE_PTRS, KB_FLAG_TEXTBOOK, KB_FLAG_PROFILE, KB_FLAG_NO_GRAPH, KB_FLAG_SINGLE_REDUCTION = 1, 2, 4, 8, 16


# ---- KError (src/error.rs:6-19) ----------------------------------------------------------------
class KError(Exception):
    status = 2


class FactorError(KError):
    status = 1


class SolveError(KError):
    status = 2


class IndefiniteMatrix(KError):
    status = 3


class IndefinitePreconditioner(KError):
    status = 4


class ZeroPivot(KError):
    status = 5

    def __init__(self, msg, row=0):
        super().__init__(msg)
        self.row = row


class Unsupported(KError):
    status = 6


_ERRORS = {1: FactorError, 2: SolveError, 3: IndefiniteMatrix, 4: IndefinitePreconditioner, 5: ZeroPivot, 6: Unsupported}


def _check(status, row=None):
    if status == 0:
        return
    msg = _ffi.last_error()
    cls = _ERRORS.get(status, SolveError)
    if cls is ZeroPivot:
        raise ZeroPivot(msg or "zero pivot", row or 0)
    raise cls(msg)


class SolveStats:
    """SolveStats<f64> (src/utils/convergence.rs:9-14)."""

    def __init__(self, iterations, final_residual, converged, breakdown=0):
        self.iterations = int(iterations)
        self.final_residual = float(final_residual)
        self.converged = bool(converged)
        self.breakdown = int(breakdown)

    def __repr__(self):
        return "SolveStats { iterations: %d, final_residual: %r, converged: %s }" % (
            self.iterations, self.final_residual, str(self.converged).lower())


class CgNormType(enum.IntEnum):        # pcg.rs:25
    Preconditioned = 0
    Unpreconditioned = 1
    Natural = 2
    NoNorm = 3


class Preconditioning(enum.IntEnum):   # gmres.rs:28-32
    NoPc = 0
    Left = 1
    Right = 2


# ---- context -----------------------------------------------------------------------------------
class Context:
    """One GPU + the library stream (+ communicator when row-partitioned)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(_ffi.lib().kb_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    @property
    def stream(self):
        """cudaStream_t (int) every kernel of this context is launched on."""
        return int(_ffi.lib().kb_ctx_stream(self._h) or 0)

    def synchronize(self):
        _check(_ffi.lib().kb_ctx_synchronize(self._h))

    def launch_count(self):
        return int(_ffi.lib().kb_ctx_launch_count(self._h))

    # Comm surface (src/parallel/mod.rs:4-35)
    def rank(self):
        return int(_ffi.lib().kb_comm_rank(self._h))

    def size(self):
        return int(_ffi.lib().kb_comm_size(self._h))

    def barrier(self):
        _check(_ffi.lib().kb_comm_barrier(self._h))

    def all_reduce(self, x):
        out = C.c_double(0.0)
        _check(_ffi.lib().kb_comm_all_reduce(self._h, float(x), C.byref(out)))
        return out.value

    def comm_init(self, rank, size, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _check(_ffi.lib().kb_comm_init(self._h, int(rank), int(size), buf))

    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        _check(_ffi.lib().kb_comm_unique_id(buf))
        return buf.raw

    def profile_reset(self):
        _check(_ffi.lib().kb_profile_reset(self._h))

    def profile(self):
        p = KbProfile()
        _check(_ffi.lib().kb_profile_get(self._h, C.byref(p)))
        out = {}
        for k in range(_ffi.KB_PROF_CLASSES):
            if p.launches[k]:
                out[_ffi.lib().kb_profile_class_name(k).decode()] = {"launches": int(p.launches[k]), "ms": float(p.ms[k])}
        return out

    def dot(self, x, y):
        x, y = _host_vec(x), _host_vec(y)
        out = C.c_double(0.0)
        _check(_ffi.lib().kb_dot(self._h, x.size, _f(x), _f(y), C.byref(out)))
        return out.value

    def norm(self, x):
        x = _host_vec(x)
        out = C.c_double(0.0)
        _check(_ffi.lib().kb_norm(self._h, x.size, _f(x), C.byref(out)))
        return out.value

    def close(self):
        if self._h:
            _ffi.lib().kb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def partition_range(n, p, r):
    """chunk = ceil(n/p); rank r owns [r*chunk, min((r+1)*chunk, n))  (src/preconditioner/asm.rs:46-57)."""
    lo, hi = C.c_uint64(0), C.c_uint64(0)
    _ffi.lib().kb_partition_range(int(n), int(p), int(r), C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


def _f(a):
    return a.ctypes.data_as(f64p)


def _u(a):
    return a.ctypes.data_as(u64p)


def _host_vec(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _is_device_tensor(v):
    return hasattr(v, "data_ptr") and hasattr(v, "is_cuda") and v.is_cuda


def _vec_arg(v, n, name, writable=False):
    """-> (pointer, is_device, keepalive, writeback)"""
    if _is_device_tensor(v):
        if str(v.dtype) != "torch.float64" or not v.is_contiguous() or v.numel() != n:
            raise SolveError("%s: device tensor must be contiguous float64 of length %d" % (name, n))
        return C.c_void_p(v.data_ptr()), True, v, None
    if isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags["C_CONTIGUOUS"] and (v.flags["WRITEABLE"] or not writable):
        if v.size != n:
            raise SolveError("%s: length %d, expected %d" % (name, v.size, n))
        return C.c_void_p(v.ctypes.data), False, v, None
    a = np.ascontiguousarray(v, dtype=np.float64)
    if a.size != n:
        raise SolveError("%s: length %d, expected %d" % (name, a.size, n))
    if writable:
        a = a.copy()
        return C.c_void_p(a.ctypes.data), False, a, v
    return C.c_void_p(a.ctypes.data), False, a, None


# ---- operator ----------------------------------------------------------------------------------
class DeviceCsr:
    """Device-resident CSR operator: implements MatVec<Vec<f64>> + Indexing + MatShape."""

    def __init__(self, handle, ctx):
        self._h = handle
        self.ctx = ctx

    @classmethod
    def from_csr(cls, nrows, ncols, row_ptr, col_idx, values, ctx=None):
        """CsrMatrix::from_csr(nrows, ncols, row_ptr, col_idx, values) (sparse.rs:26-47); usize indices."""
        ctx = ctx or default_context()
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        ci = np.ascontiguousarray(col_idx, dtype=np.uint64)
        v = np.ascontiguousarray(values, dtype=np.float64)
        if rp.size != nrows + 1 or ci.size != v.size or (rp.size and int(rp[-1]) != ci.size):
            raise SolveError("from_csr: inconsistent array lengths")
        h = C.c_void_p()
        _check(_ffi.lib().kb_csr_create(ctx.handle, int(nrows), int(ncols), _u(rp), _u(ci), _f(v), C.byref(h)))
        return cls(h, ctx)

    @classmethod
    def from_csr_shard(cls, n_global, row_lo, row_hi, row_ptr, col_idx, values, ctx):
        """Row-block shard [row_lo,row_hi) of a square operator; col_idx are global columns."""
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        ci = np.ascontiguousarray(col_idx, dtype=np.uint64)
        v = np.ascontiguousarray(values, dtype=np.float64)
        h = C.c_void_p()
        _check(_ffi.lib().kb_csr_create_dist(ctx.handle, int(n_global), int(row_lo), int(row_hi), _u(rp), _u(ci), _f(v), C.byref(h)))
        return cls(h, ctx)

    @property
    def handle(self):
        return self._h

    def nrows(self):
        return int(_ffi.lib().kb_csr_nrows(self._h))

    def ncols(self):
        return int(_ffi.lib().kb_csr_ncols(self._h))

    def nnz(self):
        return int(_ffi.lib().kb_csr_nnz(self._h))

    def spmv_kernel_kind(self):
        return int(_ffi.lib().kb_csr_spmv_kernel_kind(self._h))

    def ghosts(self):
        n = int(_ffi.lib().kb_csr_num_ghosts(self._h))
        g = np.zeros(n, dtype=np.uint64)
        _check(_ffi.lib().kb_csr_get_ghosts(self._h, _u(g)))
        return g

    def submatrix(self, indices):
        """SubmatrixExtract::submatrix(&self, indices) (sparse.rs:72-93): out[i][j] = a[indices[i]][indices[j]],
        stored zeros dropped, built on the device - what AdditiveSchwarz::setup calls per subdomain (asm.rs:58-65)."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        h = C.c_void_p()
        _check(_ffi.lib().kb_csr_submatrix(self._h, _u(idx), int(idx.size), C.byref(h)))
        return DeviceCsr(h, self.ctx)

    def to_csr(self):
        """(row_ptr, col_idx, values) as CsrMatrix::from_csr takes them (sparse.rs:26-34), read back from the device."""
        rp = np.zeros(self.nrows() + 1, dtype=np.uint64)
        ci = np.zeros(self.nnz(), dtype=np.uint64)
        v = np.zeros(self.nnz(), dtype=np.float64)
        _check(_ffi.lib().kb_csr_download(self._h, _u(rp), _u(ci), _f(v)))
        return rp, ci, v

    def matvec(self, x, y):
        """y <- A x  (MatVec::matvec)."""
        nx = self.nrows() if self.ctx.size() > 1 else self.ncols()
        px, dx, kx, _ = _vec_arg(x, nx, "x")
        py, dy, ky, wb = _vec_arg(y, self.nrows(), "y", writable=True)
        if dx != dy:
            raise SolveError("matvec: x and y must both be host arrays or both device tensors")
        if dx:
            _check(_ffi.lib().kb_csr_matvec_device(self._h, px, py))
        else:
            _check(_ffi.lib().kb_csr_matvec(self._h, C.cast(px, f64p), C.cast(py, f64p)))
            if wb is not None:
                wb[...] = ky
        return y

    def close(self):
        if self._h:
            _ffi.lib().kb_csr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- preconditioners ----------------------------------------------------------------------------
class _Pc:
    _create = None

    def __init__(self):
        self._h = C.c_void_p()
        self._a = None

    @property
    def handle(self):
        return self._h

    def setup(self, a):
        """Preconditioner::setup(&mut self, &M)."""
        self.close()
        h = C.c_void_p()
        st = getattr(_ffi.lib(), self._create)(a.handle, C.byref(h))
        if st != 0:
            msg = _ffi.last_error()
            row = int(_ffi.lib().kb_pc_bad_row(h)) if h else 0
            if h:
                _ffi.lib().kb_pc_destroy(h)
            cls = _ERRORS.get(st, SolveError)
            raise ZeroPivot(msg, row) if cls is ZeroPivot else cls(msg)
        self._h, self._a, self._n = h, a, a.nrows()
        return self

    def apply(self, r, z):
        """Preconditioner::apply(&self, r, z): z = M^-1 r."""
        if not self._h:
            raise SolveError("preconditioner used before setup()")
        n = self._n
        pr, dr, kr, _ = _vec_arg(r, n, "r")
        pz, dz, kz, wb = _vec_arg(z, n, "z", writable=True)
        if dr != dz:
            raise SolveError("apply: r and z must both be host arrays or both device tensors")
        if dr:
            _check(_ffi.lib().kb_pc_apply_device(self._h, pr, pz))
        else:
            _check(_ffi.lib().kb_pc_apply(self._h, C.cast(pr, f64p), C.cast(pz, f64p)))
            if wb is not None:
                wb[...] = kz
        return z

    def close(self):
        if self._h:
            _ffi.lib().kb_pc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Jacobi(_Pc):
    """Jacobi<f64>: M^-1 = D^-1 (src/preconditioner/jacobi.rs:26-95)."""
    _create = "kb_pc_create_jacobi"

    @property
    def inv_diag(self):
        out = np.zeros(self._n)
        _check(_ffi.lib().kb_pc_get_inv_diag(self._h, _f(out)))
        return out


class Ilu0(_Pc):
    """ILU(0) on the CSR pattern with level-scheduled triangular solves (replaces the dense
    src/preconditioner/ilu.rs:32-122).  On a row-block shard this is block-Jacobi ILU(0):
    AdditiveSchwarz with overlap 0 over the chunk partition (src/preconditioner/asm.rs:34-57)."""
    _create = "kb_pc_create_ilu0"

    @property
    def inv_diag(self):
        out = np.zeros(self._n)
        _check(_ffi.lib().kb_pc_get_inv_diag(self._h, _f(out)))
        return out

    def factors(self, nnz):
        lu = np.zeros(nnz)
        dp = np.zeros(self._n, dtype=np.uint64)
        _check(_ffi.lib().kb_pc_ilu0_get_factors(self._h, _f(lu), _u(dp)))
        return lu, dp

    def levels(self, upper=False):
        n = self._n
        nl = C.c_uint64(0)
        lp = np.zeros(n + 2, dtype=np.uint64)
        order = np.zeros(max(n, 1), dtype=np.uint64)
        _check(_ffi.lib().kb_pc_ilu0_get_levels(self._h, 1 if upper else 0, C.byref(nl), _u(lp), _u(order)))
        k = int(nl.value)
        return k, lp[:k + 1].copy(), order[:n].copy()


BlockJacobiIlu0 = Ilu0


# ---- solvers -------------------------------------------------------------------------------------
def _pc_handle(pc):
    if pc is None:
        return None
    if not pc.handle:
        raise SolveError("preconditioner used before setup()")
    return pc.handle


class _SolverBase:
    flags = 0

    def _solve_args(self, a, b, x):
        n = a.nrows()
        pb, db, kb_, _ = _vec_arg(b, n, "b")
        px, dx, kx, wb = _vec_arg(x, n, "x", writable=True)
        if db != dx:
            raise SolveError("solve: b and x must both be host arrays or both device tensors")
        flags = self.flags | (KB_FLAG_DEVICE_PTRS if dx else 0)
        return pb, px, flags, (kb_, kx, wb)


class PcgSolver(_SolverBase):
    """PcgSolver::new(tol, max_iters) (src/solver/pcg.rs:50-90); solve at pcg.rs:114-222."""

    def __init__(self, tol, max_iters):
        self.tol, self.max_iters = float(tol), int(max_iters)
        self.norm_type = CgNormType.Unpreconditioned
        self.single_reduction = False
        self.fused_reduction = False
        self.radius = None
        self.obj_target = None
        self.residual_history = []
        self.monitor = None
        self.record_history = True
        self.history_capacity = None

    def with_norm(self, norm_type):
        self.norm_type = CgNormType(norm_type)
        return self

    def with_single_reduction(self, flag):
        # pcg.rs:66-69: in the reference this flag only swaps Rayon dot for a serial loop
        # (pcg.rs:151-160); the arithmetic is unchanged, and so it is here.
        self.single_reduction = bool(flag)
        return self

    def with_fused_reduction(self, flag=True):
        """Extension: Chronopoulos-Gear recurrences, ONE reduction (one all-reduce on shards) per
        iteration - what the reference's flag name promises but its code does not do.  Not the reference's
        arithmetic: parity is checked against the oracle's restatement of this variant (Jacobi / no pc only)."""
        self.fused_reduction = bool(flag)
        return self

    def with_radius(self, radius):
        # pcg.rs:72-75 stores the value; solve (pcg.rs:114-222) never reads it - same here
        self.radius = float(radius)
        return self

    def with_obj_target(self, obj):
        # pcg.rs:77-80: stored, never read by solve
        self.obj_target = float(obj)
        return self

    def with_monitor(self, f):
        self.monitor = f
        return self

    def clear_history(self):
        self.residual_history = []

    def solve(self, a, pc, b, x):
        pb, px, flags, keep = self._solve_args(a, b, x)
        if self.fused_reduction:
            flags |= KB_FLAG_SINGLE_REDUCTION
        cap = 0
        hist = None
        if self.record_history or self.monitor:
            cap = self.history_capacity if self.history_capacity is not None else min(self.max_iters + 1, 1 << 20)
            hist = np.zeros(max(cap, 1))
        st = KbStats()
        hl = C.c_uint64(0)
        rc = _ffi.lib().kb_pcg_solve(a.handle, _pc_handle(pc), pb, px, self.tol, self.max_iters, int(self.norm_type), flags,
                                     _f(hist) if hist is not None else None, cap, C.byref(hl), C.byref(st))
        if hist is not None:
            k = min(int(hl.value), cap)
            self.residual_history.extend(hist[:k].tolist())
            if self.monitor:
                for i in range(k):
                    self.monitor(i, float(hist[i]))
        self.last_stats = SolveStats(st.iterations, st.final_residual, st.converged, st.breakdown)
        _check(rc)
        if keep[2] is not None:
            keep[2][...] = keep[1]
        return self.last_stats


class GmresSolver(_SolverBase):
    """GmresSolver::new(restart, tol, max_iters) (src/solver/gmres.rs:49-60); solve at gmres.rs:216-402."""

    def __init__(self, restart, tol, max_iters):
        self.restart, self.tol, self.max_iters = int(restart), float(tol), int(max_iters)
        self.preconditioning = Preconditioning.Left   # gmres.rs:53

    def with_preconditioning(self, mode):
        self.preconditioning = Preconditioning(mode)
        return self

    def solve(self, a, pc, b, x):
        pb, px, flags, keep = self._solve_args(a, b, x)
        st = KbStats()
        rc = _ffi.lib().kb_gmres_solve(a.handle, _pc_handle(pc), pb, px, self.restart, self.tol, self.max_iters,
                                       int(self.preconditioning), flags, C.byref(st))
        self.last_stats = SolveStats(st.iterations, st.final_residual, st.converged, st.breakdown)
        _check(rc)
        if keep[2] is not None:
            keep[2][...] = keep[1]
        return self.last_stats


class FgmresSolver(_SolverBase):
    """FgmresSolver::new(tol, max_iters, restart) (src/solver/fgmres.rs:52-66); solve_flex at fgmres.rs:114-340."""

    def __init__(self, tol, max_iters, restart):
        self.tol, self.max_iters, self.restart = float(tol), int(max_iters), int(restart)

    def solve_flex(self, a, pc, b, x):
        pb, px, flags, keep = self._solve_args(a, b, x)
        st = KbStats()
        rc = _ffi.lib().kb_fgmres_solve(a.handle, _pc_handle(pc), pb, px, self.restart, self.tol, self.max_iters, flags, C.byref(st))
        self.last_stats = SolveStats(st.iterations, st.final_residual, st.converged, st.breakdown)
        _check(rc)
        if keep[2] is not None:
            keep[2][...] = keep[1]
        return self.last_stats

    solve = solve_flex


class BiCgStabSolver(_SolverBase):
    """BiCgStabSolver::new(tol, max_iters) (src/solver/bicgstab.rs:45-47); solve at bicgstab.rs:69-293.

    Default is the reference's literal behaviour: the preconditioner is ignored (bicgstab.rs:70) and
    ``tol`` is absolute (bicgstab.rs:98,189,281).  ``textbook=True`` selects the right-preconditioned,
    relative-tolerance variant that "BiCGStab + Jacobi" needs."""

    def __init__(self, tol, max_iters, textbook=False):
        self.tol, self.max_iters = float(tol), int(max_iters)
        self.flags = KB_FLAG_TEXTBOOK if textbook else 0

    def solve(self, a, pc, b, x):
        pb, px, flags, keep = self._solve_args(a, b, x)
        st = KbStats()
        rc = _ffi.lib().kb_bicgstab_solve(a.handle, _pc_handle(pc), pb, px, self.tol, self.max_iters, flags, C.byref(st))
        self.last_stats = SolveStats(st.iterations, st.final_residual, st.converged, st.breakdown)
        _check(rc)
        if keep[2] is not None:
            keep[2][...] = keep[1]
        return self.last_stats
