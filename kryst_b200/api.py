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
KB_FLAG_HISTORY, KB_FLAG_MONITOR, KB_FLAG_BLOCK_ORTH, KB_FLAG_PIPELINED = 32, 64, 128, 256


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

    def comm_dot(self, a, b):
        """Comm::dot (parallel/mod.rs:19-22) / DistributedInnerProduct::dot (wrappers.rs:143-149): every rank passes its slice."""
        a, b = _host_vec(a), _host_vec(b)
        out = C.c_double(0.0)
        _check(_ffi.lib().kb_comm_dot(self._h, a.size, _f(a), _f(b), C.byref(out)))
        return out.value

    def comm_norm(self, x):
        """DistributedInnerProduct::norm (wrappers.rs:150-155)."""
        x = _host_vec(x)
        out = C.c_double(0.0)
        _check(_ffi.lib().kb_comm_norm(self._h, x.size, _f(x), C.byref(out)))
        return out.value

    def scatter(self, global_arr, out, root=0):
        """Comm::scatter(global, out, root) (parallel/mod.rs:9-12; mpi_comm.rs:74-84): equal chunks of len(out)."""
        out_c = np.ascontiguousarray(out)
        g = np.ascontiguousarray(global_arr, dtype=out_c.dtype) if global_arr is not None else None
        _check(_ffi.lib().kb_comm_scatter(self._h, g.ctypes.data if g is not None and g.size else None, out_c.nbytes,
                                          out_c.ctypes.data if out_c.size else None, int(root)))
        if out_c is not out:
            out[...] = out_c
        return out

    def gather(self, local, root=0):
        """Comm::gather(local, &mut out, root) (parallel/mod.rs:13-16; mpi_comm.rs:91-109): the root gets size*len(local)
        entries in rank order, every other rank an empty array."""
        loc = np.ascontiguousarray(local)
        out = np.zeros(loc.size * self.size(), dtype=loc.dtype) if self.rank() == root else np.zeros(0, dtype=loc.dtype)
        _check(_ffi.lib().kb_comm_gather(self._h, loc.ctypes.data if loc.size else None, loc.nbytes,
                                         out.ctypes.data if out.size else None, int(root)))
        return out

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

    def spmv_x_staged(self):
        """Non-zero when the SpMV stages the x tiles of every chunk in shared memory (kb_spmv_xtile.cuh)."""
        return int(_ffi.lib().kb_csr_spmv_x_staged(self._h))

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


class AdditiveSchwarz(_Pc):
    """AdditiveSchwarz::new(overlap, subdomains) (src/preconditioner/asm.rs:34-36) on one GPU.

    `subdomains`: list of index lists (global rows, distinct inside a block), or an int p for p uniform row chunks
    (asm.rs:46-57, the reference's `Vec::with_capacity(p)` idiom).  setup() extracts every block with
    SubmatrixExtract on the device and factors it; apply() sums the block results in subdomain order
    (asm.rs:76-116).  inner = "ilu0" (one ILU(0) application per block) or "jacobi".  overlap = 0 is the reference
    (it stores the field and never reads it); overlap = k grows every set by k layers of graph neighbours."""

    def __init__(self, overlap=0, subdomains=1, inner="ilu0"):
        super().__init__()
        self.overlap, self.subdomains, self.inner = int(overlap), subdomains, inner

    def setup(self, a):
        self.close()
        if isinstance(self.subdomains, int):
            nsub, ptr, idx = max(self.subdomains, 1), None, None
        else:
            nsub = len(self.subdomains)
            ptr = np.zeros(nsub + 1, dtype=np.uint64)
            ptr[1:] = np.cumsum([len(sd) for sd in self.subdomains])
            idx = np.ascontiguousarray(np.concatenate([np.asarray(sd, dtype=np.uint64) for sd in self.subdomains])
                                       if nsub and int(ptr[-1]) else np.zeros(0, dtype=np.uint64))
        h = C.c_void_p()
        st = _ffi.lib().kb_pc_create_asm(a.handle, self.overlap, nsub, _u(ptr) if ptr is not None else None,
                                         _u(idx) if idx is not None else None, {"ilu0": 0, "jacobi": 1}[self.inner], C.byref(h))
        if st != 0:
            msg = _ffi.last_error()
            row = int(_ffi.lib().kb_pc_bad_row(h)) if h else 0
            if h:
                _ffi.lib().kb_pc_destroy(h)
            cls = _ERRORS.get(st, SolveError)
            raise ZeroPivot(msg, row) if cls is ZeroPivot else cls(msg)
        self._h, self._a, self._n = h, a, a.nrows()
        return self

    def blocks(self):
        """Index lists actually used (after overlap growth)."""
        out = []
        for b in range(int(_ffi.lib().kb_pc_asm_num_blocks(self._h))):
            k = int(_ffi.lib().kb_pc_asm_block_size(self._h, b))
            idx = np.zeros(max(k, 1), dtype=np.uint64)
            _check(_ffi.lib().kb_pc_asm_block_indices(self._h, b, _u(idx)))
            out.append(idx[:k].copy())
        return out


class PC:
    """PC<T> (src/context/pc_context.rs:36-76): configuration values + the factory the reference lacks.

        PC.Jacobi, PC.Ilu0, PC.Ilup(fill), PC.BlockJacobi(blocks), PC.AdditiveSchwarz(overlap=0, subdomains=1), ...
        pc = PC.Ilu0.build(a)          # -> device preconditioner handle (kb_pc_create_from_spec)
    Variants that are not on the device path (Ssor, Ilut, Chebyshev, ApproxInv, Multicolor, AMG, Ilup with fill > 0)
    raise Unsupported from build()."""
    KINDS = ("Jacobi", "Ssor", "Ilu0", "Ilup", "Ilut", "Chebyshev", "ApproxInv", "BlockJacobi", "Multicolor", "AMG", "AdditiveSchwarz")

    def __init__(self, kind, fill=0, droptol=0.0, overlap=0, blocks=None, nblocks=0):
        self.kind, self.fill, self.droptol, self.overlap, self.blocks, self.nblocks = kind, fill, droptol, overlap, blocks, nblocks

    def __repr__(self):
        return "PC::%s" % self.kind

    def build(self, a):
        spec = _ffi.KbPcSpec()
        spec.kind = self.KINDS.index(self.kind)
        spec.fill, spec.droptol, spec.overlap = int(self.fill), float(self.droptol), int(self.overlap)
        keep = None
        if self.blocks is not None:
            nb = len(self.blocks)
            ptr = np.zeros(nb + 1, dtype=np.uint64)
            ptr[1:] = np.cumsum([len(b) for b in self.blocks])
            idx = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.uint64) for b in self.blocks]) if nb and int(ptr[-1])
                                       else np.zeros(0, dtype=np.uint64))
            spec.nblocks, spec.block_ptr, spec.block_idx = nb, _u(ptr), _u(idx)
            keep = (ptr, idx)
        else:
            spec.nblocks = int(self.nblocks)
        h = C.c_void_p()
        st = _ffi.lib().kb_pc_create_from_spec(a.handle, C.byref(spec), C.byref(h))
        del keep
        if st != 0:
            msg = _ffi.last_error()
            row = int(_ffi.lib().kb_pc_bad_row(h)) if h else 0
            if h:
                _ffi.lib().kb_pc_destroy(h)
            cls = _ERRORS.get(st, SolveError)
            raise ZeroPivot(msg, row) if cls is ZeroPivot else cls(msg)
        pc = _Pc()
        pc._h, pc._a, pc._n = h, a, a.nrows()
        return pc


PC.Jacobi = PC("Jacobi")
PC.Ssor = PC("Ssor")
PC.Ilu0 = PC("Ilu0")
PC.AMG = PC("AMG")
PC.Ilup = staticmethod(lambda fill: PC("Ilup", fill=fill))
PC.Ilut = staticmethod(lambda fill, droptol: PC("Ilut", fill=fill, droptol=droptol))
PC.Chebyshev = staticmethod(lambda degree, emin=None, emax=None: PC("Chebyshev", fill=degree))
PC.BlockJacobi = staticmethod(lambda blocks: PC("BlockJacobi", blocks=blocks))
PC.AdditiveSchwarz = staticmethod(lambda overlap=0, subdomains=1: PC("AdditiveSchwarz", overlap=overlap,
                                  blocks=None if isinstance(subdomains, int) else subdomains,
                                  nblocks=subdomains if isinstance(subdomains, int) else 0))


def ksp_solve(kind_index, a, pc, tol, max_it, restart, b, x, flags=0):
    """KspContext::solve_context through the C ABI (kb_ksp_solve)."""
    base = _SolverBase()
    base.flags = flags
    pb, px, fl, keep = base._solve_args(a, b, x)
    k = _ffi.KbKsp(int(kind_index), float(tol), int(max_it), int(restart))
    st = KbStats()
    rc = _ffi.lib().kb_ksp_solve(a.handle, _pc_handle(pc), C.byref(k), pb, px, fl, C.byref(st))
    stats = SolveStats(st.iterations, st.final_residual, st.converged, st.breakdown)
    _check(rc)
    if keep[2] is not None:
        keep[2][...] = keep[1]
    return stats


def get_history(a):
    """residual history of the last KB_FLAG_HISTORY / KB_FLAG_MONITOR solve on operator `a` (kb_get_history)."""
    n = C.c_uint64(0)
    _check(_ffi.lib().kb_get_history(a.handle, None, 0, C.byref(n)))
    out = np.zeros(max(int(n.value), 1))
    _check(_ffi.lib().kb_get_history(a.handle, _f(out), out.size, C.byref(n)))
    return out[:min(int(n.value), out.size)].copy()


class _MonitorScope:
    """Installs a host observer on an operator for the duration of one solve (kb_set_monitor + KB_FLAG_MONITOR)."""

    def __init__(self, a, fn):
        self.a, self.fn = a, fn
        self.cb = _ffi.MONITOR_FN(lambda it, res, _u: fn(int(it), float(res))) if fn else None

    def __enter__(self):
        if self.cb:
            _check(_ffi.lib().kb_set_monitor(self.a.handle, self.cb, None))
        return KB_FLAG_MONITOR if self.cb else 0

    def __exit__(self, *exc):
        if self.cb:
            _ffi.lib().kb_set_monitor(self.a.handle, _ffi.MONITOR_FN(0), None)
        return False


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
        self.pipelined = False
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

    def with_pipelined(self, flag=True):
        """Extension: pipelined (Ghysels-Vanroose) recurrences - the one reduction of an iteration is sent
        before that iteration's SpMV and received after it (KB_FLAG_PIPELINED; Jacobi / no pc only).  Parity is checked
        against the oracle's restatement of this variant."""
        self.pipelined = bool(flag)
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
        if self.pipelined:
            flags |= KB_FLAG_PIPELINED
        cap = 0
        hist = None
        if self.record_history or self.monitor:
            cap = self.history_capacity if self.history_capacity is not None else min(self.max_iters + 1, 1 << 20)
            hist = np.zeros(max(cap, 1))
        st = KbStats()
        hl = C.c_uint64(0)
        # with_monitor: slow mode - the observer runs on the host after every iteration, while the solve is in flight
        # (pcg.rs:143-145,196-198); without one the history is copied out once at the end
        with _MonitorScope(a, self.monitor) as mflag:
            rc = _ffi.lib().kb_pcg_solve(a.handle, _pc_handle(pc), pb, px, self.tol, self.max_iters, int(self.norm_type), flags | mflag,
                                         _f(hist) if hist is not None else None, cap, C.byref(hl), C.byref(st))
        if hist is not None:
            k = min(int(hl.value), cap)
            self.residual_history.extend(hist[:k].tolist())
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

    block_orthogonalisation = False

    def with_block_orthogonalisation(self, on=True):
        """Extension (the idea of src/solver/pca_gmres.rs:172-229): one classical Gram-Schmidt pass, every inner product of
        the Arnoldi step reduced together (KB_FLAG_BLOCK_ORTH); the default stays CGS2."""
        self.block_orthogonalisation = bool(on)
        return self

    record_history = False
    monitor = None

    def with_monitor(self, f):
        self.monitor = f
        return self

    def solve(self, a, pc, b, x):
        pb, px, flags, keep = self._solve_args(a, b, x)
        st = KbStats()
        if self.record_history:
            flags |= KB_FLAG_HISTORY
        if self.block_orthogonalisation:
            flags |= KB_FLAG_BLOCK_ORTH
        with _MonitorScope(a, self.monitor) as mflag:
            rc = _ffi.lib().kb_gmres_solve(a.handle, _pc_handle(pc), pb, px, self.restart, self.tol, self.max_iters,
                                           int(self.preconditioning), flags | mflag, C.byref(st))
        if self.record_history or self.monitor:
            self.residual_history = get_history(a).tolist()
        self.last_stats = SolveStats(st.iterations, st.final_residual, st.converged, st.breakdown)
        _check(rc)
        if keep[2] is not None:
            keep[2][...] = keep[1]
        return self.last_stats


class FgmresSolver(_SolverBase):
    """FgmresSolver::new(tol, max_iters, restart) (src/solver/fgmres.rs:52-66); solve_flex at fgmres.rs:114-340."""

    def __init__(self, tol, max_iters, restart):
        self.tol, self.max_iters, self.restart = float(tol), int(max_iters), int(restart)

    record_history = True       # fgmres.rs:48,290: residual_history is always kept
    monitor = None

    def with_monitor(self, f):      # fgmres.rs:92-97
        self.monitor = f
        return self

    def clear_history(self):        # fgmres.rs:99-101
        self.residual_history = []

    def solve_flex(self, a, pc, b, x):
        pb, px, flags, keep = self._solve_args(a, b, x)
        st = KbStats()
        if self.record_history:
            flags |= KB_FLAG_HISTORY
        with _MonitorScope(a, self.monitor) as mflag:
            rc = _ffi.lib().kb_fgmres_solve(a.handle, _pc_handle(pc), pb, px, self.restart, self.tol, self.max_iters, flags | mflag, C.byref(st))
        if self.record_history or self.monitor:
            self.residual_history = getattr(self, "residual_history", []) + get_history(a).tolist()
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

    record_history = False

    def solve(self, a, pc, b, x):
        pb, px, flags, keep = self._solve_args(a, b, x)
        st = KbStats()
        if self.record_history:
            flags |= KB_FLAG_HISTORY
        rc = _ffi.lib().kb_bicgstab_solve(a.handle, _pc_handle(pc), pb, px, self.tol, self.max_iters, flags, C.byref(st))
        if self.record_history:
            self.residual_history = get_history(a).tolist()
        self.last_stats = SolveStats(st.iterations, st.final_residual, st.converged, st.breakdown)
        _check(rc)
        if keep[2] is not None:
            keep[2][...] = keep[1]
        return self.last_stats
