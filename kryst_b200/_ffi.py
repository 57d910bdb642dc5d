"""ctypes binding of libkryst_b200.so (the C ABI in include/kryst_b200.h).

There is no fallback: if the CUDA library is missing or fails to load, importing this
module raises.  Nothing here (or anywhere in kryst_b200/) touches the CPU oracle.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkryst_b200.so")

u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
KB_PROF_CLASSES = 12


class KbStats(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("final_residual", C.c_double),
                ("converged", C.c_int32), ("breakdown", C.c_int32)]


class KbPcSpec(C.Structure):      # kb_pc_spec
    _fields_ = [("kind", C.c_int32), ("fill", C.c_uint64), ("droptol", C.c_double), ("overlap", C.c_uint64),
                ("nblocks", C.c_uint64), ("block_ptr", u64p), ("block_idx", u64p)]


class KbKsp(C.Structure):         # kb_ksp
    _fields_ = [("kind", C.c_int32), ("tol", C.c_double), ("max_it", C.c_uint64), ("restart", C.c_uint64)]


MONITOR_FN = C.CFUNCTYPE(None, C.c_uint64, C.c_double, C.c_void_p)


class KbProfile(C.Structure):
    _fields_ = [("launches", C.c_uint64 * KB_PROF_CLASSES), ("ms", C.c_double * KB_PROF_CLASSES)]


# every symbol include/kryst_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "kb_abi_version": (C.c_int, []),
    "kb_last_error": (C.c_char_p, []),
    "kb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "kb_ctx_destroy": (C.c_int, [C.c_void_p]),
    "kb_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "kb_ctx_device": (C.c_int, [C.c_void_p]),
    "kb_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "kb_ctx_launch_count": (C.c_uint64, [C.c_void_p]),
    "kb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "kb_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "kb_comm_rank": (C.c_int, [C.c_void_p]),
    "kb_comm_size": (C.c_int, [C.c_void_p]),
    "kb_comm_barrier": (C.c_int, [C.c_void_p]),
    "kb_comm_all_reduce": (C.c_int, [C.c_void_p, C.c_double, f64p]),
    "kb_comm_dot": (C.c_int, [C.c_void_p, C.c_uint64, f64p, f64p, f64p]),
    "kb_comm_norm": (C.c_int, [C.c_void_p, C.c_uint64, f64p, f64p]),
    "kb_comm_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]),
    "kb_comm_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]),
    "kb_partition_range": (None, [C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p]),
    "kb_csr_create": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, u64p, u64p, f64p, C.POINTER(C.c_void_p)]),
    "kb_csr_create_dist": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, f64p, C.POINTER(C.c_void_p)]),
    "kb_csr_destroy": (C.c_int, [C.c_void_p]),
    "kb_csr_nrows": (C.c_uint64, [C.c_void_p]),
    "kb_csr_ncols": (C.c_uint64, [C.c_void_p]),
    "kb_csr_nnz": (C.c_uint64, [C.c_void_p]),
    "kb_csr_matvec": (C.c_int, [C.c_void_p, f64p, f64p]),
    "kb_csr_matvec_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "kb_csr_num_ghosts": (C.c_uint64, [C.c_void_p]),
    "kb_csr_get_ghosts": (C.c_int, [C.c_void_p, u64p]),
    "kb_csr_spmv_kernel_kind": (C.c_int, [C.c_void_p]),
    "kb_csr_spmv_x_staged": (C.c_int, [C.c_void_p]),
    "kb_csr_submatrix": (C.c_int, [C.c_void_p, u64p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "kb_csr_download": (C.c_int, [C.c_void_p, u64p, u64p, f64p]),
    "kb_dot": (C.c_int, [C.c_void_p, C.c_uint64, f64p, f64p, f64p]),
    "kb_norm": (C.c_int, [C.c_void_p, C.c_uint64, f64p, f64p]),
    "kb_pc_create_jacobi": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "kb_pc_create_ilu0": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "kb_pc_create_asm": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, u64p, u64p, C.c_int, C.POINTER(C.c_void_p)]),
    "kb_pc_asm_num_blocks": (C.c_uint64, [C.c_void_p]),
    "kb_pc_asm_block_size": (C.c_uint64, [C.c_void_p, C.c_uint64]),
    "kb_pc_asm_block_indices": (C.c_int, [C.c_void_p, C.c_uint64, u64p]),
    "kb_pc_create_from_spec": (C.c_int, [C.c_void_p, C.POINTER(KbPcSpec), C.POINTER(C.c_void_p)]),
    "kb_ksp_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(KbKsp), C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(KbStats)]),
    "kb_set_monitor": (C.c_int, [C.c_void_p, MONITOR_FN, C.c_void_p]),
    "kb_get_history": (C.c_int, [C.c_void_p, f64p, C.c_uint64, u64p]),
    "kb_pc_apply": (C.c_int, [C.c_void_p, f64p, f64p]),
    "kb_pc_apply_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "kb_pc_destroy": (C.c_int, [C.c_void_p]),
    "kb_pc_bad_row": (C.c_uint64, [C.c_void_p]),
    "kb_pc_get_inv_diag": (C.c_int, [C.c_void_p, f64p]),
    "kb_pc_ilu0_get_factors": (C.c_int, [C.c_void_p, f64p, u64p]),
    "kb_pc_ilu0_get_levels": (C.c_int, [C.c_void_p, C.c_int, u64p, u64p, u64p]),
    "kb_pcg_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint64, C.c_int,
                               C.c_uint32, f64p, C.c_uint64, u64p, C.POINTER(KbStats)]),
    "kb_gmres_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_uint64,
                                 C.c_int, C.c_uint32, C.POINTER(KbStats)]),
    "kb_fgmres_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_uint64,
                                  C.c_uint32, C.POINTER(KbStats)]),
    "kb_bicgstab_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint64,
                                    C.c_uint32, C.POINTER(KbStats)]),
    "kb_profile_reset": (C.c_int, [C.c_void_p]),
    "kb_profile_get": (C.c_int, [C.c_void_p, C.POINTER(KbProfile)]),
    "kb_profile_class_name": (C.c_char_p, [C.c_int]),
}

_lib = None


def lib():
    """Load the CUDA library (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "kryst_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.kb_abi_version() != 1:
            raise ImportError("kryst_b200: ABI version mismatch")
        _lib = L
    return _lib


def last_error():
    return lib().kb_last_error().decode("utf-8", "replace")
