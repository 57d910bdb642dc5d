"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol declared in
include/kryst_b200.h, and refuses (loudly) to compute without a GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "kryst_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = header_symbols()
    for must in ("kb_csr_create", "kb_csr_matvec", "kb_pc_create_jacobi", "kb_pc_create_ilu0", "kb_pc_apply", "kb_pcg_solve",
                 "kb_gmres_solve", "kb_bicgstab_solve", "kb_comm_all_reduce", "kb_partition_range", "kb_dot", "kb_norm"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(os.path.join(ROOT, "kryst_b200", "libkryst_b200.so"))
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_header(built):
    from kryst_b200 import _ffi
    assert sorted(_ffi.SIGNATURES) == header_symbols()
    assert _ffi.lib().kb_abi_version() == 1


def test_partition_range_matches_reference_formula(built):
    import kryst_b200 as kb
    for n, p in ((10, 3), (16, 4), (7, 8), (16777216, 8), (56623104, 8), (5, 1)):
        chunk = (n + p - 1) // p
        for r in range(p):
            assert kb.partition_range(n, p, r) == (min(r * chunk, n), min((r + 1) * chunk, n))
    assert kb.partition_range(16777216, 8, 3) == (3 * 2097152, 4 * 2097152)      # 32 planes of 256^2
    assert kb.partition_range(56623104, 8, 7) == (7 * 7077888, 56623104)          # 48 planes of 384^2


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import kryst_b200 as kb
    with pytest.raises(kb.KError) as e:
        kb.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kryst_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_ffi" not in txt and "kryst_oracle" not in txt and "libkryst_oracle" not in txt, f


def test_stencil_generators_match_oracle_twin(built):
    import numpy as np
    import oracle_ffi as o
    from kryst_b200 import stencils
    for kind, N in (("poisson2d", 9), ("convdiff2d", 9), ("varcoef27", 5), ("poisson3d", 5), ("convdiff3d", 5)):
        A = o.stencil(kind, N)
        n, rp, ci, v = stencils.stencil(kind, N)
        assert n == A.n and np.array_equal(rp, A.row_ptr) and np.array_equal(ci, A.col_idx) and np.array_equal(v, A.vals)
        lo, hi = n // 3, n - 2
        As = o.stencil(kind, N, lo, hi)
        _, rp, ci, v = stencils.stencil(kind, N, lo, hi)
        assert np.array_equal(rp, As.row_ptr) and np.array_equal(ci, As.col_idx) and np.array_equal(v, As.vals)
    # nnz formulas of SURVEY §8
    assert o.stencil("poisson2d", 12).nnz == 5 * 144 - 4 * 12
    assert o.stencil("poisson3d", 6).nnz == 7 * 216 - 6 * 36
    assert o.stencil("varcoef27", 6).nnz == (3 * 6 - 2) ** 3
    # symmetric, diagonally dominant variable-coefficient operator
    a = o.stencil("varcoef27", 4).to_dense()
    assert np.array_equal(a, a.T) and np.all(np.diag(a) >= np.abs(a - np.diag(np.diag(a))).sum(axis=1) - 1e-12)


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/kryst_b200.h must compile as strict C99 (no C++/torch types)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include "kryst_b200.h"\nint main(void) { kb_stats s; kb_profile p; (void)s; (void)p; return KB_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-fsyntax-only", str(src)])


def test_integration_md_rust_block_declares_every_header_symbol():
    """INTEGRATION.md section 2 is the binding a kryst maintainer would paste: it must not drift from the header."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "kryst_b200.h")).read()
    declared = set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", header))
    rust = set(re.findall(r"pub fn (kb_[a-z0-9_]+)", open(os.path.join(root, "INTEGRATION.md")).read()))
    assert declared - rust == set(), sorted(declared - rust)
    assert rust - declared == set(), sorted(rust - declared)
