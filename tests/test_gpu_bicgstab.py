"""BiCGStab parity (GPU vs oracle, bit-exact): literal (reference semantics) and textbook (+Jacobi)."""
import numpy as np
import pytest

import oracle_ffi as o

pytestmark = pytest.mark.gpu


def _mk(kind, N, ctx):
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil(kind, N)
    return kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx), o.OCsr(n, n, rp, ci, v)


@pytest.mark.parametrize("kind,N", [("convdiff2d", 24), ("convdiff3d", 12), ("varcoef27", 10), ("poisson2d", 40)])
def test_bicgstab_literal_bit_exact(ctx, kind, N):
    import kryst_b200 as kb
    A, Ao = _mk(kind, N, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    tol_abs = 1e-8 * float(np.linalg.norm(b))          # literal mode: absolute tolerance (bicgstab.rs:98)
    x = np.zeros(Ao.n)
    st = kb.BiCgStabSolver(tol_abs, 2000).solve(A, kb.Jacobi().setup(A), b, x)     # pc is ignored (bicgstab.rs:70)
    rc, xo, so = o.bicgstab(Ao, None, b, np.zeros(Ao.n), tol_abs, 2000, variant=o.BICG_LITERAL)
    assert (st.iterations, st.converged, st.breakdown) == (so.iterations, bool(so.converged), so.breakdown)
    assert st.final_residual == so.final_residual
    assert np.array_equal(x, xo)


@pytest.mark.parametrize("kind,N", [("convdiff2d", 24), ("convdiff3d", 12), ("varcoef27", 10)])
@pytest.mark.parametrize("use_pc", [True, False])
def test_bicgstab_textbook_jacobi_bit_exact(ctx, kind, N, use_pc):
    import kryst_b200 as kb
    A, Ao = _mk(kind, N, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    x = np.zeros(Ao.n)
    st = kb.BiCgStabSolver(1e-8, 2000, textbook=True).solve(A, kb.Jacobi().setup(A) if use_pc else None, b, x)
    rc, xo, so = o.bicgstab(Ao, o.OPc.jacobi(Ao) if use_pc else None, b, np.zeros(Ao.n), 1e-8, 2000, variant=o.BICG_TEXTBOOK)
    assert (st.iterations, st.converged, st.breakdown) == (so.iterations, bool(so.converged), so.breakdown)
    assert st.final_residual == so.final_residual
    assert np.array_equal(x, xo)
    assert st.converged and np.abs(x - 1.0).max() < 1e-5


def test_bicgstab_reference_3x3(ctx):
    """src/solver/bicgstab.rs:303-328"""
    import kryst_b200 as kb
    a = np.array([[4.0 if i == j else float(i + 2 * j + 1) for j in range(3)] for i in range(3)])
    Ao = o.OCsr.from_dense(a)
    A = kb.DeviceCsr.from_csr(3, 3, Ao.row_ptr, Ao.col_idx, Ao.vals, ctx)
    xt = np.array([1.0, 2.0, 3.0])
    x = np.zeros(3)
    st = kb.BiCgStabSolver(1e-10, 100).solve(A, None, a @ xt, x)
    assert st.converged and np.abs(x - xt).max() < 1e-8
    rc, xo, so = o.bicgstab(Ao, None, a @ xt, np.zeros(3), 1e-10, 100)
    assert np.array_equal(x, xo) and st.iterations == so.iterations


def test_bicgstab_early_exit_and_limits(ctx):
    import kryst_b200 as kb
    A, Ao = _mk("convdiff2d", 16, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    for tol, mi in ((1e3, 50), (1e-30, 7), (1e-8, 0), (0.5, 50)):
        x = np.zeros(Ao.n)
        st = kb.BiCgStabSolver(tol, mi).solve(A, None, b, x)
        rc, xo, so = o.bicgstab(Ao, None, b, np.zeros(Ao.n), tol, mi)
        assert (st.iterations, st.converged, st.breakdown) == (so.iterations, bool(so.converged), so.breakdown), (tol, mi)
        assert st.final_residual == so.final_residual and np.array_equal(x, xo)
