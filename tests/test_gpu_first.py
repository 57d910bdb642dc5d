"""First GPU parity tests: SpMV, dot/norm, Jacobi, PCG — bit-exact against the oracle."""
import numpy as np
import pytest

import oracle_ffi as o

pytestmark = pytest.mark.gpu

CASES = [("poisson2d", 16), ("poisson2d", 67), ("convdiff2d", 48), ("varcoef27", 12), ("poisson3d", 24), ("convdiff3d", 17)]


def _mk(kind, N, ctx):
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil(kind, N)
    return kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx), o.OCsr(n, n, rp, ci, v)


@pytest.mark.parametrize("kind,N", CASES)
def test_spmv_bit_exact(ctx, kind, N):
    A, Ao = _mk(kind, N, ctx)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(Ao.n)
    y = np.zeros(Ao.n)
    A.matvec(x, y)
    assert np.array_equal(y, o.spmv(Ao, x))


@pytest.mark.parametrize("n", [1, 2, 511, 512, 513, 1000, 131072, 131073, 1 << 20])
def test_dot_norm_bit_exact(ctx, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    y = rng.standard_normal(n)
    assert ctx.dot(x, y) == o.dot(x, y)
    assert ctx.norm(x) == o.norm(x)
    assert abs(ctx.dot(x, y) - float(np.dot(x, y))) <= 1e-12 * max(1.0, float(np.abs(x * y).sum()))


@pytest.mark.parametrize("kind,N", CASES)
def test_jacobi_bit_exact(ctx, kind, N):
    import kryst_b200 as kb
    A, Ao = _mk(kind, N, ctx)
    pc = kb.Jacobi().setup(A)
    assert np.array_equal(pc.inv_diag, o.jacobi_inv_diag(Ao))
    r = np.random.default_rng(3).standard_normal(Ao.n)
    z = np.zeros(Ao.n)
    pc.apply(r, z)
    assert np.array_equal(z, o.OPc.jacobi(Ao).apply(r))


@pytest.mark.parametrize("kind,N", [("poisson2d", 16), ("poisson2d", 100), ("poisson3d", 24), ("varcoef27", 12)])
@pytest.mark.parametrize("use_pc", [True, False])
@pytest.mark.parametrize("persistent", ["0", "1"])
def test_pcg_bit_exact(ctx, kind, N, use_pc, persistent, monkeypatch):
    """Both drivers: CUDA-graph replay of 3 kernels/iteration, and the single persistent cooperative kernel."""
    import kryst_b200 as kb
    monkeypatch.setenv("KB_PCG_PERSISTENT", persistent)
    A, Ao = _mk(kind, N, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    pc = kb.Jacobi().setup(A) if use_pc else None
    x = np.zeros(Ao.n)
    s = kb.PcgSolver(1e-8, 5000)
    st = s.solve(A, pc, b, x)
    rc, xo, so, ho = o.pcg(Ao, o.OPc.jacobi(Ao) if use_pc else None, b, np.zeros(Ao.n), 1e-8, 5000, hist_cap=5001)
    assert rc == 0
    assert st.iterations == so.iterations and st.converged == bool(so.converged)
    assert st.final_residual == so.final_residual
    assert np.array_equal(x, xo)
    assert np.array_equal(np.array(s.residual_history), ho)


def test_handle_lifetimes_any_destroy_order(built):
    """pc -> operator -> context references: destroying parents first must be safe."""
    import kryst_b200 as kb
    from kryst_b200 import stencils
    c = kb.Context(0)
    n, rp, ci, v = stencils.stencil("poisson2d", 10)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, c)
    pc = kb.Jacobi().setup(A)
    ilu = kb.Ilu0().setup(A)
    c.close()
    A.close()
    r = np.ones(n)
    z = np.zeros(n)
    pc.apply(r, z)            # still valid: the pc keeps its operator and context alive
    assert np.all(z == 0.25)
    ilu.apply(r, z)
    pc.close()
    ilu.close()
