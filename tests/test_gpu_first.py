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
@pytest.mark.parametrize("driver", ["graph", "persistent", "resident"])
def test_pcg_bit_exact(ctx, kind, N, use_pc, driver, monkeypatch):
    """All drivers: CUDA-graph replay of 3 kernels/iteration, the persistent cooperative kernel (grid barriers), and the
    on-chip resident kernel (default where the problem fits; the 27-point operator declines it: rows longer than 8)."""
    import kryst_b200 as kb
    monkeypatch.setenv("KB_PCG_PERSISTENT", "1" if driver == "persistent" else "0")
    monkeypatch.setenv("KB_PCG_RESIDENT", "1" if driver == "resident" else "0")
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


@pytest.mark.parametrize("kind,N", [("poisson2d", 16), ("poisson2d", 100), ("poisson3d", 24), ("varcoef27", 12)])
@pytest.mark.parametrize("use_pc", [True, False])
@pytest.mark.parametrize("norm", [0, 1, 2, 3])
def test_pcg_fused_single_reduction_bit_exact(ctx, kind, N, use_pc, norm):
    """Chronopoulos-Gear PCG (one fused reduction per iteration) vs the oracle's restatement."""
    import kryst_b200 as kb
    A, Ao = _mk(kind, N, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    pc = kb.Jacobi().setup(A) if use_pc else None
    x = np.zeros(Ao.n)
    s = kb.PcgSolver(1e-8, 300).with_norm(norm).with_fused_reduction(True)
    st = s.solve(A, pc, b, x)
    rc, xo, so, ho = o.pcg_sr(Ao, o.OPc.jacobi(Ao) if use_pc else None, b, np.zeros(Ao.n), 1e-8, 300, norm_type=norm, hist_cap=301)
    assert rc == 0
    assert st.iterations == so.iterations and st.converged == bool(so.converged)
    assert st.final_residual == so.final_residual
    assert np.array_equal(x, xo)
    assert np.array_equal(np.array(s.residual_history), ho)
    if norm == 1 and so.iterations < 300:
        # same Krylov method as the literal recurrences: iteration count within 2 % (here: identical or +-1)
        rc, _, sl, _ = o.pcg(Ao, o.OPc.jacobi(Ao) if use_pc else None, b, np.zeros(Ao.n), 1e-8, 300)
        assert abs(int(sl.iterations) - int(st.iterations)) <= max(1, int(0.02 * sl.iterations))


@pytest.mark.parametrize("kind,N", [("poisson2d", 16), ("poisson2d", 100), ("poisson3d", 24), ("varcoef27", 12)])
@pytest.mark.parametrize("use_pc", [True, False])
@pytest.mark.parametrize("norm", [0, 1, 2, 3])
def test_pcg_pipelined_bit_exact(ctx, kind, N, use_pc, norm):
    """pipelined (Ghysels-Vanroose) PCG vs the oracle's restatement: iterations, history and x bit for bit."""
    import kryst_b200 as kb
    A, Ao = _mk(kind, N, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    pc = kb.Jacobi().setup(A) if use_pc else None
    x = np.zeros(Ao.n)
    s = kb.PcgSolver(1e-8, 300).with_norm(norm).with_pipelined()
    st = s.solve(A, pc, b, x)
    rc, xo, so, ho = o.pcg_pipe(Ao, o.OPc.jacobi(Ao) if use_pc else None, b, np.zeros(Ao.n), 1e-8, 300, norm_type=norm, hist_cap=301)
    assert rc == 0
    assert st.iterations == so.iterations and st.converged == bool(so.converged)
    assert st.final_residual == so.final_residual
    assert np.array_equal(x, xo)
    assert np.array_equal(np.array(s.residual_history), ho)
    if norm == 1 and so.iterations < 300:
        rc, _, sl, _ = o.pcg(Ao, o.OPc.jacobi(Ao) if use_pc else None, b, np.zeros(Ao.n), 1e-8, 300)
        assert abs(int(sl.iterations) - int(st.iterations)) <= max(1, int(0.02 * sl.iterations))


def test_pcg_pipelined_edges(ctx):
    import kryst_b200 as kb
    A, Ao = _mk("poisson2d", 20, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    x = np.full(Ao.n, 0.25)
    st = kb.PcgSolver(1e-8, 0).with_pipelined().solve(A, None, b, x)
    rc, xo, so, _ = o.pcg_pipe(Ao, None, b, np.full(Ao.n, 0.25), 1e-8, 0)
    assert (st.iterations, st.converged, st.final_residual) == (0, False, so.final_residual) and np.array_equal(x, xo)
    x = np.zeros(Ao.n)
    st = kb.PcgSolver(1e-30, 7).with_pipelined().solve(A, None, b, x)
    rc, xo, so, _ = o.pcg_pipe(Ao, None, b, np.zeros(Ao.n), 1e-30, 7)
    assert (st.iterations, st.converged) == (7, True) == (so.iterations, bool(so.converged)) and np.array_equal(x, xo)
    n, rp, ci, v = Ao.n, Ao.row_ptr, Ao.col_idx, -Ao.vals
    An = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    x = np.full(n, 3.0)
    with pytest.raises(kb.IndefiniteMatrix):
        kb.PcgSolver(1e-8, 50).with_pipelined().solve(An, None, b, x)
    assert np.all(x == 3.0)
    # solving again with the literal and the single-reduction recurrences on the same operator (separate graph caches)
    x1, x2 = np.zeros(Ao.n), np.zeros(Ao.n)
    kb.PcgSolver(1e-8, 300).solve(A, None, b, x1)
    kb.PcgSolver(1e-8, 300).with_pipelined().solve(A, None, b, x2)
    assert np.abs(x1 - x2).max() < 1e-6


def test_pcg_fused_single_reduction_edges(ctx):
    import kryst_b200 as kb
    A, Ao = _mk("poisson2d", 20, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    # max_iters = 0: stats {0, res0, false}, x untouched but written back
    x = np.full(Ao.n, 0.25)
    st = kb.PcgSolver(1e-8, 0).with_fused_reduction().solve(A, None, b, x)
    rc, xo, so, _ = o.pcg_sr(Ao, None, b, np.full(Ao.n, 0.25), 1e-8, 0)
    assert (st.iterations, st.converged, st.final_residual) == (0, False, so.final_residual) and np.array_equal(x, xo)
    # iteration cap hit: Convergence::check reports converged (F8), like the literal path
    x = np.zeros(Ao.n)
    st = kb.PcgSolver(1e-30, 7).with_fused_reduction().solve(A, None, b, x)
    rc, xo, so, _ = o.pcg_sr(Ao, None, b, np.zeros(Ao.n), 1e-30, 7)
    assert (st.iterations, st.converged) == (7, True) == (so.iterations, bool(so.converged)) and np.array_equal(x, xo)
    # negative definite operator: IndefiniteMatrix, x not written
    n, rp, ci, v = Ao.n, Ao.row_ptr, Ao.col_idx, -Ao.vals
    An = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    x = np.full(n, 3.0)
    with pytest.raises(kb.IndefiniteMatrix):
        kb.PcgSolver(1e-8, 50).with_fused_reduction().solve(An, None, b, x)
    assert np.all(x == 3.0)
    rc, _, so, _ = o.pcg_sr(o.OCsr(n, n, rp, ci, v), None, b, np.full(n, 3.0), 1e-8, 50)
    assert rc == 3
    # nonsymmetric operator: the recurrence p.Ap = delta - beta*gamma/alpha goes non-positive mid-solve; same
    # iteration, residual and error as the oracle
    Ac, Aco = _mk("convdiff2d", 20, ctx)
    bc = o.spmv(Aco, np.ones(Aco.n))
    sv = kb.PcgSolver(1e-8, 3000).with_fused_reduction()
    with pytest.raises(kb.IndefiniteMatrix):
        sv.solve(Ac, kb.Jacobi().setup(Ac), bc, np.zeros(Aco.n))
    rc, _, so, _ = o.pcg_sr(Aco, o.OPc.jacobi(Aco), bc, np.zeros(Aco.n), 1e-8, 3000)
    assert rc == 3 and (sv.last_stats.iterations, sv.last_stats.final_residual) == (so.iterations, so.final_residual)
    # ILU(0) cannot be fused into the update sweep
    with pytest.raises(kb.Unsupported):
        kb.PcgSolver(1e-8, 50).with_fused_reduction().solve(A, kb.Ilu0().setup(A), b, np.zeros(n))


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
