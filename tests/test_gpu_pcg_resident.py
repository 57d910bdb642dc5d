"""kb_pcg_resident (whole PCG solve on chip, one cooperative launch): bit-exact against the oracle - iterations, residual
history, x - for every norm type, both preconditioner choices, limits and error exits; problems that do not fit, rows
with more than 8 entries or more than 4 ghost columns per thread must fall back to the CUDA-graph path unchanged."""
import numpy as np
import pytest

import oracle_ffi as o

pytestmark = pytest.mark.gpu


def _mk(kind, N, ctx):
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil(kind, N)
    return kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx), o.OCsr(n, n, rp, ci, v)


def _solve(ctx, A, pc, b, x, solver):
    """returns (stats, kernel launches of this solve)"""
    l0 = ctx.launch_count()
    st = solver.solve(A, pc, b, x)
    return st, ctx.launch_count() - l0


@pytest.mark.parametrize("kind,N", [("poisson2d", 16), ("poisson2d", 130), ("poisson2d", 301), ("poisson3d", 24), ("poisson3d", 41)])
@pytest.mark.parametrize("use_pc", [True, False])
@pytest.mark.parametrize("norm", [0, 1, 2, 3])
def test_resident_bit_exact(ctx, kind, N, use_pc, norm, monkeypatch):
    import kryst_b200 as kb
    monkeypatch.setenv("KB_PCG_RESIDENT", "1")
    A, Ao = _mk(kind, N, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    pc = kb.Jacobi().setup(A) if use_pc else None
    x = np.zeros(Ao.n)
    s = kb.PcgSolver(1e-8, 5000).with_norm(norm)
    st, launches = _solve(ctx, A, pc, b, x, s)
    assert launches <= 4                      # r = b - A x, init pass, the resident kernel
    rc, xo, so, ho = o.pcg(Ao, o.OPc.jacobi(Ao) if use_pc else None, b, np.zeros(Ao.n), 1e-8, 5000, norm_type=norm, hist_cap=5001)
    assert rc == 0
    assert st.iterations == so.iterations and st.converged == bool(so.converged)
    assert st.final_residual == so.final_residual
    assert np.array_equal(x, xo)
    assert np.array_equal(np.array(s.residual_history), ho)


def test_resident_matches_graph_path_and_repeats(ctx, monkeypatch):
    """Same workspace, alternating drivers and right-hand sides: packets and marks of an earlier solve must not leak."""
    import kryst_b200 as kb
    A, Ao = _mk("poisson2d", 200, ctx)
    pc = kb.Jacobi().setup(A)
    rng = np.random.default_rng(5)
    for rep in range(3):
        b = o.spmv(Ao, rng.standard_normal(Ao.n))
        out = []
        for res in ("1", "0", "1"):
            monkeypatch.setenv("KB_PCG_RESIDENT", res)
            x = np.zeros(Ao.n)
            st, launches = _solve(ctx, A, pc, b, x, kb.PcgSolver(1e-10, 3000))
            assert (launches <= 4) == (res == "1")
            out.append((st.iterations, st.final_residual, st.converged, x))
        for k in (1, 2):
            assert out[0][:3] == out[k][:3] and np.array_equal(out[0][3], out[k][3])
        rc, xo, so, _ = o.pcg(Ao, o.OPc.jacobi(Ao), b, np.zeros(Ao.n), 1e-10, 3000)
        assert out[0][0] == so.iterations and np.array_equal(out[0][3], xo)


def test_resident_limits_and_error_exits(ctx, monkeypatch):
    import kryst_b200 as kb
    monkeypatch.setenv("KB_PCG_RESIDENT", "1")
    A, Ao = _mk("poisson2d", 40, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    for tol, mi, x0 in ((1e-30, 7, 0.0), (1e-8, 0, 0.25), (1e3, 50, 0.0), (1e-8, 1, 0.5)):
        x = np.full(Ao.n, x0)
        st = kb.PcgSolver(tol, mi).solve(A, None, b, x)
        rc, xo, so, _ = o.pcg(Ao, None, b, np.full(Ao.n, x0), tol, mi)
        assert (st.iterations, st.converged, st.final_residual) == (so.iterations, bool(so.converged), so.final_residual), (tol, mi)
        assert np.array_equal(x, xo)
    n, rp, ci, v = Ao.n, Ao.row_ptr, Ao.col_idx, -Ao.vals          # negative definite: p.Ap <= 0 in the first iteration
    An = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    x = np.full(n, 3.0)
    with pytest.raises(kb.IndefiniteMatrix):
        kb.PcgSolver(1e-8, 50).solve(An, None, b, x)
    assert np.array_equal(x, np.full(n, 3.0))                       # x untouched on Err (pcg.rs:171)


@pytest.mark.parametrize("n", [1, 2, 5, 511, 513, 1025])
def test_resident_tiny_and_odd_sizes(ctx, n, monkeypatch):
    import kryst_b200 as kb
    monkeypatch.setenv("KB_PCG_RESIDENT", "1")
    rp, ci, v = [0], [], []
    for i in range(n):                                              # 1-D Laplacian, diagonally dominant
        for j, a in ((i - 1, -1.0), (i, 2.5), (i + 1, -1.0)):
            if 0 <= j < n:
                ci.append(j); v.append(a)
        rp.append(len(ci))
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    Ao = o.OCsr(n, n, rp, ci, v)
    b = o.spmv(Ao, np.arange(1.0, n + 1.0))
    x = np.zeros(n)
    st, launches = _solve(ctx, A, kb.Jacobi().setup(A), b, x, kb.PcgSolver(1e-12, 500))
    assert launches <= 4
    rc, xo, so, _ = o.pcg(Ao, o.OPc.jacobi(Ao), b, np.zeros(n), 1e-12, 500)
    assert (st.iterations, st.final_residual) == (so.iterations, so.final_residual) and np.array_equal(x, xo)


def _banded_spd(n, offs, diag):
    """symmetric, strictly diagonally dominant: `diag` on the diagonal, -1 at the given +- offsets"""
    rows, cols, vals = [], [], []
    i = np.arange(n)
    for d in sorted(set([0] + [o_ for x in offs for o_ in (x, -x)])):
        j = i + d
        ok = (j >= 0) & (j < n)
        rows.append(i[ok]); cols.append(j[ok]); vals.append(np.full(int(ok.sum()), diag if d == 0 else -1.0))
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    rp = np.zeros(n + 1, dtype=np.uint64)
    np.add.at(rp, rows + 1, 1)
    return np.cumsum(rp).astype(np.uint64), cols.astype(np.uint64), vals


def test_resident_far_couplings_are_ghost_slots(ctx, monkeypatch):
    """Every off-diagonal entry leaves the CTA (one 512-row tile per CTA, couplings at distance >= 1000): 3072 ghost
    references per CTA, all served from the owners' packets."""
    import kryst_b200 as kb
    monkeypatch.setenv("KB_PCG_RESIDENT", "1")
    n = 70000
    rp, ci, v = _banded_spd(n, (1000, 2000, 3000), 7.0)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    Ao = o.OCsr(n, n, rp, ci, v)
    b = o.spmv(Ao, np.ones(n))
    x = np.zeros(n)
    st, launches = _solve(ctx, A, kb.Jacobi().setup(A), b, x, kb.PcgSolver(1e-8, 200))
    rc, xo, so, _ = o.pcg(Ao, o.OPc.jacobi(Ao), b, np.zeros(n), 1e-8, 200)
    assert launches <= 4 and st.iterations == so.iterations and st.final_residual == so.final_residual and np.array_equal(x, xo)


def test_resident_declines_what_does_not_fit(ctx, monkeypatch):
    """(a) more tiles than 4 per SM, (b) rows longer than 8, (c) more ghost references than the shared memory left beside
    the operator rows can hold: the CUDA-graph path must run instead (many launches), with the same bits as the oracle."""
    import kryst_b200 as kb
    monkeypatch.setenv("KB_PCG_RESIDENT", "1")
    # (a) 600^2 = 360 000 rows = 704 tiles > 592
    A, Ao = _mk("poisson2d", 600, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    x = np.zeros(Ao.n)
    st, launches = _solve(ctx, A, kb.Jacobi().setup(A), b, x, kb.PcgSolver(1e-8, 40))
    rc, xo, so, _ = o.pcg(Ao, o.OPc.jacobi(Ao), b, np.zeros(Ao.n), 1e-8, 40)
    assert launches > 40 and st.iterations == so.iterations and np.array_equal(x, xo)
    # (b) 27-point rows
    A, Ao = _mk("varcoef27", 10, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    x = np.zeros(Ao.n)
    st, launches = _solve(ctx, A, None, b, x, kb.PcgSolver(1e-8, 30))
    rc, xo, so, _ = o.pcg(Ao, None, b, np.zeros(Ao.n), 1e-8, 30)
    assert launches > 30 and st.iterations == so.iterations and np.array_equal(x, xo)
    # (c) 592 tiles (4 per CTA), 7 entries per row, every coupling at distance >= 5000: 12 288 ghost references per CTA
    # against room for ~3 600 beside 172 KB of operator rows
    n = 592 * 512
    rp, ci, v = _banded_spd(n, (5000, 10000, 15000), 7.0)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    Ao = o.OCsr(n, n, rp, ci, v)
    b = o.spmv(Ao, np.ones(n))
    x = np.zeros(n)
    st, launches = _solve(ctx, A, kb.Jacobi().setup(A), b, x, kb.PcgSolver(1e-8, 200))
    rc, xo, so, _ = o.pcg(Ao, o.OPc.jacobi(Ao), b, np.zeros(n), 1e-8, 200)
    assert launches > 20 and st.iterations == so.iterations and st.final_residual == so.final_residual and np.array_equal(x, xo)
