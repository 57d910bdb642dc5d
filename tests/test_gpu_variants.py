"""GPU edge cases and kernel variants: every SpMV kernel kind, ragged / empty / long rows, rectangular operators,
invalid CSR rejection, PCG norm types / error paths / history, generic-preconditioner PCG, device-tensor API."""
import os

import numpy as np
import pytest

import oracle_ffi as o

pytestmark = pytest.mark.gpu


def _rand_csr(n, m, row_len, seed, dense_rows=()):
    rng = np.random.default_rng(seed)
    rp, ci, v = [0], [], []
    for i in range(n):
        k = int(row_len(i, rng))
        if i in dense_rows:
            k = m
        cols = np.sort(rng.choice(m, size=min(k, m), replace=False)) if k > 0 else np.zeros(0, dtype=np.int64)
        ci.extend(cols.tolist())
        v.extend(rng.standard_normal(cols.size).tolist())
        rp.append(len(ci))
    return np.array(rp, dtype=np.uint64), np.array(ci, dtype=np.uint64), np.array(v)


def _mk(kind, N, ctx):
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil(kind, N)
    return kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx), o.OCsr(n, n, rp, ci, v)


def _check_spmv(ctx, n, m, rp, ci, v, exact=True, kind=None):
    import kryst_b200 as kb
    A = kb.DeviceCsr.from_csr(n, m, rp, ci, v, ctx)
    if kind is not None:
        assert A.spmv_kernel_kind() == kind
    x = np.random.default_rng(1).standard_normal(m)
    y = np.zeros(n)
    A.matvec(x, y)
    yo = o.spmv(o.OCsr(n, m, rp, ci, v), x)
    if exact:
        assert np.array_equal(y, yo)
    else:
        assert np.allclose(y, yo, rtol=1e-12, atol=1e-12 * np.abs(yo).max())
    return A


def test_reference_csr_fixtures(ctx):
    """src/matrix/sparse.rs:121-144 and tests/core_dense.rs:38-47 through the CUDA path."""
    import kryst_b200 as kb
    A = kb.DeviceCsr.from_csr(3, 3, [0, 1, 2, 3], [0, 1, 2], [1.0, 1.0, 1.0], ctx)
    y = np.zeros(3)
    A.matvec(np.array([2.0, 3.0, 5.0]), y)
    assert y.tolist() == [2.0, 3.0, 5.0]
    B = kb.DeviceCsr.from_csr(2, 3, [0, 2, 4], [0, 1, 1, 2], [1.0, 2.0, 3.0, 4.0], ctx)
    assert (B.nrows(), B.ncols()) == (2, 3)
    y = np.zeros(2)
    B.matvec(np.ones(3), y)
    assert y.tolist() == [3.0, 7.0]
    assert abs(ctx.dot([1, 2, 3], [4, -5, 6]) - 12.0) < 1e-12
    assert abs(ctx.norm([1, 2, 3]) - np.sqrt(14.0)) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 511, 512, 513, 1500])
def test_ragged_rows_with_empty_rows_bulk_kernel(ctx, n):
    rp, ci, v = _rand_csr(n, n + 7, lambda i, r: r.integers(0, 12) if i % 5 else 0, seed=n)
    _check_spmv(ctx, n, n + 7, rp, ci, v, exact=True, kind=2)


def test_plain_load_stream_kernel_forced(ctx):
    os.environ["KB_SPMV_KIND"] = "0"
    try:
        rp, ci, v = _rand_csr(3000, 3000, lambda i, r: r.integers(1, 40), seed=3)
        _check_spmv(ctx, 3000, 3000, rp, ci, v, exact=True, kind=0)
    finally:
        del os.environ["KB_SPMV_KIND"]


def test_rows_longer_than_a_stage_fall_back_and_stay_exact(ctx):
    n, m = 700, 9000
    rp, ci, v = _rand_csr(n, m, lambda i, r: r.integers(0, 9), seed=5, dense_rows=(3, 600))
    _check_spmv(ctx, n, m, rp, ci, v, exact=True, kind=0)      # 9000 > 4096 nnz per stage


def test_vector_per_row_kernel_for_long_rows(ctx):
    n = 1200
    rp, ci, v = _rand_csr(n, n, lambda i, r: r.integers(90, 140), seed=7)
    _check_spmv(ctx, n, n, rp, ci, v, exact=False, kind=1)     # different summation tree: 1e-12 instead of bit-exact


def test_product_phase_variant_27pt(ctx, monkeypatch):
    from kryst_b200 import stencils
    monkeypatch.setenv("KB_SPMV_XTILE", "0")          # kb_spmv_bulk's own product phase (the default for long rows stages x: test_gpu_xtile.py)
    n, rp, ci, v = stencils.stencil("varcoef27", 14)
    _check_spmv(ctx, n, n, rp, ci, v, exact=True, kind=2)


def test_invalid_csr_is_rejected(ctx):
    import kryst_b200 as kb
    with pytest.raises(kb.KError):
        kb.DeviceCsr.from_csr(2, 2, [0, 2, 3], [1, 0, 1], [1.0, 2.0, 3.0], ctx)      # columns not ascending
    with pytest.raises(kb.KError):
        kb.DeviceCsr.from_csr(2, 2, [0, 1, 2], [0, 5], [1.0, 2.0], ctx)              # column out of range
    with pytest.raises(kb.KError):
        kb.DeviceCsr.from_csr(3, 3, [0, 2, 1, 3], [0, 1, 2], [1.0, 2.0, 3.0], ctx)   # row_ptr not monotone
    with pytest.raises(kb.KError):
        kb.DeviceCsr.from_csr(2, 2, [0, 1, 1], [0, 0], [1.0, 1.0], ctx)              # inconsistent lengths


def test_empty_operator(ctx):
    import kryst_b200 as kb
    A = kb.DeviceCsr.from_csr(0, 0, [0], [], [], ctx)
    assert A.nrows() == 0
    st = kb.PcgSolver(1e-8, 10).solve(A, None, np.zeros(0), np.zeros(0))
    assert st.iterations == 0


@pytest.mark.parametrize("norm_type", [0, 1, 2, 3])
def test_pcg_norm_types_bit_exact(ctx, norm_type):
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil("poisson2d", 24)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    Ao = o.OCsr(n, n, rp, ci, v)
    b = o.spmv(Ao, np.ones(n))
    x = np.zeros(n)
    s = kb.PcgSolver(1e-7, 60).with_norm(norm_type)
    st = s.solve(A, kb.Jacobi().setup(A), b, x)
    rc, xo, so, ho = o.pcg(Ao, o.OPc.jacobi(Ao), b, np.zeros(n), 1e-7, 60, norm_type=norm_type, hist_cap=100)
    assert (st.iterations, st.converged) == (so.iterations, bool(so.converged))
    assert st.final_residual == so.final_residual and np.array_equal(x, xo)
    assert np.array_equal(np.array(s.residual_history), ho)


def test_pcg_natural_norm_first_entry_is_nan_when_rz_negative(ctx):
    """pcg.rs:137-146 pushes dp.sqrt() of the raw r.z before the loop (no abs): with a negative Jacobi diagonal the first
    history entry is NaN, res0 stays |r.z|.sqrt() (:134).  max_iters = 0: the loop never runs, Ok is returned."""
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil("poisson2d", 12)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, -v, ctx)            # negative definite: D^-1 < 0, r.z < 0
    Ao = o.OCsr(n, n, rp, ci, -v)
    b = o.spmv(Ao, np.ones(n))
    for nt in (0, 1, 2, 3):
        s = kb.PcgSolver(1e-8, 0).with_norm(nt)
        x = np.zeros(n)
        st = s.solve(A, kb.Jacobi().setup(A), b, x)
        rc, xo, so, ho = o.pcg(Ao, o.OPc.jacobi(Ao), b, np.zeros(n), 1e-8, 0, norm_type=nt, hist_cap=4)
        assert rc == 0 and (st.iterations, st.final_residual) == (0, so.final_residual)
        h = np.array(s.residual_history)
        assert h.shape == ho.shape == (1,) and np.array_equal(h, ho, equal_nan=True)
        assert np.isnan(h[0]) == (nt == 2)


def test_pcg_max_iters_reports_converged_and_history_capacity(ctx):
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil("poisson2d", 30)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    b = np.ones(n)
    x = np.zeros(n)
    s = kb.PcgSolver(1e-14, 5)
    s.history_capacity = 3
    st = s.solve(A, None, b, x)
    assert st.iterations == 5 and st.converged          # convergence.rs:24-25 (F8)
    assert len(s.residual_history) == 3


def test_pcg_indefinite_matrix_leaves_x_untouched(ctx):
    import kryst_b200 as kb
    A = kb.DeviceCsr.from_csr(3, 3, [0, 1, 2, 3], [0, 1, 2], [1.0, -1.0, 2.0], ctx)
    x = np.array([0.5, 0.25, 0.125])
    with pytest.raises(kb.IndefiniteMatrix):
        kb.PcgSolver(1e-10, 50).solve(A, None, np.ones(3), x)
    assert x.tolist() == [0.5, 0.25, 0.125]             # pcg.rs:171 returns Err before writing x


def test_pcg_error_class_matches_oracle_on_indefinite_inputs(ctx):
    """pAp <= 0 -> IndefiniteMatrix (pcg.rs:162-172); beta < 0 -> IndefinitePreconditioner (pcg.rs:206-213)."""
    import kryst_b200 as kb
    seen = set()
    for diag, b in (([2.0, -3.0, 5.0], [1.0, 0.5, 1.0]), ([2.0, -3.0, 5.0], [0.1, 2.0, 0.1]), ([1.0, -1.0, 2.0], [1.0, 1.0, 1.0]),
                    ([4.0, -0.5, 1.0], [1.0, 0.3, 1.0])):
        A = kb.DeviceCsr.from_csr(3, 3, [0, 1, 2, 3], [0, 1, 2], diag, ctx)
        Ao = o.OCsr(3, 3, [0, 1, 2, 3], [0, 1, 2], diag)
        rc, xo, so, _ = o.pcg(Ao, o.OPc.jacobi(Ao), b, np.zeros(3), 1e-12, 50)
        x = np.zeros(3)
        try:
            st = kb.PcgSolver(1e-12, 50).solve(A, kb.Jacobi().setup(A), np.array(b), x)
            got = 0
            assert st.iterations == so.iterations and np.array_equal(x, xo)
        except kb.IndefiniteMatrix:
            got = 3
        except kb.IndefinitePreconditioner:
            got = 4
        assert got == rc, (diag, b, got, rc)
        if got:
            assert x.tolist() == [0.0, 0.0, 0.0]
        seen.add(got)
    assert 3 in seen or 4 in seen


def test_pc_set_up_for_another_operator_is_rejected(ctx):
    import kryst_b200 as kb
    A = kb.DeviceCsr.from_csr(3, 3, [0, 1, 2, 3], [0, 1, 2], [2.0, 3.0, 5.0], ctx)
    B = kb.DeviceCsr.from_csr(3, 3, [0, 1, 2, 3], [0, 1, 2], [1.0, 1.0, 1.0], ctx)
    pc = kb.Jacobi().setup(B)
    with pytest.raises(kb.KError):
        kb.PcgSolver(1e-12, 50).solve(A, pc, np.array([1.0, 2.0, 3.0]), np.zeros(3))


@pytest.mark.parametrize("kind,N", [("poisson2d", 24), ("poisson3d", 10)])
def test_pcg_with_ilu0_preconditioner_bit_exact(ctx, kind, N):
    """PcgSolver accepts any Preconditioner (pcg.rs:126-131); ILU(0) goes through the unfused update/apply/dots path."""
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil(kind, N)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    Ao = o.OCsr(n, n, rp, ci, v)
    b = o.spmv(Ao, np.ones(n))
    x = np.zeros(n)
    st = kb.PcgSolver(1e-9, 500).solve(A, kb.Ilu0().setup(A), b, x)
    rc, xo, so, _ = o.pcg(Ao, o.OPc.ilu0(Ao), b, np.zeros(n), 1e-9, 500)
    assert rc == 0 and (st.iterations, st.converged) == (so.iterations, bool(so.converged))
    assert st.final_residual == so.final_residual and np.array_equal(x, xo)


def test_device_tensor_api_matches_host_api(ctx):
    import torch
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil("poisson3d", 12)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    pc = kb.Jacobi().setup(A)
    b = np.arange(1.0, n + 1.0)
    xh = np.zeros(n)
    sh = kb.PcgSolver(1e-9, 400).solve(A, pc, b, xh)
    bd = torch.tensor(b, dtype=torch.float64, device="cuda")
    xd = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    sd = kb.PcgSolver(1e-9, 400).solve(A, pc, bd, xd)
    assert sd.iterations == sh.iterations and np.array_equal(xd.cpu().numpy(), xh)
    yd = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    A.matvec(bd, yd)
    yh = np.zeros(n)
    A.matvec(b, yh)
    assert np.array_equal(yd.cpu().numpy(), yh)


def test_graph_and_plain_launch_paths_agree(ctx):
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil("convdiff2d", 20)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    b = np.ones(n)
    res = []
    for flags in (0, kb.api.KB_FLAG_NO_GRAPH, kb.api.KB_FLAG_PROFILE):
        x = np.zeros(n)
        s = kb.GmresSolver(8, 1e-9, 500)
        s.flags = flags
        st = s.solve(A, kb.Ilu0().setup(A), b, x)
        res.append((st.iterations, st.final_residual, x.copy()))
    for r in res[1:]:
        assert r[0] == res[0][0] and r[1] == res[0][1] and np.array_equal(r[2], res[0][2])
    assert ctx.launch_count() > 0


def test_ksp_context_dispatch(ctx):
    """KspContext::solve_context (src/context/ksp_context.rs:88-148) routes to the same device solvers."""
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil("convdiff2d", 20)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    Ao = o.OCsr(n, n, rp, ci, v)
    b = o.spmv(Ao, np.ones(n))
    x = np.zeros(n)
    st = kb.KspContext(kb.SolverKind.GmresLeft, A, kb.Ilu0().setup(A), tol=1e-9, max_it=500, restart=12).solve_context(b, x)
    rc, xo, so = o.gmres(Ao, o.OPc.ilu0(Ao), b, np.zeros(n), 12, 1e-9, 500, mode=1, variant=o.GMRES_CGS2)
    assert st.iterations == so.iterations and np.array_equal(x, xo)
    x = np.zeros(n)
    st = kb.KspContext(kb.SolverKind.GmresRight, A, kb.Jacobi().setup(A), tol=1e-9, max_it=500, restart=12).solve_context(b, x)
    rc, xo, so = o.gmres(Ao, o.OPc.jacobi(Ao), b, np.zeros(n), 12, 1e-9, 500, mode=2, variant=o.GMRES_CGS2)
    assert st.iterations == so.iterations and np.array_equal(x, xo)
    x = np.zeros(n)
    tol_abs = 1e-8 * float(np.linalg.norm(b))
    st = kb.KspContext(kb.SolverKind.Bicgstab, A, kb.Jacobi().setup(A), tol=tol_abs, max_it=500).solve_context(b, x)
    rc, xo, so = o.bicgstab(Ao, None, b, np.zeros(n), tol_abs, 500)
    assert st.iterations == so.iterations and np.array_equal(x, xo)
    ns, rps, cis, vs = stencils.stencil("poisson2d", 20)
    S = kb.DeviceCsr.from_csr(ns, ns, rps, cis, vs, ctx)
    So = o.OCsr(ns, ns, rps, cis, vs)
    bs = o.spmv(So, np.ones(ns))
    for kind, pco in ((kb.SolverKind.Pcg, o.OPc.jacobi(So)), (kb.SolverKind.Cg, None)):
        x = np.zeros(ns)
        st = kb.KspContext(kind, S, kb.Jacobi().setup(S), tol=1e-9, max_it=500).solve_context(bs, x)
        rc, xo, so, _ = o.pcg(So, pco, bs, np.zeros(ns), 1e-9, 500)
        assert st.iterations == so.iterations and np.array_equal(x, xo)
    with pytest.raises(kb.Unsupported):
        kb.KspContext(kb.SolverKind.Minres, S).solve_context(bs, np.zeros(ns))


@pytest.mark.parametrize("order", ["jacobi_then_ilu", "ilu_then_jacobi"])
@pytest.mark.parametrize("solver", ["pcg", "gmres", "bicgstab"])
def test_graph_cache_survives_preconditioner_replacement(ctx, order, solver):
    """the solvers' CUDA-graph caches live on the operator; they must never be replayed for a new
    preconditioner that happens to be allocated where a destroyed one lived.  Solve with one preconditioner, destroy
    it, solve with the other kind on the same operator: every solve is bit-identical to the oracle."""
    import kryst_b200 as kb
    A, Ao = _mk("poisson3d", 14, ctx)
    b = o.spmv(Ao, np.ones(Ao.n))
    seq = ["jacobi", "ilu0", "jacobi", "ilu0"] if order == "jacobi_then_ilu" else ["ilu0", "jacobi", "ilu0", "jacobi"]
    for kind in seq:
        pc = (kb.Jacobi() if kind == "jacobi" else kb.Ilu0()).setup(A)
        pco = o.OPc.jacobi(Ao) if kind == "jacobi" else o.OPc.ilu0(Ao)
        x = np.zeros(Ao.n)
        if solver == "pcg":
            st = kb.PcgSolver(1e-8, 2000).solve(A, pc, b, x)
            rc, xo, so, _ = o.pcg(Ao, pco, b, np.zeros(Ao.n), 1e-8, 2000)
        elif solver == "gmres":
            st = kb.GmresSolver(10, 1e-8, 2000).solve(A, pc, b, x)
            rc, xo, so = o.gmres(Ao, pco, b, np.zeros(Ao.n), 10, 1e-8, 2000, mode=o.MODE_LEFT, variant=o.GMRES_CGS2)
        else:
            st = kb.BiCgStabSolver(1e-8, 2000, textbook=True).solve(A, pc, b, x)
            rc, xo, so = o.bicgstab(Ao, pco, b, np.zeros(Ao.n), 1e-8, 2000, variant=o.BICG_TEXTBOOK)
        assert rc == 0 and st.iterations == so.iterations and st.final_residual == so.final_residual, (kind, st.iterations, so.iterations)
        assert np.array_equal(x, xo), kind
        pc.close()      # the next preconditioner may reuse this handle's address
