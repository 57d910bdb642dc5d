"""kb_spmv_xtile (x tiles staged in shared memory, 16-bit chunk-local column ids): bit-exact against the oracle for the
plain product, the fused-dot SpMVs inside PCG / BiCGStab / GMRES, odd and rectangular shapes (tail slot), operators
that do not fit (must fall back to kb_spmv_bulk), and unaligned device operands (must fall back per call)."""
import numpy as np
import pytest

import oracle_ffi as o

pytestmark = pytest.mark.gpu


def _env(monkeypatch, cfg, mode="1"):
    monkeypatch.setenv("KB_SPMV_XTILE", mode)
    monkeypatch.setenv("KB_XT_CFG", str(cfg))


def _mk(kind, N, ctx):
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil(kind, N)
    return kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx), o.OCsr(n, n, rp, ci, v)


def _banded(n, m, seed, maxlen=12, last_col=True):
    rng = np.random.default_rng(seed)
    rp, ci, v = [0], [], []
    for i in range(n):
        k = int(rng.integers(0, maxlen)) if i % 5 else 0
        centre = min(m - 1, i)
        cand = np.unique(np.clip(centre + rng.integers(-6, 7, size=k), 0, m - 1))
        if last_col and i == n - 1:
            cand = np.unique(np.append(cand, m - 1))
        ci.extend(cand.tolist()); v.extend(rng.standard_normal(cand.size).tolist()); rp.append(len(ci))
    return np.array(rp, dtype=np.uint64), np.array(ci, dtype=np.uint64), np.array(v)


@pytest.mark.parametrize("cfg", [0, 1, 2])
@pytest.mark.parametrize("kind,N", [("varcoef27", 14), ("varcoef27", 31), ("poisson3d", 17), ("convdiff3d", 24), ("poisson2d", 33),
                                    ("convdiff2d", 130)])
def test_xtile_spmv_stencils_bit_exact(ctx, kind, N, cfg, monkeypatch):
    _env(monkeypatch, cfg)
    A, Ao = _mk(kind, N, ctx)
    assert A.spmv_kernel_kind() == 2 and A.spmv_x_staged() == cfg + 1
    rng = np.random.default_rng(N)
    for _ in range(3):                      # repeated launches: ring phases start fresh in every launch
        x = rng.standard_normal(Ao.n)
        y = np.zeros(Ao.n)
        A.matvec(x, y)
        assert np.array_equal(y, o.spmv(Ao, x))


@pytest.mark.parametrize("cfg", [0, 1, 2])
@pytest.mark.parametrize("n,m", [(1, 1), (5, 5), (513, 513), (700, 707), (1500, 1501), (4099, 4099)])
def test_xtile_ragged_odd_rectangular(ctx, n, m, cfg, monkeypatch):
    import kryst_b200 as kb
    _env(monkeypatch, cfg)
    rp, ci, v = _banded(n, m, seed=n)
    A = kb.DeviceCsr.from_csr(n, m, rp, ci, v, ctx)
    assert A.spmv_x_staged() == cfg + 1
    x = np.random.default_rng(2).standard_normal(m)
    y = np.zeros(n)
    A.matvec(x, y)
    assert np.array_equal(y, o.spmv(o.OCsr(n, m, rp, ci, v), x))


def test_xtile_declines_scattered_columns(ctx, monkeypatch):
    import kryst_b200 as kb
    _env(monkeypatch, 0)
    rng = np.random.default_rng(3)
    n, m = 600, 200000
    rp, ci, v = [0], [], []
    for i in range(n):
        cols = np.sort(rng.choice(m, size=9, replace=False))
        ci.extend(cols.tolist()); v.extend(rng.standard_normal(9).tolist()); rp.append(len(ci))
    A = kb.DeviceCsr.from_csr(n, m, rp, ci, v, ctx)
    assert A.spmv_kernel_kind() == 2 and A.spmv_x_staged() == 0
    x = rng.standard_normal(m)
    y = np.zeros(n)
    A.matvec(x, y)
    assert np.array_equal(y, o.spmv(o.OCsr(n, m, rp, ci, v), x))


def test_xtile_auto_mode_only_long_rows(ctx, monkeypatch):
    _env(monkeypatch, 1, mode="2")
    A7, _ = _mk("poisson3d", 12, ctx)
    A27, _ = _mk("varcoef27", 12, ctx)
    assert A7.spmv_x_staged() == 0 and A27.spmv_x_staged() == 2


def test_xtile_is_the_default_for_long_rows_only(ctx, monkeypatch):
    monkeypatch.delenv("KB_SPMV_XTILE", raising=False)
    monkeypatch.delenv("KB_XT_CFG", raising=False)
    A7, _ = _mk("poisson3d", 12, ctx)
    A27, Ao = _mk("varcoef27", 12, ctx)
    assert A7.spmv_x_staged() == 0 and A27.spmv_x_staged() == 2       # geometry 1 first
    x = np.random.default_rng(1).standard_normal(Ao.n)
    y = np.zeros(Ao.n)
    A27.matvec(x, y)
    assert np.array_equal(y, o.spmv(Ao, x))


@pytest.mark.parametrize("prod", ["0", "1"])
def test_xtile_row_wise_and_product_phase(ctx, prod, monkeypatch):
    """Both consumer forms of the staged-x kernel (thread per row / per-nonzero product phase) on short and long rows."""
    _env(monkeypatch, 0)
    monkeypatch.setenv("KB_SPMV_PROD", prod)
    for kind, N in (("varcoef27", 18), ("poisson3d", 21)):
        A, Ao = _mk(kind, N, ctx)
        assert A.spmv_x_staged() == 1
        x = np.random.default_rng(N).standard_normal(Ao.n)
        y = np.zeros(Ao.n)
        A.matvec(x, y)
        assert np.array_equal(y, o.spmv(Ao, x))


def test_xtile_second_geometry_when_the_first_is_too_small(ctx, monkeypatch):
    """A 7-point pattern 300 wide with planes of 2000 (offsets 0, +-1, +-300, +-2000): a 256-row chunk touches five
    separate column runs of 256-258 columns, 1282 > 1280 of geometry 1, so the operator takes geometry 0 (2048)."""
    import kryst_b200 as kb
    monkeypatch.setenv("KB_SPMV_XTILE", "1")
    monkeypatch.delenv("KB_XT_CFG", raising=False)
    # banded operator: offsets 0, +-1, +-300, +-2000 (the pattern of a 7-point grid 300 wide, planes of 2000)
    n = 6000
    offs = np.array([-2000, -300, -1, 0, 1, 300, 2000])
    rp, ci, v = [0], [], []
    rng = np.random.default_rng(8)
    for i in range(n):
        cols = i + offs
        cols = cols[(cols >= 0) & (cols < n)]
        ci.extend(cols.tolist()); v.extend(rng.standard_normal(cols.size).tolist()); rp.append(len(ci))
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    assert A.spmv_x_staged() == 1
    x = rng.standard_normal(n)
    y = np.zeros(n)
    A.matvec(x, y)
    assert np.array_equal(y, o.spmv(o.OCsr(n, n, rp, ci, v), x))


def test_xtile_unaligned_device_operand_falls_back(ctx, monkeypatch):
    import torch
    _env(monkeypatch, 1)
    A, Ao = _mk("varcoef27", 12, ctx)
    assert A.spmv_x_staged() == 2
    xh = np.random.default_rng(4).standard_normal(Ao.n)
    buf = torch.zeros(Ao.n + 1, dtype=torch.float64, device="cuda:0")
    buf[1:] = torch.from_numpy(xh).to("cuda:0")
    x = buf[1:]                              # 8-byte aligned only
    assert x.data_ptr() % 16 == 8
    y = torch.zeros(Ao.n, dtype=torch.float64, device="cuda:0")
    A.matvec(x, y)
    assert np.array_equal(y.cpu().numpy(), o.spmv(Ao, xh))


@pytest.mark.parametrize("cfg", [0, 1, 2])
def test_xtile_pcg_jacobi_bit_exact(ctx, cfg, monkeypatch):
    import kryst_b200 as kb
    _env(monkeypatch, cfg)
    A, Ao = _mk("poisson3d", 24, ctx)
    assert A.spmv_x_staged() == cfg + 1
    b = o.spmv(Ao, np.ones(Ao.n))
    x = np.zeros(Ao.n)
    st = kb.PcgSolver(1e-8, 1000).solve(A, kb.Jacobi().setup(A), b, x)
    rc, xo, so, _ = o.pcg(Ao, o.OPc.jacobi(Ao), b, np.zeros(Ao.n), 1e-8, 1000)
    assert (st.iterations, st.converged) == (so.iterations, bool(so.converged))
    assert st.final_residual == so.final_residual and np.array_equal(x, xo)


@pytest.mark.parametrize("cfg", [0, 1, 2])
@pytest.mark.parametrize("kind,N", [("varcoef27", 20), ("convdiff3d", 16)])
def test_xtile_bicgstab_jacobi_bit_exact(ctx, kind, N, cfg, monkeypatch):
    import kryst_b200 as kb
    _env(monkeypatch, cfg)
    A, Ao = _mk(kind, N, ctx)
    assert A.spmv_x_staged() == cfg + 1
    b = o.spmv(Ao, np.ones(Ao.n))
    x = np.zeros(Ao.n)
    st = kb.BiCgStabSolver(1e-8, 2000, textbook=True).solve(A, kb.Jacobi().setup(A), b, x)
    rc, xo, so = o.bicgstab(Ao, o.OPc.jacobi(Ao), b, np.zeros(Ao.n), 1e-8, 2000, variant=o.BICG_TEXTBOOK)
    assert (st.iterations, st.converged, st.breakdown) == (so.iterations, bool(so.converged), so.breakdown)
    assert st.final_residual == so.final_residual and np.array_equal(x, xo)


@pytest.mark.parametrize("cfg", [0, 1, 2])
def test_xtile_matches_bulk_inside_gmres_ilu0(ctx, cfg, monkeypatch):
    """GMRES(30)+ILU(0): the staged-x SpMV must reproduce the default path bit for bit (the default path is pinned to the
    oracle in test_gpu_ilu_gmres.py)."""
    import kryst_b200 as kb
    from kryst_b200 import stencils
    n, rp, ci, v = stencils.stencil("varcoef27", 16)
    Ao = o.OCsr(n, n, rp, ci, v)
    b = o.spmv(Ao, np.ones(n))
    out = []
    for mode in ("0", "1"):
        _env(monkeypatch, cfg, mode=mode)
        A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
        assert A.spmv_x_staged() == (cfg + 1 if mode == "1" else 0)
        x = np.zeros(n)
        st = kb.GmresSolver(30, 1e-8, 500).solve(A, kb.Ilu0().setup(A), b, x)
        out.append((st.iterations, st.final_residual, st.converged, x))
    assert out[0][:3] == out[1][:3] and np.array_equal(out[0][3], out[1][3])
    assert out[1][2]
