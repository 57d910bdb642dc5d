"""CPU tests (no GPU): pin the oracle against
  (1) every known-answer test the reference holds for the hot path (tests/golden/reference_known_answers.json),
  (2) the independent numpy transliteration (tests/golden/golden_small.json, made by make_golden.py),
  (3) the survey's verified iteration anchors,
  (4) independent dense / scipy computations for the Tier-T pieces (ILU(0), triangular solves, levels).
"""
import json
import os

import numpy as np
import pytest

import oracle_ffi as o

HERE = os.path.dirname(os.path.abspath(__file__))
KA = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))
GS = json.load(open(os.path.join(HERE, "golden", "golden_small.json")))


def tridiag(n, lo, d, up):
    a = np.zeros((n, n))
    for i in range(n):
        a[i, i] = d
        if i > 0:
            a[i, i - 1] = lo
        if i + 1 < n:
            a[i, i + 1] = up
    return a


def rel_err(x, xt):
    return float(np.linalg.norm(x - xt) / np.linalg.norm(xt))


# ---- (1) reference known answers --------------------------------------------------------------------
@pytest.mark.parametrize("key", ["csr_identity_spmv", "csr_simple_pattern"])
def test_ref_csr_spmv(key):
    c = KA[key]
    A = o.OCsr(c["nrows"], c["ncols"], c["row_ptr"], c["col_idx"], c["vals"])
    assert o.lib().ko_csr_validate(A.ptr()) == 0
    assert o.spmv(A, c["x"]).tolist() == c["y"]          # assert_eq! in the reference: exact


def test_ref_dot_norm():
    c = KA["dot_norm"]
    assert abs(o.dot(c["x"], c["y"]) - c["dot"]) < c["tol"]
    assert abs(o.norm(c["x"]) - np.sqrt(c["norm_x_squared"])) < c["tol"]


def test_ref_cg_2x2():
    c = KA["cg_2x2"]
    A = o.OCsr.from_dense(c["a"])
    rc, x, st, _ = o.pcg(A, None, c["b"], [0, 0], c["tol_solver"], c["max_iters"])
    assert rc == 0 and st.converged
    assert np.allclose(x, c["x"], atol=c["tol_check"], rtol=0)


def test_ref_cg_3x3():
    c = KA["cg_3x3"]
    a = np.array(c["a"])
    A = o.OCsr.from_dense(a)
    rc, x, st, _ = o.pcg(A, None, c["b"], np.zeros(3), 1e-10, 100)
    assert rc == 0 and np.linalg.norm(a @ x - c["b"]) <= c["residual_tol"]


@pytest.mark.parametrize("key,use_pc", [("pcg_jacobi_tridiag", True), ("pcg_nopc_tridiag", False)])
def test_ref_pcg_tridiag(key, use_pc):
    c = KA[key]
    a = tridiag(c["n"], c["lower"], c["diag"], c["upper"])
    A = o.OCsr.from_dense(a)
    b = a @ np.full(c["n"], c["x_true"])
    rc, x, st, _ = o.pcg(A, o.OPc.jacobi(A) if use_pc else None, b, np.zeros(c["n"]), c["tol_solver"], c["max_iters"])
    assert rc == 0 and st.converged
    assert rel_err(x, np.full(c["n"], c["x_true"])) < c["rel_err"]
    assert st.iterations <= c["n"]


def test_ref_gmres_nonsym_tridiag():
    c = KA["gmres_nonsym_tridiag"]
    a = tridiag(c["n"], c["lower"], c["diag"], c["upper"])
    A = o.OCsr.from_dense(a)
    b = a @ np.ones(c["n"])
    for variant in (o.GMRES_LITERAL, o.GMRES_CGS2, o.GMRES_MGS2):
        rc, x, st = o.gmres(A, None, b, np.zeros(c["n"]), c["restart"], c["tol_solver"], c["max_iters"], variant=variant)
        assert st.converged and rel_err(x, np.ones(c["n"])) < c["rel_err"]


def test_ref_gmres_left_literal_ilu0():
    c = KA["gmres_left_ilu0_literal_tridiag"]
    a = tridiag(c["n"], c["lower"], c["diag"], c["upper"])
    A = o.OCsr.from_dense(a)
    b = a @ np.ones(c["n"])
    rc, x, st = o.gmres(A, o.OPc.ilu_literal(a), b, np.zeros(c["n"]), c["restart"], c["tol_solver"], c["max_iters"],
                        mode=o.MODE_LEFT, variant=o.GMRES_LITERAL)
    assert st.converged and rel_err(x, np.ones(c["n"])) < c["rel_err"]


def test_ref_gmres_4x4():
    c = KA["gmres_4x4"]
    a = np.array(c["a"])
    A = o.OCsr.from_dense(a)
    xt = np.array(c["x_true"])
    b = a @ xt
    rc, x, st = o.gmres(A, None, b, np.zeros(4), c["restart"], c["tol_solver"], c["max_iters"], variant=o.GMRES_LITERAL)
    assert st.converged and np.abs(x - xt).max() < c["tol_check"]
    rc, x, st = o.gmres(A, o.OPc.jacobi(A), b, np.zeros(4), c["restart"], c["tol_solver"], c["max_iters"], mode=o.MODE_LEFT,
                        variant=o.GMRES_LITERAL)
    assert st.converged and np.abs(x - xt).max() < c["tol_check"]
    rc, x, st = o.gmres(A, o.OPc.jacobi(A), b, np.zeros(4), c["restart"], c["tol_solver"], c["max_iters"], mode=o.MODE_RIGHT,
                        variant=o.GMRES_LITERAL)
    assert np.linalg.norm(a @ x - b) < c["right_residual_tol"]     # the reference asserts only this (gmres.rs:521-527)
    # the textbook formulations solve it outright in every mode
    for mode in (o.MODE_NONE, o.MODE_LEFT, o.MODE_RIGHT):
        rc, x, st = o.gmres(A, o.OPc.jacobi(A), b, np.zeros(4), c["restart"], c["tol_solver"], c["max_iters"], mode=mode,
                            variant=o.GMRES_CGS2)
        assert st.converged and np.abs(x - xt).max() < c["tol_check"]


def test_ref_bicgstab_3x3():
    c = KA["bicgstab_3x3"]
    a = np.array(c["a"])
    assert a.tolist() == [[4.0 if i == j else float(i + 2 * j + 1) for j in range(3)] for i in range(3)]
    A = o.OCsr.from_dense(a)
    xt = np.array(c["x_true"])
    rc, x, st = o.bicgstab(A, None, a @ xt, np.zeros(3), c["tol_solver"], c["max_iters"])
    assert st.converged and np.abs(x - xt).max() < c["tol_check"]


def test_ref_ill_cond_diag():
    c = KA["pcg_ill_cond_diag"]
    d = np.ones(c["n"])
    d[-1] = c["kappa"]
    A = o.OCsr.from_dense(np.diag(d))
    rc, x, st, _ = o.pcg(A, o.OPc.jacobi(A), np.ones(c["n"]), np.zeros(c["n"]), c["tol_solver"], c["max_iters"])
    assert rc == 0 and st.converged
    rc, x, st, _ = o.pcg(A, None, np.ones(c["n"]), np.zeros(c["n"]), c["tol_solver"], c["max_iters"])
    assert rc == 0 and st.converged


def test_ref_asm_identity_block_ilu0():
    c = KA["asm_identity"]
    A = o.OCsr.from_dense(np.eye(c["n"]))
    pc = o.OPc.ilu0(A, nblocks=c["blocks"])
    assert pc.apply(c["r"]).tolist() == c["r"]          # z == r exactly (asm.rs:124-136)


# ---- Convergence::check quirks (F8) and error paths --------------------------------------------------
def test_max_iters_reports_converged():
    a = tridiag(50, -1, 2, -1)
    A = o.OCsr.from_dense(a)
    rc, x, st, _ = o.pcg(A, None, a @ np.ones(50), np.zeros(50), 1e-14, 3)
    assert rc == 0 and st.iterations == 3 and st.converged      # convergence.rs:24-25


def test_pcg_indefinite_matrix_leaves_x():
    A = o.OCsr.from_dense(np.diag([1.0, -1.0, 2.0]))
    rc, x, st, _ = o.pcg(A, None, [1, 1, 1], np.zeros(3), 1e-10, 50)
    assert rc == 3 and not st.converged and x.tolist() == [0, 0, 0]


def test_pcg_indefinite_preconditioner():
    a = np.diag([2.0, 3.0, 5.0])
    A = o.OCsr.from_dense(a)
    # a preconditioner with a negative entry makes beta < 0 (pcg.rs:206-213)
    pc = o.OPc.jacobi(o.OCsr.from_dense(np.diag([1.0, -1.0, 1.0])))
    rc, x, st, _ = o.pcg(A, pc, [1, 2, 3], np.zeros(3), 1e-12, 50)
    assert rc in (3, 4)


def test_pcg_natural_norm_first_history_entry_has_no_abs():
    """pcg.rs:137-146: the entry pushed before the loop is dp.sqrt() of the raw r.z (NaN when the preconditioner makes
    it negative); inside the loop (:191) CgNormType::Natural is |r.z|.sqrt()."""
    A = o.OCsr.from_dense(np.diag([2.0, 3.0, 5.0]))
    neg = o.OPc.jacobi(o.OCsr.from_dense(np.diag([-1.0, -1.0, -1.0])))
    rc, x, st, h = o.pcg(A, neg, [1, 2, 3], np.zeros(3), 1e-12, 0, norm_type=2, hist_cap=4)
    assert rc == 0 and st.iterations == 0 and len(h) == 1 and np.isnan(h[0])
    assert st.final_residual == np.sqrt(14.0)                    # res0 = |r.z|.sqrt() (pcg.rs:134)
    for nt, want in ((0, np.sqrt(14.0)), (1, np.sqrt(14.0)), (3, 0.0)):
        rc, x, st, h = o.pcg(A, neg, [1, 2, 3], np.zeros(3), 1e-12, 0, norm_type=nt, hist_cap=4)
        assert h[0] == want
    pos = o.OPc.jacobi(A)
    rc, x, st, h = o.pcg(A, pos, [1, 2, 3], np.zeros(3), 1e-12, 5, norm_type=2, hist_cap=8)
    assert rc == 0 and np.all(np.isfinite(h)) and abs(h[0] - np.sqrt(0.5 + 4.0 / 3.0 + 9.0 / 5.0)) < 1e-15


# ---- (2) independent numpy transliteration ------------------------------------------------------------
def _dense(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("mg", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    kind, size = name.rsplit("_", 1)
    size = int(size)
    if kind.startswith("rowscaled_"):
        return mg.rowscaled(_dense(name[len("rowscaled_"):]))
    if kind == "poisson2d":
        return mg.poisson2d(size)
    if kind == "convdiff2d":
        return mg.convdiff2d(size)
    if kind == "tridiag_spd":
        return tridiag(size, -1, 2, -1)
    if kind == "tridiag_nonsym":
        return tridiag(size, -1, 2, 0.5)
    raise KeyError(name)


@pytest.mark.parametrize("key", sorted(k for k in GS if k.startswith("pcg_")))
def test_golden_pcg(key):
    g = GS[key]
    a = _dense(key.split("/")[1])
    A = o.OCsr.from_dense(a)
    b = a @ np.ones(a.shape[0])
    rc, x, st, h = o.pcg(A, o.OPc.jacobi(A) if key.startswith("pcg_jacobi") else None, b, np.zeros(len(b)), 1e-10, 500, hist_cap=600)
    assert rc == 0 and st.iterations == g["iterations"] and bool(st.converged) == g["converged"]
    assert np.allclose(x, g["x"], rtol=1e-10, atol=1e-12)
    assert np.allclose(h, g["history"], rtol=1e-8, atol=1e-14)


@pytest.mark.parametrize("key", sorted(k for k in GS if k.startswith("gmres_literal_") and "jacobi" in k))
def test_golden_gmres_literal(key):
    g = GS[key]
    mode = {"none": 0, "left": 1, "right": 2}[key.split("_")[2]]
    a = _dense(key.split("/")[1])
    A = o.OCsr.from_dense(a)
    b = a @ np.ones(a.shape[0])
    rc, x, st = o.gmres(A, o.OPc.jacobi(A) if mode else None, b, np.zeros(len(b)), 5, 1e-9, 400, mode=mode, variant=o.GMRES_LITERAL)
    assert st.iterations == g["iterations"] and bool(st.converged) == g["converged"]
    assert np.allclose(x, g["x"], rtol=1e-8, atol=1e-10)


def test_golden_gmres_literal_ilu_literal():
    g = GS["gmres_literal_left_iluliteral/tridiag_nonsym_10"]
    a = tridiag(10, -1, 2, 0.5)
    A = o.OCsr.from_dense(a)
    rc, x, st = o.gmres(A, o.OPc.ilu_literal(a), a @ np.ones(10), np.zeros(10), 10, 1e-12, 100, mode=1, variant=o.GMRES_LITERAL)
    assert st.iterations == g["iterations"] and bool(st.converged) == g["converged"]
    assert np.allclose(x, g["x"], rtol=1e-9)
    z = o.OPc.ilu_literal(tridiag(6, -1, 2, 0.5)).apply(np.arange(1.0, 7.0))
    assert np.allclose(z, GS["ilu_literal_apply/tridiag_nonsym_6"]["z"], rtol=1e-13)
    # the literal Ilu0 is not an ILU (documented deviation F5)
    assert np.allclose(z, [0.237, 1.525, 1.949, 4.602, 3.047, 10.031], atol=2e-3)


@pytest.mark.parametrize("key", sorted(k for k in GS if k.startswith("pcg_sr_")))
def test_golden_pcg_single_reduction(key):
    """the oracle's Chronopoulos-Gear PCG against the independent numpy restatement."""
    g = GS[key]
    a = _dense(key.split("/")[1])
    A = o.OCsr.from_dense(a)
    b = a @ np.ones(a.shape[0])
    rc, x, st, h = o.pcg_sr(A, o.OPc.jacobi(A) if key.startswith("pcg_sr_jacobi") else None, b, np.zeros(len(b)), 1e-10, 500, hist_cap=600)
    assert rc == 0 and st.iterations == g["iterations"] and bool(st.converged) == g["converged"]
    assert np.allclose(x, g["x"], rtol=1e-10, atol=1e-12)
    assert np.allclose(h, g["history"], rtol=1e-6, atol=1e-13)


@pytest.mark.parametrize("key", sorted(k for k in GS if k.startswith("fgmres_literal_")))
def test_golden_fgmres_literal(key):
    """the oracle's fgmres.rs restatement against the independent numpy transliteration, incl. the
    reference's quirk that the reported final_residual is always ||r0|| (fgmres.rs:171,334)."""
    g = GS[key]
    a = _dense(key.split("/")[1])
    A = o.OCsr.from_dense(a)
    b = a @ np.ones(a.shape[0])
    rc, x, st = o.fgmres(A, o.OPc.jacobi(A) if "_jacobi/" in key else None, b, np.zeros(len(b)), 5, 1e-9, 400)
    assert rc == 0 and st.iterations == g["iterations"] and bool(st.converged) == g["converged"]
    assert abs(st.final_residual - g["final_residual"]) <= 1e-12 * g["final_residual"]
    assert np.allclose(x, g["x"], rtol=1e-8, atol=1e-10)


def test_golden_bicgstab_literal():
    g = GS["bicgstab_literal/convdiff2d_8"]
    a = _dense("convdiff2d_8")
    A = o.OCsr.from_dense(a)
    rc, x, st = o.bicgstab(A, None, a @ np.ones(64), np.zeros(64), 1e-9, 400)
    assert st.iterations == g["iterations"] and bool(st.converged) == g["converged"]
    assert np.allclose(x, g["x"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("key", sorted(k for k in GS if k.startswith("ilu0_textbook/")))
def test_golden_ilu0_factors(key):
    a = _dense(key.split("/")[1])
    A = o.OCsr.from_dense(a)
    st, lu, dp, iud, bad = o.ilu0_factor(A)
    assert st == 0
    dense = np.zeros_like(a)
    for i in range(A.n):
        for p in range(int(A.row_ptr[i]), int(A.row_ptr[i + 1])):
            dense[i, int(A.col_idx[p])] = lu[p]
    assert np.allclose(dense, GS[key]["lu"], rtol=1e-13, atol=0)
    assert np.allclose(iud, 1.0 / np.diag(dense), rtol=1e-15)


# ---- (3) survey anchors ----------------------------------------------------------------------------------
def test_anchor_c1_909_iterations():
    c = KA["survey_anchors"]
    A = o.stencil("poisson2d", 512)
    pc = o.OPc.jacobi(A)
    b = o.spmv(A, np.ones(A.n))
    rc, x, st, _ = o.pcg(A, pc, b, np.zeros(A.n), 1e-8, 20000)
    assert st.iterations == c["c1_pcg_jacobi_b_A1"]
    rc, x, st, _ = o.pcg(A, pc, np.ones(A.n), np.zeros(A.n), 1e-8, 20000)
    assert st.iterations == c["c1_pcg_jacobi_b_1"]


def test_single_reduction_pcg_matches_literal_iteration_counts():
    """the Chronopoulos-Gear recurrences are the same Krylov method - iteration counts within 2 %
    of the literal PCG (north_star's bar) and the same solution to solver tolerance; sharded sums stay reproducible."""
    for kind, N in (("poisson2d", 128), ("poisson3d", 24), ("varcoef27", 12)):
        A = o.stencil(kind, N)
        b = o.spmv(A, np.ones(A.n))
        for pc in (None, o.OPc.jacobi(A)):
            rc1, x1, s1, _ = o.pcg(A, pc, b, np.zeros(A.n), 1e-8, 5000)
            rc2, x2, s2, h2 = o.pcg_sr(A, pc, b, np.zeros(A.n), 1e-8, 5000, hist_cap=5001)
            assert rc1 == 0 and rc2 == 0 and s2.converged
            assert abs(int(s1.iterations) - int(s2.iterations)) <= max(1, int(0.02 * s1.iterations))
            assert np.abs(x2 - x1).max() < 1e-6 and len(h2) == s2.iterations + 1
            rc3, x3, s3, _ = o.pcg_sr(A, pc, b, np.zeros(A.n), 1e-8, 5000, nshards=3)
            assert rc3 == 0 and abs(int(s3.iterations) - int(s2.iterations)) <= 1


def test_anchor_c2_small():
    c = KA["survey_anchors"]
    A = o.stencil("convdiff2d", 48)
    b = o.spmv(A, np.ones(A.n))
    z = np.zeros(A.n)
    assert o.gmres(A, o.OPc.ilu0(A), b, z, 30, 1e-8, 20000, mode=1, variant=o.GMRES_MGS2)[2].iterations == c["c2_48_textbook_left_ilu0"]
    assert o.gmres(A, o.OPc.ilu0(A), b, z, 30, 1e-8, 20000, mode=1, variant=o.GMRES_CGS2)[2].iterations == c["c2_48_textbook_left_ilu0"]
    assert o.gmres(A, None, b, z, 30, 1e-8, 20000, mode=0, variant=o.GMRES_LITERAL)[2].iterations == c["c2_48_literal_none"]
    assert o.gmres(A, None, b, z, 30, 1e-8, 20000, mode=0, variant=o.GMRES_CGS2)[2].iterations == c["c2_48_literal_none"]
    assert o.gmres(A, o.OPc.ilu0(A), b, z, 30, 1e-8, 20000, mode=1, variant=o.GMRES_LITERAL)[2].iterations == c["c2_48_literal_left_true_ilu0"]


# ---- (4) independent checks of the Tier-T pieces ------------------------------------------------------------
def test_ilu0_tridiagonal_is_exact_lu():
    sla = pytest.importorskip("scipy.linalg")
    a = tridiag(12, -1.0, 2.0, 0.5)
    A = o.OCsr.from_dense(a)
    st, lu, dp, iud, bad = o.ilu0_factor(A)
    assert st == 0
    r = np.arange(1.0, 13.0)
    z = o.ilu0_apply(A, lu, dp, iud, r)
    assert np.allclose(z, sla.solve(a, r), rtol=1e-12)      # no fill on a tridiagonal: ILU(0) == LU


def test_ilu0_residual_zero_on_pattern():
    A = o.stencil("convdiff3d", 5)
    st, lu, dp, iud, bad = o.ilu0_factor(A)
    a = A.to_dense()
    L = np.eye(A.n)
    U = np.zeros_like(a)
    for i in range(A.n):
        for p in range(int(A.row_ptr[i]), int(A.row_ptr[i + 1])):
            j = int(A.col_idx[p])
            if j < i:
                L[i, j] = lu[p]
            else:
                U[i, j] = lu[p]
    R = L @ U - a
    assert np.abs(R[a != 0]).max() < 1e-13               # defining property of ILU(0): (LU - A) vanishes on the pattern
    z = o.ilu0_apply(A, lu, dp, iud, np.ones(A.n))
    assert np.allclose(L @ (U @ z), np.ones(A.n), rtol=1e-12)


def test_ilu0_zero_pivot_and_missing_diagonal():
    A = o.OCsr.from_dense(np.array([[1.0, 1.0], [1.0, 1.0]]))
    st, lu, dp, iud, bad = o.ilu0_factor(A)
    assert st == 5 and bad == 1                            # KError::ZeroPivot(1)
    B = o.OCsr(2, 2, [0, 1, 2], [1, 0], [1.0, 1.0])
    st, *_rest, bad = o.ilu0_factor(B)
    assert st == 1 and bad == 0


@pytest.mark.parametrize("kind,N,expect", [("poisson2d", 9, 2 * 9 - 1), ("poisson3d", 6, 3 * 6 - 2), ("convdiff2d", 7, 13)])
def test_level_sets_stencil_counts(kind, N, expect):
    A = o.stencil(kind, N)
    for upper in (False, True):
        nl, lev, order, lp = o.levels(A, upper)
        assert nl == expect                                 # SURVEY App. A.2 sanity values
        assert sorted(order.tolist()) == list(range(A.n))
        for l in range(nl):
            seg = order[int(lp[l]):int(lp[l + 1])]
            assert np.all(lev[seg.astype(np.int64)] == l) and np.all(np.diff(seg.astype(np.int64)) > 0)


def test_partition_and_ghosts():
    A = o.stencil("poisson3d", 6)
    n = A.n
    for p in (1, 2, 3, 4, 8):
        cover = []
        for r in range(p):
            lo, hi = o.partition_range(n, p, r)
            chunk = (n + p - 1) // p
            assert lo == min(r * chunk, n) and hi == min((r + 1) * chunk, n)     # asm.rs:46-57
            cover += list(range(lo, hi))
            g = o.ghost_list(A, lo, hi)
            cols = A.col_idx[int(A.row_ptr[lo]):int(A.row_ptr[hi])]
            ref = np.unique(cols[(cols < lo) | (cols >= hi)])
            assert np.array_equal(g, ref)
        assert cover == list(range(n))


def test_canonical_tree_value_and_sharding():
    rng = np.random.default_rng(0)
    for n in (1, 3, 511, 512, 513, 100000):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        assert abs(o.dot(x, y) - float(np.dot(x, y))) <= 1e-12 * float(np.abs(x * y).sum() + 1)
        assert abs(o.dot(x, y, nshards=4) - float(np.dot(x, y))) <= 1e-12 * float(np.abs(x * y).sum() + 1)
    assert o.dot([], []) == 0.0


def test_block_jacobi_ilu0_blocks_match_separate_factorisations():
    A = o.stencil("convdiff3d", 6)
    p = 3
    pc = o.OPc.ilu0(A, nblocks=p)
    r = np.random.default_rng(5).standard_normal(A.n)
    z = pc.apply(r)
    for b in range(p):
        lo, hi = o.partition_range(A.n, p, b)
        sub = A.to_dense()[lo:hi, lo:hi]
        S = o.OCsr.from_dense(sub)
        st, lu, dp, iud, bad = o.ilu0_factor(S)
        assert np.array_equal(z[lo:hi], o.ilu0_apply(S, lu, dp, iud, r[lo:hi]))


def test_jacobi_zero_diagonal_rule():
    A = o.OCsr(3, 3, [0, 1, 2, 3], [0, 2, 2], [2.0, 1.0, 4.0])     # row 1 has no diagonal entry
    assert o.jacobi_inv_diag(A).tolist() == [0.5, 0.0, 0.25]          # jacobi.rs:69-71: zero -> 0


def test_ref_fgmres_2x2_fixture():
    """src/solver/fgmres.rs:531-551: [2 1; 1 3] x = b, x_true = [1, 2], Jacobi as (fixed) flexible preconditioner."""
    a = np.array([[2.0, 1.0], [1.0, 3.0]])
    A = o.OCsr.from_dense(a)
    xt = np.array([1.0, 2.0])
    rc, x, st = o.fgmres(A, o.OPc.jacobi(A), a @ xt, np.zeros(2), 25, 1e-10, 100)
    assert rc == 0 and st.converged and np.abs(x - xt).max() < 1e-6
    # literal quirk: the reported final_residual is the INITIAL residual norm (fgmres.rs:158,337)
    assert st.final_residual == o.norm(a @ xt)
    rc, x, st = o.fgmres(A, None, np.zeros(2), np.zeros(2), 25, 1e-10, 100)      # beta == 0 early return (fgmres.rs:152-154)
    assert st.converged and st.iterations == 0 and st.final_residual == 0.0


def test_cpu_baseline_uses_all_cores_even_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; bench.py's CPU arms must not inherit that."""
    import subprocess
    import sys
    code = ("import sys, os; sys.path.insert(0, %r); import oracle_ffi as o; "
            "print(o.num_threads(), o.use_all_cores(), len(os.sched_getaffinity(0)))" % HERE)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "1" and out[1] == out[2]
