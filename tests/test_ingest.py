"""the step before the path - sub-matrix extraction (SubmatrixExtract, sparse.rs:72-93) and Matrix
Market ingestion.  CPU part: the oracle against the reference's literal dense definition, and the host I/O.  GPU part
(marked): device extraction vs the oracle, bit-exact, through the C ABI."""
import os

import numpy as np
import pytest

import oracle_ffi as o


def _dense_submatrix(A, idx):
    """The reference's own definition: densify, gather, drop zeros (sparse.rs:74-90)."""
    d = A.to_dense()
    sub = d[np.ix_(idx, idx)]
    rp, ci, v = [0], [], []
    for i in range(len(idx)):
        for j in range(len(idx)):
            if sub[i, j] != 0.0:
                ci.append(j)
                v.append(sub[i, j])
        rp.append(len(ci))
    return np.array(rp, dtype=np.uint64), np.array(ci, dtype=np.uint64), np.array(v, dtype=np.float64)


def _rand_csr(n, m, density, seed, zeros=True):
    rng = np.random.default_rng(seed)
    rp, ci, v = [0], [], []
    for i in range(n):
        cols = np.flatnonzero(rng.random(m) < density)
        vals = rng.standard_normal(cols.size)
        if zeros and cols.size:
            vals[rng.integers(0, cols.size)] = 0.0        # stored explicit zero: must be dropped by submatrix
        ci.extend(cols.tolist())
        v.extend(vals.tolist())
        rp.append(len(ci))
    return o.OCsr(n, m, rp, ci, v)


SUB_CASES = [
    ("sorted-subset", lambda n, r: np.sort(r.choice(n, n // 2, replace=False))),
    ("permutation", lambda n, r: r.permutation(n)),
    ("unsorted-subset", lambda n, r: r.choice(n, n // 3, replace=False)),
    ("with-repeats", lambda n, r: r.integers(0, n, n // 2)),
    ("empty", lambda n, r: np.zeros(0, dtype=np.int64)),
    ("single", lambda n, r: np.array([n - 1])),
    ("chunk-block", lambda n, r: np.arange(n // 4, n // 2)),          # asm.rs:46-57 subdomain
]


@pytest.mark.parametrize("name,mk", SUB_CASES)
def test_oracle_submatrix_equals_reference_dense_definition(name, mk):
    A = _rand_csr(60, 60, 0.15, seed=3)
    idx = mk(60, np.random.default_rng(5))
    S = o.submatrix(A, idx)
    rp, ci, v = _dense_submatrix(A, idx)
    assert np.array_equal(S.row_ptr, rp) and np.array_equal(S.col_idx, ci) and np.array_equal(S.vals, v)


def test_oracle_submatrix_stencil_block_and_range_check():
    A = o.stencil("poisson2d", 12)
    idx = np.arange(36, 96)
    S = o.submatrix(A, idx)
    rp, ci, v = _dense_submatrix(A, idx)
    assert np.array_equal(S.row_ptr, rp) and np.array_equal(S.col_idx, ci) and np.array_equal(S.vals, v)
    with pytest.raises(IndexError):
        o.submatrix(A, [0, A.n])
    R = _rand_csr(10, 7, 0.5, seed=1)          # rectangular: indices address rows and columns
    with pytest.raises(IndexError):
        o.submatrix(R, [8])


# ---- Matrix Market host I/O (no GPU, no CUDA library: import the module by path) ---------------------------------
def _mmio():
    import importlib.util
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kryst_b200", "mmio.py")
    spec = importlib.util.spec_from_file_location("kb_mmio", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_matrix_market_round_trip_is_exact(tmp_path):
    mm = _mmio()
    A = _rand_csr(40, 33, 0.2, seed=9, zeros=True)
    p = str(tmp_path / "a.mtx")
    mm.write_matrix_market(p, A.n, A.ncols, A.row_ptr, A.col_idx, A.vals, comment="round trip")
    n, m, rp, ci, v = mm.read_matrix_market(p)
    assert (n, m) == (A.n, A.ncols)
    assert np.array_equal(rp, A.row_ptr) and np.array_equal(ci, A.col_idx) and np.array_equal(v, A.vals)   # bit-exact f64
    assert int(o.lib().ko_csr_validate(o.OCsr(n, m, rp, ci, v).ptr())) == 0          # new_checked rules hold


def test_matrix_market_symmetric_pattern_duplicates_and_errors(tmp_path):
    mm = _mmio()
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n% lower triangle\n3 3 4\n1 1 2.0\n2 1 -1.0\n3 2 -1.0\n3 3 2.0\n")
    n, m, rp, ci, v = mm.read_matrix_market(str(p))
    d = o.OCsr(n, m, rp, ci, v).to_dense()
    assert np.array_equal(d, np.array([[2.0, -1.0, 0.0], [-1.0, 0.0, -1.0], [0.0, -1.0, 2.0]]))
    p.write_text("%%MatrixMarket matrix coordinate pattern general\n2 3 3\n1 3\n2 1\n1 1\n")
    n, m, rp, ci, v = mm.read_matrix_market(str(p))
    assert (n, m) == (2, 3) and rp.tolist() == [0, 2, 3] and ci.tolist() == [0, 2, 0] and v.tolist() == [1.0, 1.0, 1.0]
    p.write_text("%%MatrixMarket matrix coordinate real general\n2 2 4\n1 1 1.0\n2 2 5.0\n1 1 0.25\n1 1 0.5\n")
    n, m, rp, ci, v = mm.read_matrix_market(str(p))          # repeats are summed in file order
    assert rp.tolist() == [0, 1, 2] and v.tolist() == [(1.0 + 0.25) + 0.5, 5.0]
    p.write_text("%%MatrixMarket matrix coordinate real skew-symmetric\n2 2 1\n2 1 3.0\n")
    n, m, rp, ci, v = mm.read_matrix_market(str(p))
    assert np.array_equal(o.OCsr(n, m, rp, ci, v).to_dense(), np.array([[0.0, -3.0], [3.0, 0.0]]))
    p.write_text("%%MatrixMarket matrix coordinate real general\n0 0 0\n")
    n, m, rp, ci, v = mm.read_matrix_market(str(p))
    assert (n, m) == (0, 0) and rp.tolist() == [0] and ci.size == 0
    for bad in ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
                "%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 0.0\n",
                "%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n",
                "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n",
                "not a header\n"):
        p.write_text(bad)
        with pytest.raises(mm.MatrixMarketError):
            mm.read_matrix_market(str(p))


# ---- device extraction ----------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,mk", SUB_CASES)
def test_gpu_submatrix_bit_exact(ctx, name, mk):
    import kryst_b200 as kb
    Ao = _rand_csr(300, 300, 0.05, seed=11)
    A = kb.DeviceCsr.from_csr(Ao.n, Ao.ncols, Ao.row_ptr, Ao.col_idx, Ao.vals, ctx)
    idx = mk(300, np.random.default_rng(2))
    S = A.submatrix(idx)
    So = o.submatrix(Ao, idx)
    rp, ci, v = S.to_csr()
    assert (S.nrows(), S.ncols(), S.nnz()) == (So.n, So.ncols, So.nnz)
    assert np.array_equal(rp, So.row_ptr) and np.array_equal(ci, So.col_idx) and np.array_equal(v, So.vals)
    if So.n:
        x = np.random.default_rng(4).standard_normal(So.n)
        y = np.zeros(So.n)
        S.matvec(x, y)
        assert np.array_equal(y, o.spmv(So, x))


@pytest.mark.gpu
def test_gpu_submatrix_feeds_the_path_like_asm_setup(ctx, tmp_path):
    """asm.rs:46-65: chunk subdomains -> submatrix -> local solve; plus Matrix Market in, download out."""
    import kryst_b200 as kb
    n, rp, ci, v = kb.stencils.stencil("poisson3d", 16)
    p = str(tmp_path / "p3d.mtx")
    kb.mmio.write_matrix_market(p, n, n, rp, ci, v)
    n2, m2, rp2, ci2, v2 = kb.mmio.read_matrix_market(p)
    A = kb.DeviceCsr.from_csr(n2, m2, rp2, ci2, v2, ctx)
    Ao = o.OCsr(n, n, rp, ci, v)
    d = A.to_csr()
    assert np.array_equal(d[0], Ao.row_ptr) and np.array_equal(d[1], Ao.col_idx) and np.array_equal(d[2], Ao.vals)
    lo, hi = kb.partition_range(n, 3, 1)
    idx = np.arange(lo, hi)
    S, So = A.submatrix(idx), o.submatrix(Ao, idx)
    b = o.spmv(So, np.ones(So.n))
    x = np.zeros(So.n)
    st = kb.GmresSolver(20, 1e-10, 400).solve(S, kb.Ilu0().setup(S), b, x)
    rc, xo, so = o.gmres(So, o.OPc.ilu0(So), b, np.zeros(So.n), 20, 1e-10, 400, mode=o.MODE_LEFT, variant=o.GMRES_CGS2)
    assert rc == 0 and (st.iterations, st.converged) == (so.iterations, bool(so.converged)) and np.array_equal(x, xo)
    with pytest.raises(kb.SolveError):
        A.submatrix([0, n])


@pytest.mark.gpu
def test_gpu_download_and_submatrix_edges(ctx):
    import kryst_b200 as kb
    E = kb.DeviceCsr.from_csr(0, 0, [0], [], [], ctx)
    assert [a.size for a in E.to_csr()] == [1, 0, 0]
    Z = kb.DeviceCsr.from_csr(3, 3, [0, 1, 2, 3], [0, 1, 2], [0.0, -0.0, 5.0], ctx)      # stored zeros are dropped
    rp, ci, v = Z.submatrix([2, 1, 0]).to_csr()
    assert rp.tolist() == [0, 1, 1, 1] and ci.tolist() == [0] and v.tolist() == [5.0]


# ---- C++ host reader (include/kryst_mmio.hpp) against the Python one ------------------------------------------------
def _cpp_mmio(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_mmio")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "test_mmio.cpp"), "-o", exe])
    return exe


def _run_cpp(exe, *args):
    import subprocess
    r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True)
    if r.returncode != 0:
        return r.returncode, r.stderr
    lines = r.stdout.split("\n")
    n, m, nnz = (int(t) for t in lines[0].split())
    rp = np.array([int(t) for t in lines[1].split()], dtype=np.uint64)
    ci = np.array([int(t) for t in lines[2].split()], dtype=np.uint64)
    v = np.array([float.fromhex(t) for t in lines[3].split()], dtype=np.float64)
    assert ci.size == nnz and v.size == nnz
    return 0, (n, m, rp, ci, v)


def test_cpp_matrix_market_reader_matches_python(tmp_path):
    mm = _mmio()
    exe = _cpp_mmio(tmp_path)
    A = _rand_csr(30, 41, 0.2, seed=4, zeros=True)
    p, q = tmp_path / "a.mtx", tmp_path / "b.mtx"
    mm.write_matrix_market(str(p), A.n, A.ncols, A.row_ptr, A.col_idx, A.vals)
    rc, (n, m, rp, ci, v) = _run_cpp(exe, p, q)
    assert rc == 0 and (n, m) == (A.n, A.ncols)
    assert np.array_equal(rp, A.row_ptr) and np.array_equal(ci, A.col_idx) and np.array_equal(v, A.vals)
    n2, m2, rp2, ci2, v2 = mm.read_matrix_market(str(q))                    # C++ writer -> Python reader: exact round trip
    assert np.array_equal(rp2, A.row_ptr) and np.array_equal(ci2, A.col_idx) and np.array_equal(v2, A.vals)
    cases = ["%%MatrixMarket matrix coordinate real symmetric\n% lower triangle\n3 3 4\n1 1 2.0\n2 1 -1.0\n3 2 -1.0\n3 3 2.0\n",
             "%%MatrixMarket matrix coordinate pattern general\n2 3 3\n1 3\n2 1\n1 1\n",
             "%%MatrixMarket matrix coordinate real general\n2 2 4\n1 1 1.0\n2 2 5.0\n1 1 0.25\n1 1 0.5\n",
             "%%MatrixMarket matrix coordinate real skew-symmetric\n2 2 1\n2 1 3.0\n",
             "%%MatrixMarket matrix coordinate integer symmetric\n4 4 5\n1 1 4\n4 1 -2\n2 2 4\n4 1 -1\n3 3 7\n",
             "%%MatrixMarket matrix coordinate real general\n0 0 0\n"]
    for text in cases:
        p.write_text(text)
        want = mm.read_matrix_market(str(p))
        rc, got = _run_cpp(exe, p)
        assert rc == 0 and got[:2] == want[:2], text
        assert all(np.array_equal(g, w) for g, w in zip(got[2:], want[2:])), text
    for bad in ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
                "%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 0.0\n",
                "%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n",
                "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n",
                "%%MatrixMarket matrix coordinate real skew-symmetric\n2 2 1\n1 1 3.0\n",
                "not a header\n"):
        p.write_text(bad)
        rc, msg = _run_cpp(exe, p)
        assert rc == 2 and "MatrixMarketError" in msg, bad
        with pytest.raises(mm.MatrixMarketError):
            mm.read_matrix_market(str(p))
