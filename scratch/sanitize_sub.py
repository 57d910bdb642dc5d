import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import kryst_b200 as kb
from kryst_b200 import stencils
ctx = kb.Context(0)
n, rp, ci, v = stencils.stencil("poisson3d", 9)
A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
rng = np.random.default_rng(1)
for idx in (np.arange(100, 500), rng.permutation(n), rng.integers(0, n, 300), np.zeros(0, dtype=np.int64)):
    S = A.submatrix(idx)
    r, c, w = S.to_csr()
    print(S.nrows(), S.nnz(), int(r[-1]))
x = np.zeros(n); b = np.zeros(n); A.matvec(np.ones(n), b)
st = kb.PcgSolver(1e-8, 50).with_fused_reduction().solve(A, kb.Jacobi().setup(A), b, x); print(st.iterations, st.converged)
print("SANITIZE_SUB_OK")
