"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import kryst_b200 as kb
from kryst_b200 import stencils
ctx = kb.Context(0)
for kind, N in (("poisson2d", 20), ("convdiff3d", 9), ("varcoef27", 6)):
    n, rp, ci, v = stencils.stencil(kind, N)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    b = np.zeros(n); A.matvec(np.ones(n), b)
    x = np.zeros(n); st = kb.PcgSolver(1e-8, 60).solve(A, kb.Jacobi().setup(A), b, x); print(kind, "pcg", st)
    x = np.zeros(n); st = kb.BiCgStabSolver(1e-8, 60, textbook=True).solve(A, kb.Jacobi().setup(A), b, x); print(kind, "bicg", st)
    ilu = kb.Ilu0().setup(A)
    x = np.zeros(n); st = kb.GmresSolver(6, 1e-8, 40).solve(A, ilu, b, x); print(kind, "gmres", st)
    print(ctx.dot(b, b), ctx.norm(b))
print("SANITIZE_RUN_OK")
