"""C1 (512^2) with the resident PCG at several grid sizes: us per iteration (fixed iteration counts, real arithmetic)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kryst_b200 as kb
from kryst_b200 import stencils
ctx = kb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
n, rp, ci, v = stencils.stencil("poisson2d", 512)
A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
pc = kb.Jacobi().setup(A)
ones = torch.ones(n, dtype=torch.float64, device="cuda")
b = torch.zeros(n, dtype=torch.float64, device="cuda")
x = torch.zeros(n, dtype=torch.float64, device="cuda")
A.matvec(ones, b)
for grid in sys.argv[1:]:
    os.environ["KB_RES_GRID"] = grid
    best = {}
    for iters in (100, 600):
        s = kb.PcgSolver(1e-30, iters)
        s.record_history = False
        t = None
        for rep in range(4):
            x.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            st = s.solve(A, pc, b, x)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            if rep and (t is None or ms < t):
                t = ms
        best[iters] = (t, st.iterations)
    print(json.dumps({"grid": grid, "us_per_iteration": 1e3 * (best[600][0] - best[100][0]) / (best[600][1] - best[100][1]), "iters": [best[100][1], best[600][1]]}), flush=True)
