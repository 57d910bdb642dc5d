"""Timing ablations of kb_pcg_resident on C1 (512^2): fixed 1000 iterations, us per iteration for each KB_RES_DEBUG mode."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kryst_b200 as kb
from kryst_b200 import stencils
ctx = kb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
N = int(os.environ.get("ABL_N", "512"))
n, rp, ci, v = stencils.stencil("poisson2d", N)
A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
pc = kb.Jacobi().setup(A)
ones = torch.ones(n, dtype=torch.float64, device="cuda")
b = torch.zeros(n, dtype=torch.float64, device="cuda")
x = torch.zeros(n, dtype=torch.float64, device="cuda")
A.matvec(ones, b)
for mode in [int(a) for a in sys.argv[1:]]:
    os.environ["KB_RES_DEBUG"] = str(mode | 16)
    best = {}
    for iters in (200, 1200):
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
    us = 1e3 * (best[1200][0] - best[200][0]) / (best[1200][1] - best[200][1])
    print(json.dumps({"N": N, "dbg": mode, "us_per_iteration": us, "ms_200": best[200][0], "ms_1200": best[1200][0], "iters": [best[200][1], best[1200][1]]}), flush=True)
