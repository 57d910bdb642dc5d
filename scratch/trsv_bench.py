import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import kryst_b200 as kb
from kryst_b200 import stencils
ctx = kb.Context(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = sys.argv[2] if len(sys.argv) > 2 else "poisson3d"
n, rp, ci, v = stencils.stencil(kind, N)
A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
del rp, ci, v
r = torch.ones(n, dtype=torch.float64, device="cuda"); z = torch.zeros(n, dtype=torch.float64, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream)
def run(tag, **env):
    for k in ("KB_TRSV_KIND","KB_TRSV_GRID","KB_TRSV_WINDOW","KB_TRSV_SLEEP","KB_TRSV_GATE"): os.environ.pop(k, None)
    for k, val in env.items(): os.environ[k] = str(val)
    pc = kb.Ilu0().setup(A)
    pc.apply(r, z); torch.cuda.synchronize()
    ctx.profile_reset()
    t0 = time.perf_counter()
    for _ in range(5): pc.apply(r, z)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(tag, env, "apply ms %.3f" % (dt * 1e3), "checksum %.6e" % float(z.sum().item()), flush=True)
    pc.close()
run("persist default")
run("persist", KB_TRSV_GATE=2)
run("persist", KB_TRSV_GATE=4)
