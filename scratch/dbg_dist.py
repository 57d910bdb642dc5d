import os, sys, faulthandler
faulthandler.enable()
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch, torch.distributed as dist
import kryst_b200 as kb
from kryst_b200 import parallel
rank, world, local = parallel.dist_env()
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ctx = kb.Context(local); parallel.init_comm(ctx)
def P(*a):
    print("[r%d]" % rank, *a, flush=True)
n, lo, hi, rp, ci, v = parallel.shard_stencil("poisson3d", 12, world, rank)
A = kb.DeviceCsr.from_csr_shard(n, lo, hi, rp, ci, v, ctx)
P("csr ok")
pc = kb.Ilu0().setup(A); P("ilu setup ok")
pc.close(); P("ilu close ok (no use)")
pc = kb.Ilu0().setup(A)
r = np.ones(hi-lo); z = np.zeros(hi-lo); pc.apply(r, z); P("apply ok")
pc.close(); P("close after apply ok")
pc = kb.Ilu0().setup(A)
b = np.ones(hi-lo); x = np.zeros(hi-lo)
st = kb.GmresSolver(10, 1e-8, 300).with_preconditioning(1).solve(A, pc, b, x); P("gmres ok", st)
pc.close(); P("close after gmres ok")
A.close(); P("A close ok")
ctx.close(); P("ctx closed")
