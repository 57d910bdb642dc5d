import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
import kryst_b200 as kb
from kryst_b200 import parallel, _ffi
import ctypes as C
rank, world, local = parallel.dist_env()
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ctx = kb.Context(local); parallel.init_comm(ctx)
n, lo, hi, rp, ci, v = parallel.shard_stencil("poisson3d", 256, world, rank)
A = kb.DeviceCsr.from_csr_shard(n, lo, hi, rp, ci, v, ctx)
x = torch.ones(hi - lo, dtype=torch.float64, device="cuda"); y = torch.zeros(hi - lo, dtype=torch.float64, device="cuda")
for _ in range(5): A.matvec(x, y)
ctx.profile_reset()
# matvec_device = halo exchange (push + recv) + spmv, synchronous
lib = _ffi.lib()
dist.barrier()
t0 = time.perf_counter()
for _ in range(200): A.matvec(x, y)
dt = (time.perf_counter() - t0) / 200
print("rank", rank, "matvec (push+recv+spmv+sync) us", round(dt * 1e6, 1), os.environ.get("KB_DEBUG_PUSH_LOCAL"), flush=True)
dist.barrier(); A.close(); ctx.close(); dist.destroy_process_group()
