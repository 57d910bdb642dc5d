"""Timeline of the pencil-march solves: per-pencil entry/first/end times (ns) -> free-run step time, hop lag."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["KB_MARCH_TRACE"] = "1"
import kryst_b200 as kb
from kryst_b200 import stencils, _ffi
import torch
kind, N = sys.argv[1], int(sys.argv[2])
ctx = kb.default_context(0)
n, rp, ci, v = stencils.stencil(kind, N)
A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
os.environ["KB_MARCH_GROUP"] = sys.argv[3] if len(sys.argv) > 3 else "4"
os.environ["KB_TRSV_MARCH"] = "1"
pc = kb.Ilu0().setup(A)
r = torch.randn(n, dtype=torch.float64, device="cuda"); z = torch.zeros_like(r)
lib = _ffi.lib()
lib.kb_debug_march_trace.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
stream = torch.cuda.ExternalStream(ctx.stream)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream); pc.apply(r, z); e1.record(stream); e1.synchronize()
    print("apply ms", e0.elapsed_time(e1))
buf = np.zeros(8 * 1 << 22, dtype=np.uint64)
px, py = C.c_int(0), C.c_int(0)
npn = lib.kb_debug_march_trace(pc.handle, buf.ctypes.data, C.byref(px), C.byref(py))
print("pencils", npn, px.value, py.value)
for u, name in ((0, "L"), (1, "U")):
    t = buf[u * npn * 4:(u + 1) * npn * 4].reshape(npn, 4).astype(np.int64)
    t0 = t[:, 0].min()
    ent, first, end, stalls = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3]
    run = end - first
    print(name, "total us %.1f" % (end.max() / 1e3), "run/pencil us: min %.1f med %.1f max %.1f" % (run.min() / 1e3, np.median(run) / 1e3, run.max() / 1e3),
          "stalls/pencil med %d max %d" % (np.median(stalls), stalls.max()))
    P = first.reshape(py.value, px.value)
    print("  first-step time us along a (b=0):", np.round(P[0, :min(8, px.value)] / 1e3, 1))
    print("  first-step time us along b (a=0):", np.round(P[:min(8, py.value), 0] / 1e3, 1))
    Pe = ent.reshape(py.value, px.value)
    print("  entry us along diag:", np.round(np.array([Pe[min(i * py.value // 8, py.value - 1), min(i * px.value // 8, px.value - 1)] for i in range(9)]) / 1e3, 1))
    print("  first us along diag:", np.round(np.array([P[min(i * py.value // 8, py.value - 1), min(i * px.value // 8, px.value - 1)] for i in range(9)]) / 1e3, 1))
    E = end.reshape(py.value, px.value)
    print("  end   us along diag:", np.round(np.array([E[min(i * py.value // 8, py.value - 1), min(i * px.value // 8, px.value - 1)] for i in range(9)]) / 1e3, 1))
    print("  total stalls", stalls.sum(), " groups with stalls", (stalls > 0).sum())
    print("  end us corner:", E[-1, -1] / 1e3, " run of pencil 0 us:", run.reshape(py.value, px.value)[0, 0] / 1e3)
