"""Timeline of the packet-face tile solves: per tile pick-up / preload / distance-wait / end times (ns)."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["KB_TILES_TRACE"] = "1"
import kryst_b200 as kb
from kryst_b200 import stencils, _ffi
import torch
kind, N = sys.argv[1], int(sys.argv[2])
ctx = kb.default_context(0)
n, rp, ci, v = stencils.stencil(kind, N)
A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
pc = kb.Ilu0().setup(A)
r = torch.randn(n, dtype=torch.float64, device="cuda"); z = torch.zeros_like(r)
lib = _ffi.lib()
lib.kb_debug_tiles_trace.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
stream = torch.cuda.ExternalStream(ctx.stream)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream); pc.apply(r, z); e1.record(stream); e1.synchronize()
    print("apply ms", e0.elapsed_time(e1))
buf = np.zeros(8 * 1 << 20, dtype=np.uint64)
tx, ty, tz = C.c_int(0), C.c_int(0), C.c_int(0)
nt = lib.kb_debug_tiles_trace(pc.handle, buf.ctypes.data, C.byref(tx), C.byref(ty), C.byref(tz))
print("tiles", nt, tx.value, ty.value, tz.value)
t = buf[:nt * 4].reshape(nt, 4).astype(np.int64)
t0 = t[:, 0].min()
pick, pre, wait, end = (t[:, k] - t0 for k in range(4))
print("L total us %.1f" % (end.max() / 1e3), " tile(0,0,0): preload %.2f steps %.2f" % ((pre-pick)[0]/1e3, (end-wait)[0]/1e3))
print("per tile us: preload med %.2f  distance-wait med %.2f  steps med %.2f (min %.2f max %.2f)  total med %.2f" % (
    np.median(pre - pick) / 1e3, np.median(wait - pre) / 1e3, np.median(end - wait) / 1e3, (end - wait).min() / 1e3, (end - wait).max() / 1e3, np.median(end - pick) / 1e3))
T = end.reshape(tz.value, ty.value, tx.value)
print("end time us along I (J=K=0):", np.round(T[0, 0, :8] / 1e3, 1))
print("end time us along the diagonal:", np.round(np.array([T[k, k, k] for k in range(min(8, tx.value))]) / 1e3, 1))
S = (end - wait).reshape(tz.value, ty.value, tx.value)
print("steps us along the diagonal:", np.round(np.array([S[k, k, k] for k in range(min(8, tx.value))]) / 1e3, 1))
