"""Free-running step time of the lean march on a single-group grid (nx x ny x nz), with KB_LEAN_DBG experiments."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["KB_MARCH_TRACE"] = "1"
os.environ["KB_TRSV_MARCH"] = "1"
import kryst_b200 as kb
from kryst_b200 import _ffi
import torch
nx, ny, nz = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
def poisson(nx, ny, nz):
    n = nx * ny * nz
    r = np.arange(n); i = r % nx; j = (r // nx) % ny; k = r // (nx * ny)
    cols = [r - nx * ny, r - nx, r - 1, r, r + 1, r + nx, r + nx * ny]
    ok = [k >= 1, j >= 1, i >= 1, np.ones(n, bool), i + 1 < nx, j + 1 < ny, k + 1 < nz]
    vals = [-1.0, -1.0, -1.0, 6.0, -1.0, -1.0, -1.0]
    C_ = np.stack(cols, 1); M = np.stack(ok, 1); V = np.broadcast_to(np.array(vals), (n, 7))
    rp = np.concatenate([[0], np.cumsum(M.sum(1))]).astype(np.uint64)
    return n, rp, C_[M].astype(np.uint64), V[M].astype(np.float64).copy()
ctx = kb.default_context(0)
n, rp, ci, v = poisson(nx, ny, nz)
A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
pc = kb.Ilu0().setup(A)
r = torch.randn(n, dtype=torch.float64, device="cuda"); z = torch.zeros_like(r)
lib = _ffi.lib()
lib.kb_debug_march_trace.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
for rep in range(3):
    pc.apply(r, z)
buf = np.zeros(8 * 1 << 20, dtype=np.uint64)
px, py = C.c_int(0), C.c_int(0)
npn = lib.kb_debug_march_trace(pc.handle, buf.ctypes.data, C.byref(px), C.byref(py))
out = []
for u, name in ((0, "L"), (1, "U")):
    t = buf[u * npn * 4:(u + 1) * npn * 4].reshape(npn, 4).astype(np.int64)
    run = (t[:, 2] - t[:, 1])
    out.append("%s step ns %.0f" % (name, run[0] / (nz + 14.0)))
print("dbg", os.environ.get("KB_LEAN_DBG", "0"), "grid", nx, ny, nz, "pencils", npn, " ".join(out))
