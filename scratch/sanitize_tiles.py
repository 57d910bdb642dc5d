"""Tile triangular solves under compute-sanitizer (racecheck)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import kryst_b200 as kb
from kryst_b200 import stencils
ctx = kb.Context(0)
for kind, N in (("poisson3d", 11), ("convdiff2d", 37)):
    n, rp, ci, v = stencils.stencil(kind, N)
    A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    ilu = kb.Ilu0().setup(A)
    r = np.arange(n, dtype=np.float64) / n + 1.0
    z = np.zeros(n); ilu.apply(r, z)
    print(kind, N, float(z.sum()))
print("SANITIZE_TILES_OK")
