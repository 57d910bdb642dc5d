"""CPU emulation of kb_trsv_march's index logic: pencils in level order, lanes (la,lb), c = t-la-lb, mailbox by c."""
import numpy as np, sys
def build(gx, gy, gz, n, rng):
    # full-stencil factor in DIA form, natural coords
    sB = gx if gz > 1 else 0
    sC = gx*gy if gz > 1 else gx
    r = np.arange(n); i = r % gx; j = (r//gx) % gy if sB else 0*r
    L = np.zeros((3,n)); U = np.zeros((3,n))
    v = lambda: rng.standard_normal(n)*0.3
    L[0] = np.where(i>=1, v(), 0); L[1] = np.where((j>=1) & (sB>0), v(), 0); L[2] = np.where(r-sC>=0, v(), 0)
    U[0] = np.where((i+1<gx)&(r+1<n), v(), 0); U[1] = np.where((sB>0)&(j+1<gy)&(r+sB<n), v(), 0); U[2] = np.where(r+sC<n, v(), 0)
    d = 1.0/(1.0+rng.random(n))
    return L, U, d, sB, sC
def seq(L,U,d,sB,sC,n,rhs):
    y = np.zeros(n)
    for r in range(n):
        s = rhs[r]
        if r-sC>=0: s = s - L[2][r]*y[r-sC]
        else: s = s - 0.0
        if sB: s = s - L[1][r]*(y[r-sB] if r-sB>=0 else 0.0)
        s = s - L[0][r]*(y[r-1] if r>=1 else 0.0)
        y[r] = s
    z = np.zeros(n)
    for r in range(n-1,-1,-1):
        s = y[r]
        s = s - U[0][r]*(z[r+1] if r+1<n else 0.0)
        if sB: s = s - U[1][r]*(z[r+sB] if r+sB<n else 0.0)
        s = s - U[2][r]*(z[r+sC] if r+sC<n else 0.0)
        z[r] = s*d[r]
    return y, z
def march(coef, dg, rhs, n, gx, gy, gz, upper):
    two_d = gz == 1
    if two_d: nx, ny, nz, LX, LY = gx, 1, gy, 32, 1
    else: nx, ny, nz, LX, LY = gx, gy, gz, 8, 4
    px, py = -(-nx//LX), -(-ny//LY)
    faces = LX+LY
    mail = {}
    out = np.full(n, np.nan)
    order = sorted(range(px*py), key=lambda p: (p % px) + (p // px))
    plane = nx*ny
    nsteps = nz + LX + LY - 2
    for pencil in order:
        Pa, Pb = pencil % px, pencil // px
        yprev = np.zeros(32)
        for t in range(nsteps):
            new = np.zeros(32)
            for lane in range(32):
                la, lb = lane % LX, lane // LX
                ca, cb = Pa*LX+la, Pb*LY+lb
                in_ab = ca < nx and cb < ny
                gi = nx-1-ca if upper else ca; gj = ny-1-cb if upper else cb
                c = t-la-lb
                row = gi + nx*gj + plane*((nz-1-c) if upper else c)
                act = in_ab and 0 <= c < nz and row < n
                if not act: continue
                ya = yprev[lane-1] if la > 0 else (mail[(pencil,0,lb,c)] if Pa > 0 else 0.0)
                yb = yprev[lane-LX] if lb > 0 else (mail[(pencil,1,la,c)] if Pb > 0 else 0.0)
                vA, vB, vC = coef[0][row], coef[1][row], coef[2][row]
                if not upper:
                    s = rhs[row] - vC*yprev[lane]; s = s - vB*yb; s = s - vA*ya
                else:
                    s = rhs[row] - vA*ya; s = s - vB*yb; s = s - vC*yprev[lane]; s = s*dg[row]
                new[lane] = s; out[row] = s
                if la == LX-1 and Pa+1 < px: mail[(pencil+1,0,lb,c)] = s
                if lb == LY-1 and Pb+1 < py: mail[(pencil+px,1,la,c)] = s
            yprev = new
    return out
rng = np.random.default_rng(1)
for (gx,gy,gz,n) in [(10,9,5,10*9*5),(17,6,4,17*6*4-23),(40,7,1,40*7-11),(70,70,1,4900),(9,5,3,9*5*2+20)]:
    L,U,d,sB,sC = build(gx,gy,gz,n,rng)
    rhs = rng.standard_normal(n)
    y,z = seq(L,U,d,sB,sC,n,rhs)
    ym = march(L, d, rhs, n, gx, gy, gz, False)
    zm = march(U, d, ym, n, gx, gy, gz, True)
    print((gx,gy,gz,n), np.array_equal(y,ym), np.array_equal(z,zm))
