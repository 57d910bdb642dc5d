"""CPU emulation of kb_trsv_march v2: skewed coefficient slots, rhs/out delay lines, mailbox by consumer step, ragged windows."""
import numpy as np
D = 6
def build(gx, gy, gz, n, rng):
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
        s = s - L[2][r]*(y[r-sC] if r-sC>=0 else 0.0)
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
def shape(gx,gy,gz):
    if gz == 1: return gx, 1, gy, 32, 1
    return gx, gy, gz, 8, 4
def skew(L, U, d, n, gx, gy, gz, sB, sC):
    nx, ny, nz, LX, LY = shape(gx,gy,gz)
    px, py = -(-nx//LX), -(-ny//LY)
    nsteps = nz + LX + LY - 2
    lo = np.zeros((3, px*py*nsteps*32)); up = np.zeros((4, px*py*nsteps*32))
    for r in range(n):
        i = r % gx; j = (r//gx) % gy if sB else 0; c0 = r // sC
        a0, b0 = i, j
        pa, pb, qa, qb = a0//LX, b0//LY, a0%LX, b0%LY
        slot = ((pa + px*pb)*nsteps + (c0+qa+qb))*32 + qa + LX*qb
        lo[:, slot] = L[:, r]
        a1, b1, c1 = nx-1-a0, ny-1-b0, nz-1-c0
        pa, pb, qa, qb = a1//LX, b1//LY, a1%LX, b1%LY
        slot = ((pa + px*pb)*nsteps + (c1+qa+qb))*32 + qa + LX*qb
        up[:3, slot] = U[:, r]; up[3, slot] = d[r]
    return lo, up
def march(coef, rhs, n, gx, gy, gz, upper):
    nx, ny, nz, LX, LY = shape(gx,gy,gz)
    SK = LX+LY-2; RD = D+SK+1; OD = SK+1; FACES = 1 if LY == 1 else LX+LY
    px, py = -(-nx//LX), -(-ny//LY)
    nsteps = nz + SK
    mail = {}
    out = np.full(n, np.nan)
    order = sorted(range(px*py), key=lambda p: (p % px) + (p // px))
    plane = nx*ny
    for pencil in order:
        Pa, Pb = pencil % px, pencil // px
        lanes = []
        for lane in range(32):
            la, lb = lane % LX, lane // LX
            ca, cb = Pa*LX+la, Pb*LY+lb
            in_ab = ca < nx and cb < ny
            gi = nx-1-ca if upper else ca; gj = ny-1-cb if upper else cb
            row0 = gi + nx*gj
            def kcount(x):
                room = n - row0 - x
                if not in_ab or room <= 0: return 0
                return min(nz, -(-room//plane))
            if not upper: c_lo, c_hi, cA_lo, cB_lo = 0, kcount(0), 0, 0
            else: c_hi = nz; c_lo = nz-kcount(0); cA_lo = nz-kcount(1); cB_lo = nz-kcount(nx)
            rbase = row0 + plane*(nz-1) if upper else row0
            rstep = -plane if upper else plane
            lanes.append(dict(la=la, lb=lb, sk=la+lb, c_lo=c_lo, c_hi=c_hi, cA_lo=cA_lo, cB_lo=cB_lo, rbase=rbase, rstep=rstep,
                              rring=[None]*RD, oring=[None]*OD, yprev=0.0))
        # loader windows
        pkw = []
        for lane in range(32):
            if lane < FACES:
                cl = lane*LX if lane < LY else (lane-LY)
                skw = lane if lane < LY else lane-LY
                need = (Pa > 0) if lane < LY else (Pb > 0)
                lo_ = lanes[cl]['cA_lo'] if lane < LY else lanes[cl]['cB_lo']
                pkw.append((lo_+skw, lanes[cl]['c_hi']+skw) if need else (0,0))
            else: pkw.append((0,0))
        for t in range(-D, nsteps):
            # prefetch rhs slab t+D at end of step t (prologue: t<0 loads slabs 0..D-1)
            tl = t + D
            if t >= 0:
                pv = [0.0]*32
                for lane in range(32):
                    lo_, hi_ = pkw[lane]
                    if lo_ <= t < hi_:
                        pv[lane] = mail[(pencil, t, lane)]
                new = []
                for lane in range(32):
                    Ln = lanes[lane]
                    c = t - Ln['sk']
                    act = Ln['c_lo'] <= c < Ln['c_hi']
                    slot = (pencil*nsteps + t)*32 + lane
                    vA, vB, vC = coef[0][slot], coef[1][slot], coef[2][slot]
                    rh = Ln['rring'][c % RD] if act else 0.0
                    ya = lanes[lane-1]['yprev'] if Ln['la'] > 0 else pv[Ln['lb']]
                    if LY > 1: yb = lanes[lane-LX]['yprev'] if Ln['lb'] > 0 else pv[LY+Ln['la']]
                    else: yb = 0.0
                    if not act: vA=vB=vC=0.0
                    if not upper:
                        s = rh - vC*Ln['yprev']
                        if LY > 1: s = s - vB*yb
                        s = s - vA*ya
                    else:
                        s = rh - vA*ya
                        if LY > 1: s = s - vB*yb
                        s = s - vC*Ln['yprev']; s = s*(coef[3][slot] if act else 1.0)
                    if not act: s = 0.0
                    if act:
                        Ln['oring'][c % OD] = s
                        if Ln['la'] == LX-1 and Pa+1 < px: mail[(pencil+1, t-(LX-1), Ln['lb'])] = s
                        if LY > 1 and Ln['lb'] == LY-1 and Pb+1 < py: mail[(pencil+px, t-(LY-1), LY+Ln['la'])] = s
                    new.append(s)
                for lane in range(32):
                    lanes[lane]['yprev'] = new[lane]
                    Ln = lanes[lane]
                    cd = t - SK
                    if Ln['c_lo'] <= cd < Ln['c_hi']:
                        out[Ln['rbase'] + Ln['rstep']*cd] = Ln['oring'][cd % OD]
            if tl < nsteps:
                for lane in range(32):
                    Ln = lanes[lane]
                    if Ln['c_lo'] <= tl < Ln['c_hi']:
                        Ln['rring'][tl % RD] = rhs[Ln['rbase'] + Ln['rstep']*tl]
    return out
rng = np.random.default_rng(1)
for (gx,gy,gz,n) in [(10,9,5,10*9*5),(17,6,4,17*6*4-23),(40,7,1,40*7-11),(70,70,1,4900),(9,5,3,9*5*2+20),(14,14,3,14*14*2+37),(14,14,8,14*14*8)]:
    L,U,d,sB,sC = build(gx,gy,gz,n,rng)
    rhs = rng.standard_normal(n)
    y,z = seq(L,U,d,sB,sC,n,rhs)
    lo, up = skew(L,U,d,n,gx,gy,gz,sB,sC)
    ym = march(lo, rhs, n, gx, gy, gz, False)
    zm = march(up, ym, n, gx, gy, gz, True)
    print((gx,gy,gz,n), np.array_equal(y,ym), np.array_equal(z,zm))
