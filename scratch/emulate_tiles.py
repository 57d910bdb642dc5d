"""CPU emulation of kb_trsv_tiles (same index logic) against the oracle's ILU(0) apply."""
import sys, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import oracle_ffi as o

def detect(rp, col, n):
    smin, smax = 2**31 - 1, 0
    for r in range(n):
        for p in range(rp[r], rp[r + 1]):
            d = abs(r - col[p])
            if d > 1: smin, smax = min(smin, d), max(smax, d)
    if smax == 0 or smin < 2 or (smax != smin and smax % smin): return None
    nx = smin; sy = smax
    ny = (n + nx - 1) // nx if smax == smin else smax // smin
    nz = 1 if smax == smin else (n + smax - 1) // smax
    three_d = sy > nx
    for r in range(n):
        i, j = r % nx, (r // nx) % ny
        for p in range(rp[r], rp[r + 1]):
            c = col[p]
            if c == r: continue
            d = abs(r - c)
            if d == 1: ok = (i >= 1) if c < r else (i + 1 < nx)
            elif d == nx: ok = (not three_d) or ((j >= 1) if c < r else (j + 1 < ny))
            elif three_d and d == sy: ok = True
            else: ok = False
            if not ok: return None
    return nx, ny, nz

def solve(rp, col, lu, dptr, invd, rhs, n, grid, upper, T=256):
    nx, ny, nz = grid
    bx, by, bz, R = (32, 32, 1, 4) if nz == 1 else (8, 8, 8, 2)
    tx, ty, tz = -(-nx // bx), -(-ny // by), -(-nz // bz)
    nt = tx * ty * tz
    ids = sorted(range(nt), key=lambda t: (t % tx) + ((t // tx) % ty) + t // (tx * ty))
    if upper: ids = ids[::-1]
    flags = np.zeros(nt, dtype=int)
    out = np.full(n, np.nan)
    bxy, trows, nlevels = bx * by, bx * by * bz, bx + by + bz - 2
    sx, sy = nx, nx * ny
    for tile in ids:
        TI, TJ, TK = tile % tx, (tile // tx) % ty, tile // (tx * ty)
        for tid in range(1, 8):
            da, db, dc = tid & 1, (tid >> 1) & 1, (tid >> 2) & 1
            PI, PJ, PK = (TI + da, TJ + db, TK + dc) if upper else (TI - da, TJ - db, TK - dc)
            if 0 <= PI < tx and 0 <= PJ < ty and 0 <= PK < tz:
                assert flags[PI + tx * (PJ + ty * PK)] == 1, "predecessor not finished: deadlock in the real kernel"
        ytile = np.full(1024, np.nan)
        rows = []
        for q in range(R * T):
            if q >= trows: continue
            li, lj, lk = q % bx, (q // bx) % by, q // bxy
            gi, gj, gk = TI * bx + li, TJ * by + lj, TK * bz + lk
            r = gi + nx * (gj + ny * gk)
            if not (gi < nx and gj < ny and gk < nz and r < n): continue
            pd = dptr[r]
            p0, p1 = (pd + 1, rp[r + 1]) if upper else (rp[r], pd)
            assert p1 - p0 <= 3
            ent = []
            for p in range(p0, p1):
                c = col[p]; d = c - r if upper else r - c
                slot = -1
                if d == 1:
                    if (li + 1 < bx) if upper else (li >= 1): slot = q + 1 if upper else q - 1
                elif d == sx:
                    if (lj + 1 < by) if upper else (lj >= 1): slot = q + bx if upper else q - bx
                else:
                    if (lk + 1 < bz) if upper else (lk >= 1): slot = q + bxy if upper else q - bxy
                ent.append((slot, lu[p], out[c] if slot < 0 else None))
                if slot < 0: assert not np.isnan(out[c]), "external value not ready"
            rows.append((q, r, li + lj + lk, ent))
        for s0 in range(nlevels):
            step = nlevels - 1 - s0 if upper else s0
            for q, r, lvl, ent in rows:
                if lvl != step: continue
                s = rhs[r]
                for slot, cv, xv in ent:
                    v = ytile[slot] if slot >= 0 else xv
                    assert not np.isnan(v)
                    s = s - cv * v
                if upper: s = s * invd[r]
                ytile[q] = s; out[r] = s
        flags[tile] = 1
    return out

def check(A, label):
    st, lu, dp, iud, bad = o.ilu0_factor(A)
    assert st == 0
    n = A.n; rp = A.row_ptr.astype(np.int64); col = A.col_idx.astype(np.int64); dp = dp.astype(np.int64)
    g = detect(rp, col, n)
    rhs = np.random.default_rng(1).standard_normal(n)
    ref = o.ilu0_apply(A, lu, dp.astype(np.uint64), iud, rhs)
    if g is None:
        print(label, "-> not grid-like (general path)"); return
    y = solve(rp, col, lu, dp, iud, rhs, n, g, False)
    z = solve(rp, col, lu, dp, iud, y, n, g, True)
    print(label, "grid", g, "bit-exact:", np.array_equal(z, ref))
    assert np.array_equal(z, ref)

check(o.stencil("poisson2d", 33), "poisson2d 33")
check(o.stencil("convdiff2d", 70), "convdiff2d 70")
check(o.stencil("poisson3d", 12), "poisson3d 12")
check(o.stencil("convdiff3d", 17), "convdiff3d 17")
check(o.stencil("varcoef27", 8), "varcoef27 8")
# a z-slab shard (block-Jacobi block): rows [lo,hi) with couplings outside dropped
A = o.stencil("poisson3d", 12)
lo, hi = 12 * 12 * 4, 12 * 12 * 9
S = o.submatrix(A, np.arange(lo, hi))
check(S, "poisson3d 12 slab[4:9]")
S = o.submatrix(A, np.arange(0, 12 * 12 * 2 + 30))       # ragged last plane
check(S, "poisson3d 12 ragged")
