"""Host-side model of the block-wavefront triangular-solve schedule (kryst_b200/csrc/kb_trsv_tiles.cu).

The CUDA kernel's index logic (pattern detection, tile decomposition, processing order, in-tile wavefront, which operand
comes from shared memory and which from a finished neighbour tile) is restated here in Python and executed sequentially
against the oracle's ILU(0) apply.  It checks on the CPU what makes the device schedule correct:
  * every predecessor tile precedes its dependants in the processing order (the persistent grid cannot deadlock),
  * every operand is available when it is read (no in-tile or cross-tile read-before-write),
  * per-row operation order is the oracle's, so results are bit-identical, incl. partial tiles, z-slab shards and
    ragged last planes; non-grid patterns (27-point) are rejected by the detector and keep the level-scheduled path.
The GPU parity tests of the real kernel are tests/test_gpu_ilu_gmres.py::test_trsv_schedules_bit_exact.
"""
import numpy as np
import pytest

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
        return None
    y = solve(rp, col, lu, dp, iud, rhs, n, g, False)
    z = solve(rp, col, lu, dp, iud, y, n, g, True)
    assert np.array_equal(z, ref), label
    return tuple(int(v) for v in g)


@pytest.mark.parametrize("kind,N,grid", [("poisson2d", 33, (33, 33, 1)), ("convdiff2d", 40, (40, 40, 1)),
                                         ("poisson3d", 10, (10, 10, 10)), ("convdiff3d", 9, (9, 9, 9)), ("varcoef27", 6, None)])
def test_tile_schedule_model_matches_oracle(kind, N, grid):
    assert check(o.stencil(kind, N), "%s %d" % (kind, N)) == grid


def test_tile_schedule_model_on_slab_blocks():
    A = o.stencil("poisson3d", 10)
    assert check(o.submatrix(A, np.arange(10 * 10 * 3, 10 * 10 * 8)), "slab") == (10, 10, 5)
    assert check(o.submatrix(A, np.arange(0, 10 * 10 * 2 + 17)), "ragged") == (10, 10, 3)


def test_detector_rejects_wrapped_and_irregular_patterns():
    # periodic wrap along a line (entry r-1 present at i == 0) must not pass as a box grid
    A = o.stencil("poisson2d", 8)
    d = A.to_dense()
    d[8, 7] = d[7, 8] = -1.0
    W = o.OCsr.from_dense(d)
    assert detect(W.row_ptr.astype(np.int64), W.col_idx.astype(np.int64), W.n) is None
    T = o.OCsr.from_dense(np.diag(np.full(70, 2.0)) + np.diag(np.full(69, -1.0), 1) + np.diag(np.full(69, -1.0), -1))
    assert detect(T.row_ptr.astype(np.int64), T.col_idx.astype(np.int64), T.n) is None      # 1-D: no second axis


# ---- next step: the same schedule for 9-/27-point patterns through a unimodular skew -------------------
# Executable specification for the kernel that is not written yet: cube tiles are illegal for these patterns (a lower
# neighbour (i+1, j-1, k) can sit in a later cube); in skewed coordinates u = i+j+2k, v = j+k, w = k every dependency is
# non-positive, cube tiles in (u,v,w) are convex, their predecessors are the <= 7 tiles one step back, and the in-tile
# wavefront i+2j+4k is simply lu+lv+lw.
def detect_box(rp, col, n):
    """-> (nx, ny, nz) if every off-diagonal entry is a (+-1, +-1, +-1) box-stencil neighbour, else None."""
    offs = set()
    for r in range(n):
        for p in range(rp[r], rp[r + 1]):
            if col[p] != r:
                offs.add(abs(r - col[p]))
    big = sorted(d for d in offs if d > 1)
    if not big:
        return None
    for nx in (big[0], big[0] + 1):
        if nx < 3:
            continue
        for nxy in {big[-1], big[-1] - nx - 1, big[-1] - nx, big[-1] - 1, nx * ((n + nx - 1) // nx)}:
            if nxy < nx or nxy % nx:
                continue
            ny = nxy // nx
            nz = (n + nxy - 1) // nxy
            good = True
            for r in range(n):
                i, j, k = r % nx, (r // nx) % ny, r // nxy
                for p in range(rp[r], rp[r + 1]):
                    c = col[p]
                    ci, cj, ck = c % nx, (c // nx) % ny, c // nxy
                    if max(abs(ci - i), abs(cj - j), abs(ck - k)) > 1:
                        good = False
                        break
                if not good:
                    break
            if good:
                return (int(nx), int(nz), 1) if ny == 1 else (int(nx), int(ny), int(nz))     # two-dimensional: canonical form
    return None


def solve_skewed(rp, col, lu, dptr, invd, rhs, n, grid, upper, B=4):
    nx, ny, nz = grid
    two_d = nz == 1
    def skew(i, j, k):
        return (i + j, j, 0) if two_d else (i + j + 2 * k, j + k, k)
    umax, vmax, wmax = skew(nx - 1, ny - 1, nz - 1)
    tx, ty, tz = umax // B + 1, vmax // B + 1, wmax // B + 1
    tiles = {}
    for r in range(n):
        u, v, w = skew(r % nx, (r // nx) % ny, r // (nx * ny))
        tiles.setdefault((u // B, v // B, w // B), []).append((u + v + w, r))
    order = sorted(tiles, key=lambda t: (t[0] + t[1] + t[2], t[2], t[1], t[0]))
    if upper:
        order = order[::-1]
    done, out = set(), np.full(n, np.nan)
    tile_of = {}
    for t, rows in tiles.items():
        for _, r in rows:
            tile_of[r] = t
    for t in order:
        for da, db, dc in [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1) if a + b + c]:
            p = (t[0] + da, t[1] + db, t[2] + dc) if upper else (t[0] - da, t[1] - db, t[2] - dc)
            assert p not in tiles or p in done, "predecessor tile not finished"
        for _, r in sorted(tiles[t], reverse=upper):          # in-tile wavefront lu+lv+lw
            pd = dptr[r]
            p0, p1 = (pd + 1, rp[r + 1]) if upper else (rp[r], pd)
            s = rhs[r]
            for p in range(p0, p1):
                c = col[p]
                tc = tile_of[c]
                d = tuple((tc[x] - t[x]) if upper else (t[x] - tc[x]) for x in range(3))
                assert min(d) >= 0 and max(d) <= 1, "dependency outside the 7 predecessor tiles"
                assert not np.isnan(out[c]), "operand not ready"
                s = s - lu[p] * out[c]
            out[r] = s * invd[r] if upper else s
        done.add(t)
    return out


@pytest.mark.parametrize("kind,N,grid", [("varcoef27", 7, (7, 7, 7)), ("poisson3d", 9, (9, 9, 9)), ("convdiff2d", 21, (21, 21, 1))])
def test_skewed_tile_schedule_is_legal_and_exact(kind, N, grid):
    A = o.stencil(kind, N)
    st, lu, dp, iud, bad = o.ilu0_factor(A)
    rp, col, dp = A.row_ptr.astype(np.int64), A.col_idx.astype(np.int64), dp.astype(np.int64)
    assert detect_box(rp, col, A.n) == grid
    rhs = np.random.default_rng(2).standard_normal(A.n)
    y = solve_skewed(rp, col, lu, dp, iud, rhs, A.n, grid, False)
    z = solve_skewed(rp, col, lu, dp, iud, y, A.n, grid, True)
    assert np.array_equal(z, o.ilu0_apply(A, lu, dp.astype(np.uint64), iud, rhs))
