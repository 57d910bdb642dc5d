// kb_trsv_tiles.cu — block-wavefront triangular solves for grid-structured ILU(0) factors.
//
// The level-scheduled solves (kb_ilu0.cu) pay one inter-CTA hop (~1-2 us through L2 flags) per level: 766 hops
// per solve on the 256^3 7-point operator, 2047 on the 1024^2 5-point one.  When the factor's pattern is that of a
// lexicographically numbered box grid whose lower neighbours are (i-1,j,k), (i,j-1,k), (i,j,k-1) — detected from the
// CSR pattern alone, never assumed — the rows are grouped into BX x BY x BZ tiles:
//   * a tile depends only on the (up to 7) tiles at (I-a, J-b, K-c), a,b,c in {0,1}: the tile DAG is acyclic and its
//     depth is TX+TY+TZ-2 (94 for 256^3 with 8^3 tiles, 63 for 1024^2 with 32^2 tiles) instead of the row-level depth;
//   * inside a tile ONE CTA walks the BX+BY+BZ-2 internal wavefront steps with __syncthreads (tens of ns), the
//     tile's solution living in shared memory; values from finished neighbour tiles are read once from L2;
//   * tiles are processed by a persistent co-resident grid in block-level order; a per-tile flag (release/acquire)
//     is the only inter-CTA communication.  No sentinel fill, no value polling.
// Every row is still computed as s = rhs - sum_{stored entries in ascending column order} l_ij * y_j (then * 1/u_ii
// for U): identical operation order to the oracle, so the result is bit-identical whatever the schedule.
// Patterns that do not pass the check (27-point stencils, general matrices) keep the level-scheduled kernels.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kb_objects.h"

struct KbTileSolve {
    int nx = 0, ny = 0, nz = 0;          // detected grid
    int bx = 0, by = 0, bz = 0;          // tile shape
    int tx = 0, ty = 0, tz = 0;          // tiles per dimension
    int ntiles = 0, rows_per_thread = 0, grid[2] = {0, 0};
    int* order[2] = {nullptr, nullptr};  // tile ids in processing order (lower: ascending block level; upper: descending)
    int* flags = nullptr;                // [2][ntiles] completion flags, cleared before every apply
    unsigned* err = nullptr;             // borrowed: the preconditioner's error word (set when a spin times out)
    // packet mailbox of the fine-grained variant (kb_trsv_tiles_ll): per tile three incoming faces of 16-byte packets
    ulonglong2* mail = nullptr;
    unsigned* sync = nullptr;            // [0],[1] epoch of L / U ; [2],[3] finish tickets
    int face_len = 0;                    // packets per tile: by*bz + bx*bz + bx*by
    int variant = 0;                     // 0 flag hand-off, 1 packets + communication warp (fine-grained), 2 packets at tile granularity
    unsigned long long* trace = nullptr; // diagnostics (KB_TILES_TRACE=1)
};

// ---- pattern detection ---------------------------------------------------------------------------------------
// pass 1: smallest and largest |r - c| > 1 over all off-diagonal entries
__global__ void k_tiles_offsets(const int* __restrict__ rp, const int* __restrict__ col, int n, int* __restrict__ smin, int* __restrict__ smax) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int lo = 0x7fffffff, hi = 0;
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
        const int d = abs(r - col[p]);
        if (d > 1) { lo = min(lo, d); hi = max(hi, d); }
    }
    if (hi > 0) { atomicMin(smin, lo); atomicMax(smax, hi); }
}
// pass 2: every off-diagonal entry is a +-1 step along exactly one grid axis and does not wrap around a line / plane
__global__ void k_tiles_verify(const int* __restrict__ rp, const int* __restrict__ col, int n, int nx, int ny, int sx, int sy,
                               int* __restrict__ bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int i = r % nx, j = (r / nx) % ny;
    const bool three_d = sy > sx;
    bool ok = true;
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
        const int c = col[p];
        if (c == r) continue;
        const int d = abs(r - c);
        if (c < 0 || c >= n) ok = false;
        else if (d == 1) ok = ok && (c < r ? i >= 1 : i + 1 < nx);
        else if (d == sx) ok = ok && (!three_d || (c < r ? j >= 1 : j + 1 < ny));
        else if (three_d && d == sy) ok = ok && true;
        else ok = false;
    }
    if (!ok) atomicExch(bad, 1);
}

// ---- the solve -------------------------------------------------------------------------------------------------
struct KbTileArgs {
    const int* __restrict__ rp; const int* __restrict__ col; const double* __restrict__ lu; const int* __restrict__ dptr;
    const double* __restrict__ inv_diag;
    const double* __restrict__ rhs; double* out;
    int n, nx, ny, nz, bx, by, bz, tx, ty, tz, ntiles;
    const int* __restrict__ order; int* flags; unsigned* err;
    const KbCtl* skip_ctl; int skip_mask;
    ulonglong2* mail; unsigned* sync; int face_len; int lag;      // kb_trsv_tiles_ll
    unsigned long long* trace;                                     // diagnostics: per tile {pick-up, preload done, distance wait done, end} ns
};

#define KB_TILE_ROWS_MAX 1024

template <bool UPPER, int R>
__global__ void __launch_bounds__(KB_THREADS) kb_trsv_tiles(KbTileArgs a) {
    if (kb_skip(a.skip_ctl, a.skip_mask)) return;
    __shared__ double ytile[KB_TILE_ROWS_MAX];
    const int tid = threadIdx.x;
    const int bxy = a.bx * a.by, trows = bxy * a.bz;
    const int nlevels = a.bx + a.by + a.bz - 2;          // internal wavefront depth of a full tile
    const int sx = a.nx;                                 // column stride of a j-step; anything else that is not 1 is a k-step
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const int tile = a.order[t];
        const int TI = tile % a.tx, TJ = (tile / a.tx) % a.ty, TK = tile / (a.tx * a.ty);
        // ---- everything that does not depend on other tiles: this thread's rows and their stored entries
        int row[R], lvl[R], di[R][3];
        double rh[R], dg[R], cv[R][3], xv[R][3];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int q = m * KB_THREADS + tid;
            row[m] = -1; lvl[m] = -1; rh[m] = 0.0; dg[m] = 1.0;
#pragma unroll
            for (int e = 0; e < 3; ++e) { di[m][e] = -2; cv[m][e] = 0.0; xv[m][e] = 0.0; }     // -2: no entry
            if (q < trows) {
                const int li = q % a.bx, lj = (q / a.bx) % a.by, lk = q / bxy;
                const int gi = TI * a.bx + li, gj = TJ * a.by + lj, gk = TK * a.bz + lk;
                const int r = gi + a.nx * (gj + a.ny * gk);
                if (gi < a.nx && gj < a.ny && gk < a.nz && r < a.n) {
                    row[m] = r;
                    lvl[m] = li + lj + lk;
                    rh[m] = a.rhs[r];
                    const int pd = a.dptr[r];
                    const int p0 = UPPER ? pd + 1 : a.rp[r];
                    const int p1 = UPPER ? a.rp[r + 1] : pd;
                    if (UPPER) dg[m] = a.inv_diag[r];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        if (p0 + e < p1) {
                            const int c = a.col[p0 + e];
                            cv[m][e] = a.lu[p0 + e];
                            const int d = UPPER ? c - r : r - c;
                            // inside the tile -> shared-memory slot, otherwise -1 (value fetched after the flag wait)
                            int slot = -1;
                            if (d == 1) { if (UPPER ? li + 1 < a.bx : li >= 1) slot = UPPER ? q + 1 : q - 1; }
                            else if (d == sx) { if (UPPER ? lj + 1 < a.by : lj >= 1) slot = UPPER ? q + a.bx : q - a.bx; }
                            else { if (UPPER ? lk + 1 < a.bz : lk >= 1) slot = UPPER ? q + bxy : q - bxy; }
                            di[m][e] = slot;
                            if (slot < 0) xv[m][e] = __longlong_as_double((long long)c);      // park the column id until the wait is over
                        }
                    }
                }
            }
        }
        // ---- wait for the (up to 7) predecessor tiles
        if (tid >= 1 && tid < 8) {
            const int da = tid & 1, db = (tid >> 1) & 1, dc = (tid >> 2) & 1;
            const int PI = UPPER ? TI + da : TI - da, PJ = UPPER ? TJ + db : TJ - db, PK = UPPER ? TK + dc : TK - dc;
            if (PI >= 0 && PI < a.tx && PJ >= 0 && PJ < a.ty && PK >= 0 && PK < a.tz) {
                const int* f = a.flags + (PI + a.tx * (PJ + a.ty * PK));
                unsigned spins = 0;
                int v;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                    if (v == 0 && a.lag > 0) __nanosleep(a.lag);   // polite polling (KB_TILES_SLEEP ns): fewer strong loads competing with the working tiles
                    if (v == 0 && (++spins & 1023u) == 0u) {      // bounded: give up (and let everybody give up) instead of hanging
                        if (spins > (1u << 26)) atomicExch(a.err, 1u);
                        if (*reinterpret_cast<volatile unsigned*>(a.err)) break;
                    }
                } while (v == 0);
            }
        }
        __syncthreads();
        // ---- values owned by finished tiles (L2 reads; never cached in L1)
#pragma unroll
        for (int m = 0; m < R; ++m)
#pragma unroll
            for (int e = 0; e < 3; ++e)
                if (di[m][e] == -1) xv[m][e] = __ldcg(a.out + (int)__double_as_longlong(xv[m][e]));
        // ---- internal wavefront
        for (int s0 = 0; s0 < nlevels; ++s0) {
            const int step = UPPER ? nlevels - 1 - s0 : s0;
#pragma unroll
            for (int m = 0; m < R; ++m) {
                if (lvl[m] == step) {
                    // the three operand loads are unconditional (clamped slot) and independent, so they overlap; an
                    // absent entry contributes s - 0.0*0.0, which leaves every finite s (and -0.0) unchanged
                    double v[3];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double y = ytile[max(di[m][e], 0)];
                        v[e] = di[m][e] >= 0 ? y : xv[m][e];
                    }
                    double s = rh[m];
#pragma unroll
                    for (int e = 0; e < 3; ++e) s = s - cv[m][e] * v[e];
                    if (UPPER) s = s * dg[m];
                    ytile[m * KB_THREADS + tid] = s;
                    a.out[row[m]] = s;
                }
            }
            __syncthreads();
        }
        // ---- publish: the barrier above ordered every thread's stores before this release
        if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.flags + tile), "r"(1) : "memory");
    }
}

// ---- fine-grained variant: faces travel as packets, a tile starts as soon as its FIRST rows can ---------------------
// kb_trsv_tiles hands a tile to its successors with one release/acquire flag after its last in-tile step: a tile level
// costs the hand-off (~3.6 us: release fence, flag poll, L2 reads of the face) plus all bx+by+bz-2 in-tile steps, 7.7 us
// on 256^3.  But a successor's first rows only need the predecessor's FACE rows, which are finished bx-1 (by-1, bz-1)
// steps after the predecessor's first row.  Here every face value is published the moment it is computed, as a 16-byte
// packet {lo32|tag, hi32|tag} in the consumer tile's mailbox (the data is its own flag: no release fence, no flag
// array, no per-apply memset - the tag is the apply's epoch), and the consumer polls exactly the packets the rows of
// its current in-tile step need.  Packets are loaded without blocking at tile start and again two steps ahead, so after
// the first one or two blocking polls a tile runs a fixed distance (~12 steps, 2.3 us) behind its predecessors instead
// of a whole tile level.  Per-row arithmetic and its order are unchanged: same bits.
__device__ __forceinline__ ulonglong2 kb_tl_load(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void kb_tl_store(ulonglong2* p, double val, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(val), t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ __forceinline__ bool kb_tl_ok(const ulonglong2& v, unsigned tag) { return (unsigned)(v.x >> 32) == tag && (unsigned)(v.y >> 32) == tag; }
__device__ __forceinline__ double kb_tl_value(const ulonglong2& v) { return __longlong_as_double((long long)((v.x & 0xffffffffull) | (v.y << 32))); }

// Implementation: the 8 compute warps run EXACTLY the step of kb_trsv_tiles (three unconditional shared-memory operand
// loads, the row sum, one store) - a first version that polled packets inside that step needed ~500 instructions per
// step and ran 3x slower than the flag kernel.  All packet traffic belongs to a 9th "communication" warp:
//   * lane l owns one line of one incoming face (A: fixed lj, B / C: fixed li) and walks it as the in-tile wavefront
//     advances; while the compute warps work on step s it makes sure the packets of step s+1 are in the shared-memory
//     halo (ytile[1024 + ...], where the compute rows find them like any in-tile operand), prefetching three steps
//     further through a small shared-memory ring, and publishes the face rows finished in step s-1;
//   * before step 0 it waits for the packets of step `lag`: a tile that starts the moment it can runs at exactly its
//     producers' pace, finds every packet "not yet there" when it looks ahead and pays a blocking L2 poll per step.
// One __syncthreads per step orders halo writes before their use, as it orders the in-tile values.
#define KB_TL_HALO_BASE KB_TILE_ROWS_MAX
#define KB_TL_HALO_MAX 512
#define KB_TL_THREADS (KB_THREADS + 32)
#define KB_TL_PD 3                     // packets are requested this many steps before the step that consumes them

template <bool UPPER, int R>
__global__ void __launch_bounds__(KB_TL_THREADS) kb_trsv_tiles_ll(KbTileArgs a) {
    if (kb_skip(a.skip_ctl, a.skip_mask)) return;
    __shared__ double ytile[KB_TL_HALO_BASE + KB_TL_HALO_MAX];
    __shared__ ulonglong2 pring[4][32];
    const int tid = threadIdx.x;
    const int bxy = a.bx * a.by, trows = bxy * a.bz;
    const int nlevels = a.bx + a.by + a.bz - 2;          // internal wavefront depth of a full tile
    const int sx = a.nx;                                 // column stride of a j-step; anything else that is not 1 is a k-step
    const int offB = a.by * a.bz, offC = offB + a.bx * a.bz;      // face offsets inside a tile's mailbox / the halo: A | B | C
    // L and U share the mailbox: their tags never coincide (even / odd), and both advance once per apply
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(a.sync + (UPPER ? 1 : 0)) + 1u;
    const unsigned tag = 2u * epoch + (UPPER ? 1u : 0u);
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const int tile = a.order[t];
        const int TI = tile % a.tx, TJ = (tile / a.tx) % a.ty, TK = tile / (a.tx * a.ty);
        unsigned long long tr0 = 0ull;
        if (a.trace && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr0));
        if (tid >= KB_THREADS) {
            // ================= communication warp =================
            const int l = tid - KB_THREADS;
            // this lane's incoming line: face (0 A, 1 B, 2 C), fixed mirrored coordinate u, extents of the walk
            int face = -1, u = 0, wext = 0;
            if (l < a.by) { face = 0; u = l; wext = a.bz; }
            else if (l < a.by + a.bx) { face = 1; u = l - a.by; wext = a.bz; }
            else if (l < a.by + 2 * a.bx) { face = 2; u = l - a.by - a.bx; wext = a.by; }
            // does the neighbour tile on the incoming / outgoing side exist?
            const bool has_in = face == 0 ? (UPPER ? TI + 1 < a.tx : TI > 0) : face == 1 ? (UPPER ? TJ + 1 < a.ty : TJ > 0) : face == 2 ? (UPPER ? TK + 1 < a.tz : TK > 0) : false;
            const bool has_out = face == 0 ? (UPPER ? TI > 0 : TI + 1 < a.tx) : face == 1 ? (UPPER ? TJ > 0 : TJ + 1 < a.ty) : face == 2 ? (UPPER ? TK > 0 : TK + 1 < a.tz) : false;
            const int in_tile_off = face == 0 ? (UPPER ? 1 : -1) : face == 1 ? (UPPER ? a.tx : -a.tx) : (UPPER ? a.tx * a.ty : -a.tx * a.ty);
            const ulonglong2* in_mail = a.mail + (size_t)tile * a.face_len;
            ulonglong2* out_mail = a.mail + (size_t)(tile - in_tile_off) * a.face_len;       // the tile on the other side consumes what this one produces
            const int foff = face == 0 ? 0 : face == 1 ? offB : offC;
            const int fdepth = face == 0 ? a.bx : face == 1 ? a.by : a.bz;              // extent across the face: outgoing rows lag by fdepth-1 steps
            // (li,lj,lk) of the line's element w on the incoming (in=true) or outgoing face, in actual coordinates
            auto coords = [&](int w, bool in, int& li, int& lj, int& lk) {
                const int uu = UPPER ? -1 : 0;      // marker only (keeps the lambda's intent readable)
                (void)uu;
                if (face == 0) { lj = UPPER ? a.by - 1 - u : u; lk = UPPER ? a.bz - 1 - w : w; li = (in != UPPER) ? 0 : a.bx - 1; }
                else if (face == 1) { li = UPPER ? a.bx - 1 - u : u; lk = UPPER ? a.bz - 1 - w : w; lj = (in != UPPER) ? 0 : a.by - 1; }
                else { li = UPPER ? a.bx - 1 - u : u; lj = UPPER ? a.by - 1 - w : w; lk = (in != UPPER) ? 0 : a.bz - 1; }
            };
            auto face_index = [&](int li, int lj, int lk) -> int { return face == 0 ? lj + a.by * lk : face == 1 ? li + a.bx * lk : li + a.bx * lj; };
            // incoming packet of in-tile step s for this lane: slot in the face (-1: none)
            auto in_slot = [&](int s) -> int {
                const int w = s - u;
                if (face < 0 || !has_in || w < 0 || w >= wext) return -1;
                int li, lj, lk;
                coords(w, true, li, lj, lk);
                const int gi = TI * a.bx + li, gj = TJ * a.by + lj, gk = TK * a.bz + lk;
                if (gi >= a.nx || gj >= a.ny || gk >= a.nz) return -1;
                const long long r = (long long)gi + (long long)a.nx * (gj + (long long)a.ny * gk);
                if (r >= a.n) return -1;
                // (U solve on a ragged grid: the row across the face is row r + stride and may not exist)
                if (UPPER && r + (face == 0 ? 1 : face == 1 ? a.nx : (long long)a.nx * a.ny) >= a.n) return -1;
                return foff + face_index(li, lj, lk);
            };
            auto fetch = [&](int slot, ulonglong2 q) {        // blocking: until the packet carries this apply's tag
                unsigned spins = 0;
                while (!kb_tl_ok(q, tag)) {
                    if (++spins > 8u) __nanosleep(40);
                    q = kb_tl_load(in_mail + slot);
                    if ((spins & 1023u) == 0u) {
                        if (spins > (1u << 22)) atomicExch(a.err, 1u);
                        if (*reinterpret_cast<volatile unsigned*>(a.err)) break;
                    }
                }
                ytile[KB_TL_HALO_BASE + slot] = kb_tl_value(q);
            };
            // prologue: request the packets of steps 0..PD, keep the distance, deliver step 0
            // (requests are cp.async copies straight into the shared-memory ring: a plain load would have to return
            //  before its value can be parked anywhere, i.e. one exposed L2 round trip per step)
            auto request = [&](int s) {
                const int sl = s < nlevels ? in_slot(s) : -1;
                if (sl >= 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&pring[s & 3][l])), "l"(in_mail + sl) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
#pragma unroll
            for (int s = 0; s <= KB_TL_PD; ++s) request(s);
            {
                const int ls = min(a.lag & 0xff, nlevels - 1);
                const int sl = in_slot(ls);
                if (sl >= 0 && ls > KB_TL_PD) { ulonglong2 q = kb_tl_load(in_mail + sl); unsigned spins = 0; while (!kb_tl_ok(q, tag)) { __nanosleep(100); q = kb_tl_load(in_mail + sl); if (++spins > (1u << 22)) { atomicExch(a.err, 1u); break; } if ((spins & 255u) == 0u && *reinterpret_cast<volatile unsigned*>(a.err)) break; } }
                const int s0 = in_slot(0);
                asm volatile("cp.async.wait_group %0;" ::"n"(KB_TL_PD) : "memory");
                if (s0 >= 0) fetch(s0, pring[0][l]);
            }
            __syncthreads();                                  // halo of step 0 is in place
            for (int s = 0; s < nlevels; ++s) {
                // publish the face rows solved in step s-1 (their values are in ytile since the last barrier)
                if (s > 0 && face >= 0 && has_out) {
                    const int w = (s - 1) - (fdepth - 1) - u;
                    if (w >= 0 && w < wext) {
                        int li, lj, lk;
                        coords(w, false, li, lj, lk);
                        const int gi = TI * a.bx + li, gj = TJ * a.by + lj, gk = TK * a.bz + lk;
                        const long long r = (long long)gi + (long long)a.nx * (gj + (long long)a.ny * gk);
                        if (gi < a.nx && gj < a.ny && gk < a.nz && r < a.n) kb_tl_store(out_mail + foff + face_index(li, lj, lk), ytile[li + a.bx * (lj + a.by * lk)], tag);
                    }
                }
                // the halo of step s+1, then the request for step s+1+PD
                asm volatile("cp.async.wait_group %0;" ::"n"(KB_TL_PD - 1) : "memory");
                if (s + 1 < nlevels) { const int sl = in_slot(s + 1); if (sl >= 0) fetch(sl, pring[(s + 1) & 3][l]); }
                request(s + 1 + KB_TL_PD);
                __syncthreads();                              // end of step s
            }
            // the rows of the last step
            if (face >= 0 && has_out) {
                const int w = (nlevels - 1) - (fdepth - 1) - u;
                if (w >= 0 && w < wext) {
                    int li, lj, lk;
                    coords(w, false, li, lj, lk);
                    const int gi = TI * a.bx + li, gj = TJ * a.by + lj, gk = TK * a.bz + lk;
                    const long long r = (long long)gi + (long long)a.nx * (gj + (long long)a.ny * gk);
                    if (gi < a.nx && gj < a.ny && gk < a.nz && r < a.n) kb_tl_store(out_mail + foff + face_index(li, lj, lk), ytile[li + a.bx * (lj + a.by * lk)], tag);
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                  // ytile may be overwritten by the next tile now
            continue;
        }
        // ================= compute warps: the step of kb_trsv_tiles =================
        int row[R], lvl[R], di[R][3];
        double rh[R], dg[R], cv[R][3];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int q = m * KB_THREADS + tid;
            row[m] = -1; lvl[m] = -1; rh[m] = 0.0; dg[m] = 1.0;
#pragma unroll
            for (int e = 0; e < 3; ++e) { di[m][e] = -2; cv[m][e] = 0.0; }     // -2: no entry
            if (q < trows) {
                const int li = q % a.bx, lj = (q / a.bx) % a.by, lk = q / bxy;
                const int gi = TI * a.bx + li, gj = TJ * a.by + lj, gk = TK * a.bz + lk;
                const int r = gi + a.nx * (gj + a.ny * gk);
                if (gi < a.nx && gj < a.ny && gk < a.nz && r < a.n) {
                    row[m] = r;
                    lvl[m] = li + lj + lk;
                    rh[m] = a.rhs[r];
                    const int pd = a.dptr[r];
                    const int p0 = UPPER ? pd + 1 : a.rp[r];
                    const int p1 = UPPER ? a.rp[r + 1] : pd;
                    if (UPPER) dg[m] = a.inv_diag[r];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        if (p0 + e < p1) {
                            const int c = a.col[p0 + e];
                            cv[m][e] = a.lu[p0 + e];
                            const int d = UPPER ? c - r : r - c;
                            // inside the tile -> its shared-memory slot, otherwise the halo slot the communication warp fills
                            int slot;
                            if (d == 1) slot = (UPPER ? li + 1 < a.bx : li >= 1) ? (UPPER ? q + 1 : q - 1) : KB_TL_HALO_BASE + (lj + a.by * lk);
                            else if (d == sx) slot = (UPPER ? lj + 1 < a.by : lj >= 1) ? (UPPER ? q + a.bx : q - a.bx) : KB_TL_HALO_BASE + offB + (li + a.bx * lk);
                            else slot = (UPPER ? lk + 1 < a.bz : lk >= 1) ? (UPPER ? q + bxy : q - bxy) : KB_TL_HALO_BASE + offC + (li + a.bx * lj);
                            di[m][e] = slot;
                        }
                    }
                }
            }
        }
        __syncthreads();                                      // halo of step 0 is in place
        for (int s0 = 0; s0 < nlevels; ++s0) {
            const int step = UPPER ? nlevels - 1 - s0 : s0;
#pragma unroll
            for (int m = 0; m < R; ++m) {
                if (lvl[m] == step) {
                    // the three operand loads are unconditional (clamped slot) and independent, so they overlap; an
                    // absent entry contributes s - 0.0*0.0, which leaves every finite s (and -0.0) unchanged
                    double v[3];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double y = ytile[max(di[m][e], 0)];
                        v[e] = di[m][e] >= 0 ? y : 0.0;
                    }
                    double s = rh[m];
#pragma unroll
                    for (int e = 0; e < 3; ++e) s = s - cv[m][e] * v[e];
                    if (UPPER) s = s * dg[m];
                    ytile[m * KB_THREADS + tid] = s;
                    a.out[row[m]] = s;
                }
            }
            __syncthreads();
        }
        __syncthreads();                                      // (pairs with the communication warp's last barrier)
        if (a.trace && tid == 0) {
            unsigned long long tr3;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr3));
            unsigned long long* q = a.trace + ((size_t)(UPPER ? a.ntiles : 0) + tile) * 4;
            q[0] = tr0; q[1] = tr0; q[2] = tr0; q[3] = tr3;
        }
    }
    // the last CTA to finish publishes the epoch: every CTA has read it by then
    if (tid == 0) {
        __threadfence();
        const unsigned tk = atomicAdd(a.sync + 2 + (UPPER ? 1 : 0), 1u);
        if (tk == gridDim.x - 1u) {
            a.sync[2 + (UPPER ? 1 : 0)] = 0u;
            a.sync[UPPER ? 1 : 0] = epoch;
            __threadfence();
        }
    }
}

// ---- packet faces, tile-granular ("kb_trsv_tiles_pk") ------------------------------------------------------------
// Same tile-level pipeline as kb_trsv_tiles, but the hand-off is made of the data itself: every face row publishes
// its value as a 16-byte tagged packet the moment it is solved, and a successor tile starts by polling, thread by
// thread, exactly the packets its own rows read across a face.  Compared with the flag hand-off this removes the
// release fence after the last step, the flag poll + barrier, the dependent L2 read of the neighbours' values and the
// per-apply flag memset; the in-tile step stays the lean one (operands that cross a face sit in registers).
template <bool UPPER, int R>
__global__ void __launch_bounds__(KB_THREADS, R <= 2 ? 4 : 1) kb_trsv_tiles_pk(KbTileArgs a) {
    if (kb_skip(a.skip_ctl, a.skip_mask)) return;
    __shared__ double ytile[KB_TILE_ROWS_MAX];
    const int tid = threadIdx.x;
    const int bxy = a.bx * a.by, trows = bxy * a.bz;
    const int nlevels = a.bx + a.by + a.bz - 2;
    const int sx = a.nx;
    const int offB = a.by * a.bz, offC = offB + a.bx * a.bz;
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(a.sync + (UPPER ? 1 : 0)) + 1u;
    const unsigned tag = 2u * epoch + (UPPER ? 1u : 0u);
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const int tile = a.order[t];
        const int TI = tile % a.tx, TJ = (tile / a.tx) % a.ty, TK = tile / (a.tx * a.ty);
        const int my_mail = tile * a.face_len;
        int row[R], lvl[R], di[R][3], ext[R][3], outp[R][3];
        double rh[R], dg[R], cv[R][3], xv[R][3];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int q = m * KB_THREADS + tid;
            row[m] = -1; lvl[m] = -1; rh[m] = 0.0; dg[m] = 1.0;
#pragma unroll
            for (int e = 0; e < 3; ++e) { di[m][e] = -2; cv[m][e] = 0.0; xv[m][e] = 0.0; ext[m][e] = -1; outp[m][e] = -1; }
            if (q < trows) {
                const int li = q % a.bx, lj = (q / a.bx) % a.by, lk = q / bxy;
                const int gi = TI * a.bx + li, gj = TJ * a.by + lj, gk = TK * a.bz + lk;
                const int r = gi + a.nx * (gj + a.ny * gk);
                if (gi < a.nx && gj < a.ny && gk < a.nz && r < a.n) {
                    row[m] = r;
                    lvl[m] = li + lj + lk;
                    rh[m] = a.rhs[r];
                    const int pd = a.dptr[r];
                    const int p0 = UPPER ? pd + 1 : a.rp[r];
                    const int p1 = UPPER ? a.rp[r + 1] : pd;
                    if (UPPER) dg[m] = a.inv_diag[r];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        if (p0 + e < p1) {
                            const int c = a.col[p0 + e];
                            cv[m][e] = a.lu[p0 + e];
                            const int d = UPPER ? c - r : r - c;
                            int slot = -1;
                            if (d == 1) { if (UPPER ? li + 1 < a.bx : li >= 1) slot = UPPER ? q + 1 : q - 1; else ext[m][e] = my_mail + (lj + a.by * lk); }
                            else if (d == sx) { if (UPPER ? lj + 1 < a.by : lj >= 1) slot = UPPER ? q + a.bx : q - a.bx; else ext[m][e] = my_mail + offB + (li + a.bx * lk); }
                            else { if (UPPER ? lk + 1 < a.bz : lk >= 1) slot = UPPER ? q + bxy : q - bxy; else ext[m][e] = my_mail + offC + (li + a.bx * lj); }
                            di[m][e] = slot;
                        }
                    }
                    if (!UPPER) {
                        if (li == a.bx - 1 && gi + 1 < a.nx) outp[m][0] = (tile + 1) * a.face_len + (lj + a.by * lk);
                        if (lj == a.by - 1 && gj + 1 < a.ny) outp[m][1] = (tile + a.tx) * a.face_len + offB + (li + a.bx * lk);
                        if (lk == a.bz - 1 && gk + 1 < a.nz) outp[m][2] = (tile + a.tx * a.ty) * a.face_len + offC + (li + a.bx * lj);
                    } else {
                        if (li == 0 && gi >= 1) outp[m][0] = (tile - 1) * a.face_len + (lj + a.by * lk);
                        if (lj == 0 && gj >= 1) outp[m][1] = (tile - a.tx) * a.face_len + offB + (li + a.bx * lk);
                        if (lk == 0 && gk >= 1) outp[m][2] = (tile - a.tx * a.ty) * a.face_len + offC + (li + a.bx * lj);
                    }
                }
            }
        }
        // ---- the hand-off: every thread waits for the packets its own rows read across a face.  All of a thread's
        //      packets are requested before the first one is looked at (one L2 round trip, not one per packet).
        {
            ulonglong2 qv[R][3];
#pragma unroll
            for (int m = 0; m < R; ++m)
#pragma unroll
                for (int e = 0; e < 3; ++e) { qv[m][e] = make_ulonglong2(0ull, 0ull); if (ext[m][e] >= 0) qv[m][e] = kb_tl_load(a.mail + ext[m][e]); }
            unsigned spins = 0;
            bool pending = true;
            while (pending) {
                pending = false;
#pragma unroll
                for (int m = 0; m < R; ++m)
#pragma unroll
                    for (int e = 0; e < 3; ++e)
                        if (ext[m][e] >= 0 && !kb_tl_ok(qv[m][e], tag)) { qv[m][e] = kb_tl_load(a.mail + ext[m][e]); pending = true; }
                if (pending) {
                    if (++spins > 2u) __nanosleep(64);
                    if ((spins & 1023u) == 0u) {
                        if (spins > (1u << 22)) atomicExch(a.err, 1u);
                        if (*reinterpret_cast<volatile unsigned*>(a.err)) break;
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < R; ++m)
#pragma unroll
                for (int e = 0; e < 3; ++e) if (ext[m][e] >= 0) xv[m][e] = kb_tl_value(qv[m][e]);
        }
        __syncthreads();          // (the previous tile's readers of ytile are done as well)
        // ---- internal wavefront: the step of kb_trsv_tiles plus the face packets
        for (int s0 = 0; s0 < nlevels; ++s0) {
            const int step = UPPER ? nlevels - 1 - s0 : s0;
#pragma unroll
            for (int m = 0; m < R; ++m) {
                if (lvl[m] == step) {
                    double v[3];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double y = ytile[max(di[m][e], 0)];
                        v[e] = di[m][e] >= 0 ? y : xv[m][e];
                    }
                    double s = rh[m];
#pragma unroll
                    for (int e = 0; e < 3; ++e) s = s - cv[m][e] * v[e];
                    if (UPPER) s = s * dg[m];
                    ytile[m * KB_THREADS + tid] = s;
                    a.out[row[m]] = s;
                    if ((outp[m][0] & outp[m][1] & outp[m][2]) != -1) {      // a face row (rare): publish
#pragma unroll
                        for (int e = 0; e < 3; ++e) if (outp[m][e] >= 0) kb_tl_store(a.mail + outp[m][e], s, tag);
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        __threadfence();
        const unsigned tk = atomicAdd(a.sync + 2 + (UPPER ? 1 : 0), 1u);
        if (tk == gridDim.x - 1u) {
            a.sync[2 + (UPPER ? 1 : 0)] = 0u;
            a.sync[UPPER ? 1 : 0] = epoch;
            __threadfence();
        }
    }
}

// ---- host ------------------------------------------------------------------------------------------------------
void kb_tiles_grid(const KbTileSolve* t, int* nx, int* ny, int* nz) { *nx = t->nx; *ny = t->ny; *nz = t->nz; }
void kb_tiles_free(KbTileSolve* t) {
    if (!t) return;
    KB_FREE(t->order[0]); KB_FREE(t->order[1]); KB_FREE(t->flags); KB_FREE(t->mail); KB_FREE(t->sync); KB_FREE(t->trace);
    delete t;
}

template <bool UPPER, int R>
static int tiles_occupancy(int* occ, int variant = 0) {
    if (variant == 2) KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kb_trsv_tiles_pk<UPPER, R>, KB_THREADS, 0));
    else if (variant == 1) KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kb_trsv_tiles_ll<UPPER, R>, KB_TL_THREADS, 0));
    else KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kb_trsv_tiles<UPPER, R>, KB_THREADS, 0));
    return KB_OK;
}

// *out stays nullptr (and KB_OK is returned) when the pattern is not a 5-/7-point box grid.
int kb_tiles_build(kb_pc_s* pc, unsigned* d_err, KbTileSolve** out) {
    *out = nullptr;
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    const int n = (int)A->n;
    if (n < 64) return KB_OK;
    const unsigned nb = (unsigned)((n + 255) / 256);
    int* d_w = nullptr;     // [0] smin, [1] smax, [2] bad
    KB_TRY(kb_alloc(&d_w, 4));
    int h_w[3] = {0x7fffffff, 0, 0};
    KB_CUDA(cudaMemcpyAsync(d_w, h_w, sizeof(h_w), cudaMemcpyHostToDevice, c->stream));
    { KbLaunch L(c, KB_K_OTHER); k_tiles_offsets<<<nb, 256, 0, c->stream>>>(pc->l_rp, pc->l_col, n, d_w, d_w + 1); }
    KB_CUDA(cudaMemcpyAsync(h_w, d_w, sizeof(h_w), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    const int s1 = h_w[0], s2 = h_w[1];
    if (s2 == 0 || s1 < 2 || (s2 != s1 && s2 % s1 != 0)) { cudaFree(d_w); return KB_OK; }
    const int nx = s1, sy = s2 == s1 ? s1 : s2;                      // sy == nx: two-dimensional
    const int ny = s2 == s1 ? (n + nx - 1) / nx : s2 / s1;
    const int nz = s2 == s1 ? 1 : (n + s2 - 1) / s2;
    { KbLaunch L(c, KB_K_OTHER); k_tiles_verify<<<nb, 256, 0, c->stream>>>(pc->l_rp, pc->l_col, n, nx, ny, nx, sy, d_w + 2); }
    KB_CUDA(cudaMemcpyAsync(h_w, d_w, sizeof(h_w), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_w);
    if (h_w[2]) return KB_OK;

    KbTileSolve* t = new KbTileSolve;
    t->nx = nx; t->ny = ny; t->nz = nz;
    if (nz == 1) { t->bx = 32; t->by = 32; t->bz = 1; }
    else { t->bx = 8; t->by = 8; t->bz = 8; }
    if (const char* e = getenv("KB_TILES_SHAPE")) {      // tuning knob: "bx,by,bz"
        int b[3] = {0, 0, 0};
        if (sscanf(e, "%d,%d,%d", &b[0], &b[1], &b[2]) == 3 && b[0] > 0 && b[1] > 0 && b[2] > 0 && (long long)b[0] * b[1] * b[2] <= KB_TILE_ROWS_MAX) {
            t->bx = b[0]; t->by = b[1]; t->bz = nz == 1 ? 1 : b[2];
        }
    }
    { const int rows = t->bx * t->by * t->bz; t->rows_per_thread = rows <= KB_THREADS ? 1 : rows <= 2 * KB_THREADS ? 2 : 4; }
    t->tx = (nx + t->bx - 1) / t->bx; t->ty = (ny + t->by - 1) / t->by; t->tz = (nz + t->bz - 1) / t->bz;
    const long long nt = (long long)t->tx * t->ty * t->tz;
    if (nt > (1 << 24)) { delete t; return KB_OK; }
    t->ntiles = (int)nt;
    // processing order: stable sort by block level I+J+K (ascending for L, descending for U)
    std::vector<int> ord((size_t)nt), lev((size_t)nt);
    for (int k = 0, id = 0; k < t->tz; ++k)
        for (int j = 0; j < t->ty; ++j)
            for (int i = 0; i < t->tx; ++i, ++id) { ord[id] = id; lev[id] = i + j + k; }
    std::stable_sort(ord.begin(), ord.end(), [&](int p, int q) { return lev[p] < lev[q]; });
    int st = KB_OK;
    do {
        if ((st = kb_alloc(&t->order[0], (size_t)nt)) != KB_OK || (st = kb_alloc(&t->order[1], (size_t)nt)) != KB_OK ||
            (st = kb_alloc(&t->flags, 2 * (size_t)nt)) != KB_OK) break;
        t->err = d_err;
        if (cudaMemcpyAsync(t->order[0], ord.data(), (size_t)nt * sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        cudaStreamSynchronize(c->stream);
        std::reverse(ord.begin(), ord.end());
        if (cudaMemcpyAsync(t->order[1], ord.data(), (size_t)nt * sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        // fine-grained packet variant (3-D grids; KB_TILES_LL=0 keeps the flag hand-off)
        // Measured on B200 (256^3, per solve): release/acquire flag hand-off 0.71 ms.  Packet faces (the data is its own
        // flag): 1.92 ms with the polling inside the compute step (~500 instructions per step), 1.21 ms with a
        // communication warp (fine-grained: a tile starts as soon as its first rows can), 1.09-1.29 ms with packets at tile
        // granularity - every variant has ~200 polling threads per waiting tile where the flag kernel has 7, and the
        // strong loads of ~400 waiting tiles slow the L2 for everybody.  The flag hand-off stays the default
        // (KB_TILES_LL=1 / 2 select the packet variants; both are bit-identical and covered by the GPU tests).
        const int ll_env = getenv("KB_TILES_LL") ? atoi(getenv("KB_TILES_LL")) : 0;      // 0 flags, 1 fine-grained (communication warp), 2 packets at tile granularity
        t->variant = nz > 1 ? ll_env : 0;
        if (t->variant == 1 && !(t->by + 2 * t->bx <= 32 && t->by * t->bz + t->bx * t->bz + t->bx * t->by <= KB_TL_HALO_MAX)) t->variant = 2;
        const bool ll = t->variant != 0;
        if (ll) {
            t->face_len = t->by * t->bz + t->bx * t->bz + t->bx * t->by;
            const size_t slots = (size_t)nt * t->face_len;
            if (slots < ((size_t)1 << 31) - 64) {
                if ((st = kb_alloc(&t->mail, slots + 16)) != KB_OK || (st = kb_alloc(&t->sync, 4)) != KB_OK) break;
                cudaMemsetAsync(t->mail, 0, slots * sizeof(ulonglong2), c->stream);       // tag 0 is never used
                cudaMemsetAsync(t->sync, 0, 4 * sizeof(unsigned), c->stream);
                if (getenv("KB_TILES_TRACE") && (st = kb_alloc(&t->trace, (size_t)nt * 8)) != KB_OK) break;
                if (cudaStreamSynchronize(c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
            }
        }
        // the whole grid must be co-resident: a tile may wait on a tile owned by any other CTA
        int occ[2] = {1, 1};
        const int use_ll = t->mail != nullptr ? t->variant : 0;
        if (!t->mail) t->variant = 0;
        if (t->rows_per_thread == 4) { if ((st = tiles_occupancy<false, 4>(&occ[0], use_ll)) != KB_OK || (st = tiles_occupancy<true, 4>(&occ[1], use_ll)) != KB_OK) break; }
        else if (t->rows_per_thread == 2) { if ((st = tiles_occupancy<false, 2>(&occ[0], use_ll)) != KB_OK || (st = tiles_occupancy<true, 2>(&occ[1], use_ll)) != KB_OK) break; }
        else { if ((st = tiles_occupancy<false, 1>(&occ[0], use_ll)) != KB_OK || (st = tiles_occupancy<true, 1>(&occ[1], use_ll)) != KB_OK) break; }
        for (int u = 0; u < 2; ++u) {
            int g = std::max(1, occ[u]) * c->sm_count;
            if (getenv("KB_TILES_GRID")) g = std::min(g, std::max(1, atoi(getenv("KB_TILES_GRID"))));
            t->grid[u] = std::min(g, t->ntiles);
        }
    } while (0);
    if (st != KB_OK) { kb_set_error("ilu0: tile schedule setup failed"); kb_tiles_free(t); return st; }
    *out = t;
    return KB_OK;
}

int kb_tiles_apply(kb_pc_s* pc, KbTileSolve* t, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask) {
    kb_ctx_s* c = pc->a->ctx;
    KbTileArgs a{};
    a.rp = pc->l_rp; a.col = pc->l_col; a.lu = pc->lu; a.dptr = pc->diag_ptr; a.inv_diag = pc->inv_diag;
    a.n = (int)pc->a->n; a.nx = t->nx; a.ny = t->ny; a.nz = t->nz; a.bx = t->bx; a.by = t->by; a.bz = t->bz;
    a.tx = t->tx; a.ty = t->ty; a.tz = t->tz; a.ntiles = t->ntiles; a.err = t->err; a.skip_ctl = skip_ctl; a.skip_mask = skip_mask;
    a.mail = t->mail; a.sync = t->sync; a.face_len = t->face_len;
    a.lag = getenv("KB_TILES_LAG") ? atoi(getenv("KB_TILES_LAG")) : 3;
    a.trace = t->trace;
    if (t->mail && t->variant == 2) {       // packets at tile granularity: no flags, no memset
        for (int u = 0; u < 2; ++u) {
            a.order = t->order[u]; a.flags = nullptr; a.rhs = u == 0 ? d_r : pc->tmp; a.out = u == 0 ? pc->tmp : d_z;
            KbLaunch L(c, KB_K_TRSV);
            if (t->rows_per_thread == 4) { if (u) kb_trsv_tiles_pk<true, 4><<<t->grid[1], KB_THREADS, 0, c->stream>>>(a); else kb_trsv_tiles_pk<false, 4><<<t->grid[0], KB_THREADS, 0, c->stream>>>(a); }
            else if (t->rows_per_thread == 2) { if (u) kb_trsv_tiles_pk<true, 2><<<t->grid[1], KB_THREADS, 0, c->stream>>>(a); else kb_trsv_tiles_pk<false, 2><<<t->grid[0], KB_THREADS, 0, c->stream>>>(a); }
            else { if (u) kb_trsv_tiles_pk<true, 1><<<t->grid[1], KB_THREADS, 0, c->stream>>>(a); else kb_trsv_tiles_pk<false, 1><<<t->grid[0], KB_THREADS, 0, c->stream>>>(a); }
        }
        KB_CUDA(cudaGetLastError());
        return KB_OK;
    }
    if (t->mail) {       // fine-grained packet variant: no flags, no memset
        for (int u = 0; u < 2; ++u) {
            a.order = t->order[u]; a.flags = nullptr; a.rhs = u == 0 ? d_r : pc->tmp; a.out = u == 0 ? pc->tmp : d_z;
            KbLaunch L(c, KB_K_TRSV);
            if (t->rows_per_thread == 4) { if (u) kb_trsv_tiles_ll<true, 4><<<t->grid[1], KB_TL_THREADS, 0, c->stream>>>(a); else kb_trsv_tiles_ll<false, 4><<<t->grid[0], KB_TL_THREADS, 0, c->stream>>>(a); }
            else if (t->rows_per_thread == 2) { if (u) kb_trsv_tiles_ll<true, 2><<<t->grid[1], KB_TL_THREADS, 0, c->stream>>>(a); else kb_trsv_tiles_ll<false, 2><<<t->grid[0], KB_TL_THREADS, 0, c->stream>>>(a); }
            else { if (u) kb_trsv_tiles_ll<true, 1><<<t->grid[1], KB_TL_THREADS, 0, c->stream>>>(a); else kb_trsv_tiles_ll<false, 1><<<t->grid[0], KB_TL_THREADS, 0, c->stream>>>(a); }
        }
        KB_CUDA(cudaGetLastError());
        return KB_OK;
    }
    a.lag = getenv("KB_TILES_SLEEP") ? atoi(getenv("KB_TILES_SLEEP")) : 0;      // (the flag kernel reuses the field as its poll back-off)
    KB_CUDA(cudaMemsetAsync(t->flags, 0, 2 * (size_t)t->ntiles * sizeof(int), c->stream));
    {
        a.order = t->order[0]; a.flags = t->flags; a.rhs = d_r; a.out = pc->tmp;
        KbLaunch L(c, KB_K_TRSV);
        if (t->rows_per_thread == 4) kb_trsv_tiles<false, 4><<<t->grid[0], KB_THREADS, 0, c->stream>>>(a);
        else if (t->rows_per_thread == 2) kb_trsv_tiles<false, 2><<<t->grid[0], KB_THREADS, 0, c->stream>>>(a);
        else kb_trsv_tiles<false, 1><<<t->grid[0], KB_THREADS, 0, c->stream>>>(a);
    }
    {
        a.order = t->order[1]; a.flags = t->flags + t->ntiles; a.rhs = pc->tmp; a.out = d_z;
        KbLaunch L(c, KB_K_TRSV);
        if (t->rows_per_thread == 4) kb_trsv_tiles<true, 4><<<t->grid[1], KB_THREADS, 0, c->stream>>>(a);
        else if (t->rows_per_thread == 2) kb_trsv_tiles<true, 2><<<t->grid[1], KB_THREADS, 0, c->stream>>>(a);
        else kb_trsv_tiles<true, 1><<<t->grid[1], KB_THREADS, 0, c->stream>>>(a);
    }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// diagnostics (not part of the ABI header): per-tile timeline of the last apply; returns tiles per solve
int kb_tiles_trace_get(KbTileSolve* t, unsigned long long* out, int* tx, int* ty, int* tz) {
    if (!t || !t->trace) return 0;
    cudaMemcpy(out, t->trace, (size_t)t->ntiles * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    *tx = t->tx; *ty = t->ty; *tz = t->tz;
    return t->ntiles;
}
