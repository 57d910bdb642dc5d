// kb_trsv_march.cu — pencil-marching triangular solves for grid-structured ILU(0) factors (the default for
// full 5-/7-point box-grid stencils; the block-wavefront tiles of kb_trsv_tiles.cu remain the fallback).
//
// A lower solve on a lexicographic box grid is y(a,b,c) = r - cC*y(a,b,c-1) - cB*y(a,b-1,c) - cA*y(a-1,b,c): a
// hyperplane wavefront.  Level scheduling pays one inter-CTA hop per hyperplane (766 on 256^3), the tile kernel one
// per tile level plus a CTA barrier per in-tile step.  Here ONE WARP owns a pencil of LX x LY grid lines and marches
// along the third axis: lane (la,lb) handles c = t - la - lb at step t, so that
//   * y(a,b,c-1) is the lane's own previous value (a register),
//   * y(a-1,b,c) and y(a,b-1,c) were produced one step earlier by lanes l-1 and l-LX: two warp shuffles,
//   * no block barrier, no shared-memory exchange between lanes, no fence anywhere in the loop.
// Memory layout is what makes it stream (v1 of this kernel read the skewed wavefront straight from row-major arrays:
// 32 different cache lines per warp access, 0.45 us per step):
//   * the factor is kept a second time "pre-skewed": coef[(pencil*nsteps + t)*32 + lane] for the three neighbour
//     coefficients (and 1/u_ii), i.e. exactly the order the march consumes them - one contiguous 256-byte read per
//     stream per step, no column indices (a solve moves 40 (L) / 48 (U) bytes per row);
//   * rhs and the solution stay row-major; every lane loads ITS element of the un-skewed slab c = t + D (four 64-byte
//     segments per warp) into a lane-private shared-memory delay line and consumes it D + la + lb steps later; results
//     go through a second delay line and are stored when their slab is complete.  All staging is cp.async, D deep;
//   * values crossing a pencil face travel as 16-byte packets {lo32|tag, hi32|tag} through an L2-resident mailbox
//     indexed by the CONSUMER's step (the 12 packets a warp needs per step are contiguous).  The data carries its own
//     flag (the NCCL-LL idea): the producer needs no release fence, the consumer no acquire; the tag is the apply's
//     epoch, so the mailbox is never cleared.  Packets are prefetched PD steps ahead by the first LX+LY lanes and
//     handed to the face lanes by shuffle.
// The U solve is the same march in mirrored coordinates.  2-D grids march along j with 32 x 1 pencils.
// Per-row operation order is the oracle's (ascending column: L = c-1, b-1, a-1; U = a+1, b+1, c+1; mul then sub;
// U multiplies by 1/u_ii last), absent neighbours contribute "- 0.0 * 0.0", which leaves every value unchanged:
// results are bit-identical to the level-scheduled solves.
// Requirement checked at setup from the factor's pattern alone: every in-grid neighbour of a row is stored
// (full stencil).  Anything else keeps the tile / level-scheduled kernels.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kb_objects.h"

#define KB_MARCH_WARPS 16            // pencils marched concurrently by one CTA (one per warp), upper bound
#define KB_MARCH_D 6                 // coefficient / rhs prefetch distance (steps) through the cp.async rings
#define KB_MARCH_PD 4                // neighbour-packet prefetch distance (steps), register ring; must divide LX and LY

struct KbMarch {
    int nx = 0, ny = 0, nz = 0;      // march-space grid: a (lanes), b (lanes), c (march axis)
    int lx = 0, ly = 0;              // pencil cross-section, lx * ly == 32
    int px = 0, py = 0, npencils = 0, nsteps = 0, faces = 0;
    int ga = 4, gb = 4, gpx = 0, gpy = 0, ngroups = 0;   // CTA groups of ga x gb pencils
    int n = 0;
    double* coef[2] = {nullptr, nullptr};   // [upper]: pre-skewed and interleaved, [(pencil*nsteps + t)*NC + e][32], e = A,B,C(,1/u_ii); NC = 3 (L) / 4 (U)
    int* order = nullptr;            // group ids by level (ascending)
    ulonglong2* mail = nullptr;      // packets: [pencil][consumer step][face lane]
    unsigned* sync = nullptr;        // [0],[1] epoch of L / U ; [2],[3] finish tickets
    unsigned* err = nullptr;         // borrowed: the preconditioner's error word
    unsigned long long* trace = nullptr;   // diagnostics, allocated on demand (KB_MARCH_TRACE=1)
    int grid = 1, warps = KB_MARCH_WARPS;
    int lean = 1, threads = 0, lag = 8;   // lean kernel (kb_trsv_lean): default; KB_MARCH_LEAN=0 keeps the v2 kernel
    size_t smem[2] = {0, 0};
};

// ---- setup: pre-skewed copy of the factor + full-stencil check ---------------------------------------------------
// gx, gy: natural grid strides (sB == 0: two-dimensional grid, marched along j).  (nx,ny,nz), LX, LY: march space.
__global__ void k_march_skew(const int* __restrict__ rp, const int* __restrict__ col, const int* __restrict__ dp, const double* __restrict__ lu,
                             const double* __restrict__ inv_ud, int n, int gx, int gy, int sB, int sC, int nx, int ny, int nz, int LX, int LY,
                             int px, int nsteps, double* __restrict__ cL, double* __restrict__ cU, int* __restrict__ bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int i = r % gx, j = sB ? (r / gx) % gy : 0;
    double la = 0.0, lb = 0.0, lc = 0.0, ua = 0.0, ub = 0.0, uc = 0.0;
    int seen = 0, want = 0;
    const int d = dp[r];
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
        if (p == d) continue;
        const int c = col[p];
        const int off = p < d ? r - c : c - r;
        if (off <= 0) { atomicExch(bad, 1); continue; }
        int which = -1;
        if (off == 1) which = 0; else if (sB && off == sB) which = 1; else if (off == sC) which = 2;
        if (which < 0 || ((seen >> (which + (p < d ? 0 : 3))) & 1)) { atomicExch(bad, 1); continue; }
        seen |= 1 << (which + (p < d ? 0 : 3));
        const double v = lu[p];
        if (p < d) { if (which == 0) la = v; else if (which == 1) lb = v; else lc = v; }
        else { if (which == 0) ua = v; else if (which == 1) ub = v; else uc = v; }
    }
    // neighbours the grid says must exist (and are rows of this block)
    if (i >= 1) want |= 1;
    if (sB && j >= 1) want |= 2;
    if (r - sC >= 0) want |= 4;
    if (i + 1 < gx && r + 1 < n) want |= 8;
    if (sB && j + 1 < gy && r + sB < n) want |= 16;
    if (r + sC < n) want |= 32;
    if (seen != want) atomicExch(bad, 1);
    // march coordinates of this row: natural for L, mirrored for U
    const int a0 = i, b0 = j, c0 = r / sC;
    {
        const int pa = a0 / LX, pb = b0 / LY, qa = a0 % LX, qb = b0 % LY;
        const size_t slot = ((size_t)(pa + px * pb) * nsteps + (c0 + qa + qb)) * (3 * 32) + (qa + LX * qb);
        cL[slot] = la; cL[slot + 32] = lb; cL[slot + 64] = lc;
    }
    {
        const int a1 = nx - 1 - a0, b1 = ny - 1 - b0, c1 = nz - 1 - c0;
        const int pa = a1 / LX, pb = b1 / LY, qa = a1 % LX, qb = b1 % LY;
        const size_t slot = ((size_t)(pa + px * pb) * nsteps + (c1 + qa + qb)) * (4 * 32) + (qa + LX * qb);
        cU[slot] = ua; cU[slot + 32] = ub; cU[slot + 64] = uc; cU[slot + 96] = inv_ud[r];
    }
}

// ---- device helpers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void kb_cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void kb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void kb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ ulonglong2 kb_pkt_load(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void kb_pkt_store(ulonglong2* p, double val, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(val);
    const unsigned long long t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ __forceinline__ bool kb_pkt_ok(const ulonglong2& v, unsigned tag) {
    return (unsigned)(v.x >> 32) == tag && (unsigned)(v.y >> 32) == tag;
}
__device__ __forceinline__ double kb_pkt_value(const ulonglong2& v) {
    return __longlong_as_double((long long)((v.x & 0xffffffffull) | (v.y << 32)));
}
// wait for a neighbour packet that the prefetch found stale: bounded, bails out when anybody raised the error flag
__device__ __noinline__ ulonglong2 kb_pkt_wait(const ulonglong2* p, unsigned tag, unsigned* err) {
    unsigned spins = 0;
    ulonglong2 v = kb_pkt_load(p);
    while (!kb_pkt_ok(v, tag)) {
        if (++spins > 64u) __nanosleep(spins > 8192u ? 1000 : 50);
        v = kb_pkt_load(p);
        if ((spins & 1023u) == 0u) {
            if (spins > (1u << 22)) atomicExch(err, 1u);
            if (*reinterpret_cast<volatile unsigned*>(err)) break;
        }
    }
    return v;
}
__device__ __forceinline__ void kb_mbar_init_cta(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kb_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void kb_mbar_wait_cta(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ unsigned long long kb_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

struct KbMarchArgs {
    const double* __restrict__ coef;   // pre-skewed, interleaved
    const double* __restrict__ rhs; double* out;
    int n, nx, ny, nz, px, py, npencils;
    int nsteps;                      // local steps of one pencil: nz + LX + LY - 2, rounded up to a multiple of 4
    int lag;                         // lean kernel: steps a consumer group keeps behind its L2 producers
    int gpx, gpy, ngroups;           // CTA groups of GA x GB pencils
    const int* __restrict__ order;   // group ids by level (ascending)
    ulonglong2* mail;
    unsigned* sync; unsigned* err;
    const KbCtl* skip_ctl; int skip_mask;
    unsigned long long* trace;       // diagnostics (nullptr normally): per pencil {entry, first step done, end, packet stalls} in ns
};

template <int LX, int LY>
struct KbMarchShape {
    static constexpr int SK = LX + LY - 2;           // largest lane skew
    static constexpr int RD = KB_MARCH_D + SK + 1;   // rhs delay line (slabs)
    static constexpr int OD = SK + 1;                // output delay line
    static constexpr int FACES = LY == 1 ? 1 : LX + LY;
    __host__ __device__ static constexpr int warp_doubles(bool upper) { return ((upper ? 4 : 3) * KB_MARCH_D + RD + OD) * 32; }
    __host__ __device__ static constexpr int face_doubles(int warps) { return 2 * warps * (LX + LY) + 2; }    // [parity][warp][A: LY | B: LX] + the step mbarrier
};

// One CTA marches a GROUP of GA x GB pencils in lockstep (one __syncthreads per step): pencil (wa,wb) runs LX*wa + LY*wb
// steps behind pencil (0,0), which is exactly one step more than the wavefront needs, so a face value written to shared
// memory in step t-1 is the one its neighbour consumes in step t.  Only the faces between CTAs go through the L2 mailbox.
template <bool UPPER, int LX, int LY, int GA, int GB>
__global__ void __launch_bounds__(GA * GB * 32, 1) kb_trsv_march(KbMarchArgs a) {
    if (kb_skip(a.skip_ctl, a.skip_mask)) return;
    typedef KbMarchShape<LX, LY> SH;
    constexpr int NC = UPPER ? 4 : 3;
    constexpr int D = KB_MARCH_D, PD = KB_MARCH_PD, SK = SH::SK, RD = SH::RD, OD = SH::OD, FACES = SH::FACES;
    constexpr int WARPS = GA * GB, FW = LX + LY;
    static_assert(LX * LY == 32, "a pencil is one warp");
    static_assert(FACES <= 32, "one loader lane per face packet");
    static_assert(LX % PD == 0 && (LY == 1 || LY % PD == 0), "pencil delays inside a group must be multiples of the packet ring length");
    extern __shared__ __align__(16) double kb_march_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wa = warp % GA, wb = warp / GA;
    double* cring = kb_march_smem + (size_t)warp * SH::warp_doubles(UPPER) + lane;      // [D][NC][32]
    double* rring = cring + NC * D * 32;                                                  // [RD][32]
    double* oring = rring + RD * 32;                                                      // [OD][32]
    double* faces = kb_march_smem + (size_t)WARPS * SH::warp_doubles(UPPER);              // [2][WARPS][FW]

    // L and U share the mailbox: their tags never coincide (even / odd), and both advance once per apply
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(a.sync + (UPPER ? 1 : 0)) + 1u;
    const unsigned tag = 2u * epoch + (UPPER ? 1u : 0u);
    const int la = lane % LX, lb = lane / LX, sk = la + lb;
    const long long plane = (long long)a.nx * a.ny;
    const int nsteps = a.nsteps;                                    // local steps of one pencil
    const int offset = LX * wa + LY * wb;                           // this pencil's delay inside the group
    const int nsteps_cta = nsteps + LX * (GA - 1) + LY * (GB - 1);

    for (int gi_ = blockIdx.x; gi_ < a.ngroups; gi_ += gridDim.x) {
        const int group = a.order[gi_];
        const int Pa = (group % a.gpx) * GA + wa, Pb = (group / a.gpx) * GB + wb;
        const bool valid = Pa < a.px && Pb < a.py;
        const int pencil = Pa + a.px * Pb;
        const int ca = Pa * LX + la, cb = Pb * LY + lb;
        const bool in_ab = valid && ca < a.nx && cb < a.ny;
        const int gi = UPPER ? a.nx - 1 - ca : ca, gj = UPPER ? a.ny - 1 - cb : cb;
        const long long row0 = (long long)gi + (long long)a.nx * gj;           // row at grid k = 0
        // number of planes k with row0 + x + plane*k < n (the grid may end in a ragged plane)
        auto kcount = [&](long long x) -> int {
            const long long room = (long long)a.n - row0 - x;
            if (!in_ab || room <= 0) return 0;
            const long long k = (room + plane - 1) / plane;
            return (int)(k > a.nz ? a.nz : k);
        };
        // windows of march indices c: rows that exist / rows whose a-1 (b-1) neighbour exists
        int c_lo, c_hi, cA_lo, cB_lo;
        if (!UPPER) { c_lo = 0; c_hi = kcount(0); cA_lo = 0; cB_lo = 0; }
        else { c_hi = a.nz; c_lo = a.nz - kcount(0); cA_lo = a.nz - kcount(1); cB_lo = a.nz - kcount(a.nx); }
        const long long rstep = UPPER ? -plane : plane;                        // row of c: rbase + rstep * c
        const long long rbase = UPPER ? row0 + plane * (a.nz - 1) : row0;
        // where this pencil's face values go / come from: shared memory inside the group, L2 packets across groups
        const bool a_from_smem = wa > 0, b_from_smem = wb > 0;
        const bool a_to_smem = wa + 1 < GA, b_to_smem = LY > 1 && wb + 1 < GB;
        const bool expA = valid && !a_to_smem && la == LX - 1 && Pa + 1 < a.px;
        const bool expB = valid && LY > 1 && !b_to_smem && lb == LY - 1 && Pb + 1 < a.py;
        // loader lanes: lane f < FACES fetches the packet of face lane f (A face: consumer lane (0,f); B face: (f-LY,0))
        const int cl = lane < LY ? lane * LX : (lane - LY) % 32;
        const int f_lo_a = __shfl_sync(0xffffffffu, cA_lo, cl), f_lo_b = __shfl_sync(0xffffffffu, cB_lo, cl), f_hi = __shfl_sync(0xffffffffu, c_hi, cl);
        int pk_lo = 0, pk_hi = 0;                 // local steps at which this loader lane's packet exists
        if (lane < FACES) {
            const int skew = lane < LY ? lane : lane - LY;
            const bool need = lane < LY ? (!a_from_smem && Pa > 0) : (!b_from_smem && Pb > 0);
            if (need && valid) { pk_lo = (lane < LY ? f_lo_a : f_lo_b) + skew; pk_hi = f_hi + skew; }
        }
        const ulonglong2* mail_in = a.mail + (size_t)pencil * nsteps * FACES + lane;
        ulonglong2* mail_oa = a.mail + (size_t)(pencil + 1) * nsteps * FACES + lb;
        ulonglong2* mail_ob = a.mail + (size_t)(pencil + a.px) * nsteps * FACES + LY + la;
        // local-step windows of this lane
        const int t_lo = c_lo + sk, t_hi = c_hi + sk;                 // row exists
        const int f_lo = c_lo + SK, f_hi2 = c_hi + SK;                // its slab is complete -> store
        // shared-memory face slots
        double* my_fa = faces + (size_t)warp * FW;                    // + parity * WARPS * FW ; A: [lb], B: [LY + la]
        const double* in_fa = faces + (size_t)(warp - 1) * FW;        // pencil (wa-1, wb)
        const double* in_fb = faces + (size_t)(warp - GA) * FW;       // pencil (wa, wb-1)

        const double* p_coef = a.coef + (size_t)pencil * nsteps * (NC * 32) + lane;       // coefficients of local step 0
        auto prefetch = [&](int t, int cslot, int rslot) {      // coefficients of local step t, rhs slab c = t
            if (valid && t < nsteps) {
                double* s = cring + cslot * (NC * 32);
                const double* g = p_coef + (size_t)t * (NC * 32);
                kb_cp_async8(s, g);
                if (LY > 1) kb_cp_async8(s + 32, g + 32);
                kb_cp_async8(s + 64, g + 64);
                if (UPPER) kb_cp_async8(s + 96, g + 96);
                if (t >= c_lo && t < c_hi) kb_cp_async8(rring + rslot * 32, a.rhs + (rbase + rstep * t));
            }
            kb_cp_async_commit();          // one group per step, empty or not: the waits below count groups
        };
#pragma unroll
        for (int t = 0; t < D; ++t) prefetch(t, t, t);
        // Cross-CTA faces: start only when the producer is PD+1 steps further than the wavefront needs, so that the
        // packet prefetches below always find their data (all pencils then run at the same pace and keep the distance).
        if (pk_hi > pk_lo) {
            const int tw = min(pk_lo + PD + 1, pk_hi - 1);
            const ulonglong2 q = kb_pkt_wait(mail_in + (size_t)tw * FACES, tag, a.err);
            (void)q;
        }
        __syncwarp();
        ulonglong2 pk[PD];
#pragma unroll
        for (int u = 0; u < PD; ++u) {
            pk[u] = make_ulonglong2(0ull, 0ull);
            if (u >= pk_lo && u < pk_hi) pk[u] = kb_pkt_load(mail_in + (size_t)u * FACES);
        }
        // running ring positions (no modulo in the loop).  Operands are read one step ahead (see the tail below).
        int cs = 0;                          // coefficient slot of local step tl           (tl mod D)
        int r_use = (RD - sk) % RD;          // rhs slot of this lane's slab tl - sk        ((tl - sk) mod RD)
        int r_ld = D % RD;                   // rhs slot of slab tl + D
        int o_w = (OD - sk) % OD;            // output slot of slab tl - sk
        int o_r = (OD - SK) % OD;            // output slot of the slab completed at step tl: tl - SK
        // operands of local step 0
        kb_cp_async_wait<D - 1>();
        double vA = cring[0], vB = LY > 1 ? cring[32] : 0.0, vC = cring[64], dg = UPPER ? cring[96] : 1.0, rh = rring[r_use * 32];
        double ya_s = 0.0, yb_s = 0.0, yprev = 0.0;       // neighbour values shuffled at the end of the previous step
        unsigned long long tr_entry = 0ull, tr_first = 0ull; unsigned tr_stalls = 0u;
        if (a.trace) tr_entry = kb_gtime();
        __syncthreads();                     // face buffers of the previous group are no longer read
        for (int t0 = 0; t0 < nsteps_cta; t0 += PD) {
#pragma unroll
            for (int u = 0; u < PD; ++u) {
                const int t = t0 + u;
                if (t >= nsteps_cta) break;                     // CTA-uniform
                const int tl = t - offset;
                const bool in_range = valid && tl >= 0 && tl < nsteps;      // warp-uniform
                // ---- neighbours' step t-1 values -> s -> this pencil's faces
                double s = 0.0;
                if (in_range) {
                    const bool act = tl >= t_lo && tl < t_hi;
                    double ya = ya_s, yb = yb_s;
                    const int par = ((t - 1) & 1) * (WARPS * FW);
                    if (a_from_smem) { if (la == 0) ya = in_fa[par + lb]; }
                    if (LY > 1 && b_from_smem) { if (lb == 0) yb = in_fb[par + LY + la]; }
                    if (!a_from_smem || (LY > 1 && !b_from_smem)) {      // warp-uniform: this pencil has an L2 face (or a domain face)
                        double pv = 0.0;
                        if (tl >= pk_lo && tl < pk_hi) {            // loader lanes whose packet exists at this step
                            ulonglong2 q = pk[u];
                            if (!kb_pkt_ok(q, tag)) { q = kb_pkt_wait(mail_in + (size_t)tl * FACES, tag, a.err); ++tr_stalls; }
                            pv = kb_pkt_value(q);
                        }
                        pk[u] = make_ulonglong2(0ull, 0ull);
                        if (tl + PD >= pk_lo && tl + PD < pk_hi) pk[u] = kb_pkt_load(mail_in + (size_t)(tl + PD) * FACES);
                        if (!a_from_smem) { const double pa = __shfl_sync(0xffffffffu, pv, lb); if (la == 0) ya = pa; }       // 0.0 at a domain face
                        if (LY > 1 && !b_from_smem) { const double pb = __shfl_sync(0xffffffffu, pv, LY + la); if (lb == 0) yb = pb; }
                    }
                    if (!UPPER) { s = rh - vC * yprev; if (LY > 1) s = s - vB * yb; s = s - vA * ya; }      // ascending column: c-1, b-1, a-1
                    else { s = rh - vA * ya; if (LY > 1) s = s - vB * yb; s = s - vC * yprev; s = s * dg; }  // a+1, b+1, c+1, then 1/u_ii
                    if (!act) s = 0.0;
                    yprev = s;
                    // faces: to the neighbour pencil of this group through shared memory, else as L2 packets
                    const int parw = (t & 1) * (WARPS * FW);
                    if (a_to_smem) { if (la == LX - 1) my_fa[parw + lb] = s; }
                    else if (expA && act) kb_pkt_store(mail_oa + (size_t)(tl - (LX - 1)) * FACES, s, tag);      // consumer step of (0,lb): c + lb
                    if (LY > 1) {
                        if (b_to_smem) { if (lb == LY - 1) my_fa[parw + LY + la] = s; }
                        else if (expB && act) kb_pkt_store(mail_ob + (size_t)(tl - (LY - 1)) * FACES, s, tag);  // consumer step of (la,0): c + la
                    }
                }
                // ---- delay lines, next step's operands, refills
                if (in_range) {
                    ya_s = __shfl_up_sync(0xffffffffu, s, 1);
                    yb_s = LY > 1 ? __shfl_up_sync(0xffffffffu, s, LX) : 0.0;
                    oring[o_w * 32] = s;
                    // slab tl - SK is complete: every lane stores its element (un-skewed, coalesced)
                    if (tl >= f_lo && tl < f_hi2) a.out[rbase + rstep * (tl - SK)] = oring[o_r * 32];
                    // operands of step tl + 1 into registers, then refill the slots they leave
                    cs = cs + 1 == D ? 0 : cs + 1;
                    r_use = r_use + 1 == RD ? 0 : r_use + 1;
                    kb_cp_async_wait<D - 2>();
                    {
                        const double* cs_p = cring + cs * (NC * 32);
                        vA = cs_p[0]; if (LY > 1) vB = cs_p[32]; vC = cs_p[64]; if (UPPER) dg = cs_p[96];
                        rh = rring[r_use * 32];
                    }
                    // (the ring slot of step tl was read into registers one step ago: its refill cannot overtake a read)
                    prefetch(tl + D, cs == 0 ? D - 1 : cs - 1, r_ld);
                    r_ld = r_ld + 1 == RD ? 0 : r_ld + 1;
                    o_w = o_w + 1 == OD ? 0 : o_w + 1;
                    o_r = o_r + 1 == OD ? 0 : o_r + 1;
                    if (a.trace && tl == 0) tr_first = kb_gtime();
                }
                __syncthreads();                 // step barrier: faces written in step t are read in step t + 1
            }
        }
        kb_cp_async_wait<0>();
        __syncwarp();
        if (a.trace && valid) {
            const unsigned st_all = __reduce_add_sync(0xffffffffu, tr_stalls);
            if (lane == 0) {
                unsigned long long* q = a.trace + ((size_t)(UPPER ? a.npencils : 0) + pencil) * 4;
                q[0] = tr_entry; q[1] = tr_first; q[2] = kb_gtime(); q[3] = st_all;
            }
        }
    }
    // the last CTA to finish publishes the epoch: every CTA has read it by then
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(a.sync + 2 + (UPPER ? 1 : 0), 1u);
        if (t == gridDim.x - 1u) {
            a.sync[2 + (UPPER ? 1 : 0)] = 0u;
            a.sync[UPPER ? 1 : 0] = epoch;
            __threadfence();
        }
    }
}

// ---- lean march ("v3") -------------------------------------------------------------------------------------------
// Same schedule, same pre-skewed factor copy, same mailbox as kb_trsv_march above, but the step is stripped to the
// recurrence.  The v2 step above is ~235 SASS instructions on one dependent chain (ring bookkeeping, five cp.async,
// packet prefetch / validation / hand-over by shuffle, un-skewing output ring): 0.6 us per step with 4 warps, 1.2 us
// with 16, times nx+ny+nz steps.  Here the CTA is warp-specialised:
//   * compute warps (one pencil each) execute per step: 4-5 LDS (coefficients, rhs), 2 shuffles, 2 predicated LDS
//     (incoming face values), the 6-7 FP64 operations of the row, 2 predicated STS (outgoing faces), one predicated
//     STG (the solution, stored skewed: the 32-byte sectors are completed in L2 by the neighbouring lanes within
//     LX+LY steps, long before they are evicted), one cp.async (rhs delay line) and the step barrier;
//   * coefficients arrive through a 4-stage ring of 2 steps each, ONE cp.async.bulk (TMA engine, mbarrier
//     complete_tx) per stage and warp: consecutive steps of a pencil are contiguous in the pre-skewed copy;
//   * ALL L2 packet traffic belongs to helper warps: lane f owns one line of the group's outer faces, validates /
//     prefetches the incoming packets (register ring, PD steps ahead), drops their values into the shared-memory face
//     slots the compute lanes read like any in-group face, and publishes the outgoing face values one step after they
//     were computed.  Compute warps contain no global loads on the dependent chain and no packet code at all.
// Per-row arithmetic and its order are those of kb_trsv_march: same bits.
#define KM_PD 4                      // packet prefetch distance (steps) = unroll of the step loop

template <bool UPPER, int LX, int LY, int GA, int GB, int RD>
struct KmShape {
    static constexpr int NC = UPPER ? 4 : 3;
    static constexpr int SK = LX + LY - 2;
    static constexpr int D = RD - SK - 1;            // rhs prefetch distance (slabs)
    static constexpr int NCW = GA * GB;              // compute warps
    static constexpr int NA = GB * LY;               // lines of the group's A faces (in and out)
    static constexpr int NB = LY > 1 ? GA * LX : 0;  // lines of the B faces
    static constexpr int NHW = (NA + NB + 31) / 32;  // helper warps
    static constexpr int THREADS = (NCW + NHW) * 32;
    static constexpr int CSTEPS = 8;                 // coefficient ring: 4 stages x 2 steps
    static constexpr int WARP_DOUBLES = CSTEPS * NC * 32 + RD * 32;
    static constexpr int FA = (GA + 1) * GB * LY;    // faceA[par][wa' = 0..GA][wb][lb]: slot 0 = from L2, slot wa+1 = written by pencil (wa,wb)
    static constexpr int FB = LY > 1 ? (GB + 1) * GA * LX : 0;
    static constexpr size_t SMEM = ((size_t)NCW * WARP_DOUBLES + 2 * (FA + FB)) * sizeof(double) + (size_t)NCW * 4 * sizeof(unsigned long long);
    static_assert((RD & (RD - 1)) == 0 && D >= 3, "rhs delay line: power of two, deep enough");
};

struct KmLane { int c_lo, c_hi, cA_lo, cB_lo; long long rbase, rstep; };
// march-space windows of lane (la,lb) of pencil (Pa,Pb): rows that exist (c_lo..c_hi) and rows whose a-1 / b-1 neighbour exists
template <bool UPPER, int LX, int LY>
__device__ __forceinline__ KmLane km_lane(const KbMarchArgs& a, int Pa, int Pb, int la, int lb) {
    KmLane g;
    const long long plane = (long long)a.nx * a.ny;
    const int ca = Pa * LX + la, cb = Pb * LY + lb;
    const bool in_ab = Pa >= 0 && Pb >= 0 && Pa < a.px && Pb < a.py && ca < a.nx && cb < a.ny;
    const int gi = UPPER ? a.nx - 1 - ca : ca, gj = UPPER ? a.ny - 1 - cb : cb;
    const long long row0 = (long long)gi + (long long)a.nx * gj;
    auto kcount = [&](long long x) -> int {
        const long long room = (long long)a.n - row0 - x;
        if (!in_ab || room <= 0) return 0;
        const long long k = (room + plane - 1) / plane;
        return (int)(k > a.nz ? a.nz : k);
    };
    if (!UPPER) { g.c_lo = 0; g.c_hi = kcount(0); g.cA_lo = 0; g.cB_lo = 0; }
    else { g.c_hi = a.nz; g.c_lo = a.nz - kcount(0); g.cA_lo = a.nz - kcount(1); g.cB_lo = a.nz - kcount(a.nx); }
    g.rstep = UPPER ? -plane : plane;
    g.rbase = UPPER ? row0 + plane * (a.nz - 1) : row0;
    return g;
}
__device__ __forceinline__ unsigned km_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void km_bulk(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(km_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(km_u32(dst)),
                 "l"(src), "r"(bytes), "r"(km_u32(bar)), "l"(pol)
                 : "memory");
}

template <bool UPPER, int LX, int LY, int GA, int GB, int RD>
__global__ void __launch_bounds__(KmShape<UPPER, LX, LY, GA, GB, RD>::THREADS, 1) kb_trsv_lean(KbMarchArgs a) {
    if (kb_skip(a.skip_ctl, a.skip_mask)) return;
    typedef KmShape<UPPER, LX, LY, GA, GB, RD> SH;
    constexpr int NC = SH::NC, SK = SH::SK, D = SH::D, NCW = SH::NCW, NA = SH::NA, NB = SH::NB, FA = SH::FA, FB = SH::FB;
    constexpr int RM = RD - 1, PD = KM_PD, FACES = LY == 1 ? 1 : LX + LY, THREADS = SH::THREADS;
    constexpr unsigned STAGE_BYTES = 2u * NC * 32u * sizeof(double);
    static_assert(LX * LY == 32, "a pencil is one warp");
    static_assert(LX % 4 == 0 && (LY == 1 || LY % 4 == 0), "pencil delays inside a group must be multiples of the unroll");
    extern __shared__ __align__(128) double km_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* faceA = km_smem + (size_t)NCW * SH::WARP_DOUBLES;        // [2][GA+1][GB][LY]
    double* faceB = faceA + 2 * FA;                                   // [2][GB+1][GA][LX]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(faceB + 2 * FB);     // [NCW][4]

    // L and U share the mailbox: their tags never coincide (even / odd), and both advance once per apply
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(a.sync + (UPPER ? 1 : 0)) + 1u;
    const unsigned tag = 2u * epoch + (UPPER ? 1u : 0u);
    const int nsteps = a.nsteps;                                          // local steps of one pencil (multiple of 4)
    const int nsteps_cta = nsteps + LX * (GA - 1) + LY * (GB - 1) + 4;   // + one block in which only the helpers work (last outgoing faces)
    if (threadIdx.x < NCW * 4) kb_mbar_init_cta(full + threadIdx.x, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned kc = 0;                  // coefficient stages consumed by this warp so far (slot = kc & 3, phase = (kc >> 2) & 1)
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));

    for (int gi_ = blockIdx.x; gi_ < a.ngroups; gi_ += gridDim.x) {
        const int group = a.order[gi_];
        const int Pa0 = (group % a.gpx) * GA, Pb0 = (group / a.gpx) * GB;
        for (int i = threadIdx.x; i < 2 * (FA + FB); i += THREADS) faceA[i] = 0.0;
        __syncthreads();
        if (warp < NCW) {
            // ================= compute warp: one pencil =================
            const int wa = warp % GA, wb = warp / GA;
            const int Pa = Pa0 + wa, Pb = Pb0 + wb;
            const bool valid = Pa < a.px && Pb < a.py;
            const int pencil = Pa + a.px * Pb;
            const int la = lane % LX, lb = lane / LX, sk = la + lb;
            const KmLane g = km_lane<UPPER, LX, LY>(a, Pa, Pb, la, lb);
            const int offset = LX * wa + LY * wb;                      // this pencil's delay inside the group
            const int t_lo = g.c_lo + sk;
            const unsigned span = (unsigned)(g.c_hi - g.c_lo);         // rows of this lane
            double* cring = km_smem + (size_t)warp * SH::WARP_DOUBLES;  // [4 stages][2 steps][NC][32]
            double* rring = cring + SH::CSTEPS * NC * 32 + lane;        // [RD][32], lane-private column
            unsigned long long* fullw = full + warp * 4;
            const double* p_coef = a.coef + (size_t)pencil * nsteps * (NC * 32);
            const double* fa_in = faceA + (wa * GB + wb) * LY + lb;    // read by la == 0 (+ parity * FA)
            double* fa_out = faceA + ((wa + 1) * GB + wb) * LY + lb;   // written by la == LX-1
            const double* fb_in = faceB + (wb * GA + wa) * LX + la;    // read by lb == 0 (+ parity * FB)
            double* fb_out = faceB + ((wb + 1) * GA + wa) * LX + la;   // written by lb == LY-1
            if (valid) {
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (2 * j < nsteps) { const unsigned sl = (kc + j) & 3u; km_bulk(cring + sl * (2 * NC * 32), p_coef + (size_t)j * (2 * NC * 32), STAGE_BYTES, fullw + sl, pol); }
                }
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    if ((unsigned)(c - g.c_lo) < span) kb_cp_async8(rring + (c & RM) * 32, a.rhs + (g.rbase + g.rstep * c));
                    kb_cp_async_commit();
                }
            }
            long long r_ld = g.rbase + g.rstep * D;       // row of slab tl + D
            long long r_out = g.rbase - g.rstep * sk;     // row of this lane at step tl: c = tl - sk
            double s = 0.0;
            __syncthreads();                              // the helpers' step-0 face values are in place
            unsigned long long tr_entry = 0ull, tr_first = 0ull;
            if (a.trace) tr_entry = kb_gtime();
            for (int t0 = 0; t0 < nsteps_cta; t0 += 4) {
                const int tl0 = t0 - offset;
                const bool in_range = valid && tl0 >= 0 && tl0 < nsteps;      // warp-uniform, same for the 4 steps
                if (a.trace && tl0 == 0) tr_first = kb_gtime();
                const unsigned half = (kc >> 1) & 1u, ph = (kc >> 2) & 1u;
                const double* cb = cring + half * (4 * NC * 32) + lane;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (in_range) {
                        const int tl = tl0 + u;
                        if ((u & 1) == 0) kb_mbar_wait_cta(fullw + (half << 1) + (u >> 1), ph);
                        const double* cp = cb + u * (NC * 32);
                        const double vA = cp[0], vB = LY > 1 ? cp[32] : 0.0, vC = cp[64], dg = UPPER ? cp[96] : 1.0;
                        kb_cp_async_wait<D - 1>();
                        const double rh = rring[((tl - sk) & RM) * 32];
                        double ya = __shfl_up_sync(0xffffffffu, s, 1);
                        double yb = LY > 1 ? __shfl_up_sync(0xffffffffu, s, LX) : 0.0;
                        const int par_r = (u & 1) ^ 1, par_w = u & 1;          // t0 is a multiple of 4
                        if (la == 0) ya = fa_in[par_r * FA];
                        if (LY > 1 && lb == 0) yb = fb_in[par_r * FB];
                        double v;
                        if (!UPPER) { v = rh - vC * s; if (LY > 1) v = v - vB * yb; v = v - vA * ya; }          // ascending column: c-1, b-1, a-1
                        else { v = rh - vA * ya; if (LY > 1) v = v - vB * yb; v = v - vC * s; v = v * dg; }      // a+1, b+1, c+1, then 1/u_ii
                        const bool act = (unsigned)(tl - t_lo) < span;
                        s = act ? v : 0.0;
                        if (la == LX - 1) fa_out[par_w * FA] = s;
                        if (LY > 1 && lb == LY - 1) fb_out[par_w * FB] = s;
                        if (act) a.out[r_out] = s;
                        r_out += g.rstep;
                        // rhs slab tl + D into the delay line
                        if ((unsigned)(tl + D - g.c_lo) < span) kb_cp_async8(rring + ((tl + D) & RM) * 32, a.rhs + r_ld);
                        kb_cp_async_commit();
                        r_ld += g.rstep;
                        if (u & 1) {      // this stage is consumed: refill its slot with the stage 4 ahead
                            __syncwarp();
                            if (lane == 0 && tl + 7 < nsteps) {
                                const unsigned sl = (half << 1) + (u >> 1);
                                km_bulk(cring + sl * (2 * NC * 32), p_coef + (size_t)(tl + 7) * (NC * 32), STAGE_BYTES, fullw + sl, pol);
                            }
                        }
                    }
                    __syncthreads();                      // step barrier: faces written in step t are read in step t + 1
                }
                if (in_range) kc += 2u;
            }
            kb_cp_async_wait<0>();
            if (a.trace && valid && lane == 0) {
                unsigned long long* q = a.trace + ((size_t)(UPPER ? a.npencils : 0) + pencil) * 4;
                q[0] = tr_entry; q[1] = tr_first; q[2] = kb_gtime();
            }
        } else {
            // ================= helper warp: the group's L2 faces =================
            const int line = (warp - NCW) * 32 + lane;
            const bool isA = line < NA, live = line < NA + NB;
            int wa_c = 0, wb_c = 0, la_c = 0, lb_c = 0, wa_p = 0, wb_p = 0, la_p = 0, lb_p = 0, fidx = 0;
            if (isA) { wb_c = wb_p = line / LY; lb_c = lb_p = line % LY; wa_p = GA - 1; la_p = LX - 1; fidx = lb_c; }
            else if (live) { const int q = line - NA; wa_c = wa_p = q / LX; la_c = la_p = q % LX; wb_p = GB - 1; lb_p = LY - 1; fidx = LY + la_c; }
            const int Pa_c = Pa0 + wa_c, Pb_c = Pb0 + wb_c, Pa_p = Pa0 + wa_p, Pb_p = Pb0 + wb_p;
            const KmLane gc = km_lane<UPPER, LX, LY>(a, Pa_c, Pb_c, la_c, lb_c);
            const KmLane gp = km_lane<UPPER, LX, LY>(a, Pa_p, Pb_p, la_p, lb_p);
            const bool has_in = live && Pa_c < a.px && Pb_c < a.py && (isA ? Pa_c > 0 : Pb_c > 0);
            const bool has_out = live && Pa_p < a.px && Pb_p < a.py && (isA ? Pa_p + 1 < a.px : Pb_p + 1 < a.py);
            const int skew_c = isA ? lb_c : la_c;
            int pk_lo = 0, pk_hi = 0;                      // consumer-local steps at which an incoming packet exists
            if (has_in) { pk_lo = (isA ? gc.cA_lo : gc.cB_lo) + skew_c; pk_hi = gc.c_hi + skew_c; if (pk_hi < pk_lo) pk_hi = pk_lo; }
            int tp_lo = 0, tp_hi = 0;                      // producer-local steps at which an outgoing value exists
            if (has_out) { tp_lo = gp.c_lo + la_p + lb_p; tp_hi = gp.c_hi + la_p + lb_p; if (tp_hi < tp_lo) tp_hi = tp_lo; }
            const unsigned in_span = (unsigned)(pk_hi - pk_lo), out_span = (unsigned)(tp_hi - tp_lo);
            const int off_c = LX * wa_c + LY * wb_c, off_p = LX * wa_p + LY * wb_p;
            const ulonglong2* mail_in = a.mail + (size_t)(Pa_c + a.px * Pb_c) * nsteps * FACES + fidx;                    // + tl * FACES
            ulonglong2* mail_out = a.mail + ((size_t)(Pa_p + a.px * Pb_p + (isA ? 1 : a.px)) * nsteps) * FACES + fidx;   // + (tlp - depth) * FACES
            const int depth = isA ? LX - 1 : LY - 1;
            double* in_slot = isA ? faceA + wb_c * LY + lb_c : faceB + wa_c * LX + la_c;                                   // slot 0 (+ parity * FA/FB)
            const double* out_slot = isA ? faceA + (GA * GB + wb_p) * LY + lb_p : faceB + (GB * GA + wa_p) * LX + la_p;   // slot GA / GB
            const int fstride = isA ? FA : FB;
            // keep the distance: start only when the producer is `lag` steps further than the wavefront needs
            if (in_span > 0u) {
                const int tw = min(pk_lo + a.lag, pk_hi - 1);
                (void)kb_pkt_wait(mail_in + (size_t)tw * FACES, tag, a.err);
            }
            // packet of compute step t' lives at consumer-local step t' - off_c; ring slot t' & 3
            ulonglong2 pk[PD];
            unsigned tr_stalls = 0u;
            {
                double v0 = 0.0;
                const int tl = 0 - off_c;
                if ((unsigned)(tl - pk_lo) < in_span) v0 = kb_pkt_value(kb_pkt_wait(mail_in + (size_t)tl * FACES, tag, a.err));
                if (live) in_slot[1 * fstride] = v0;       // step 0 reads parity (0 - 1) & 1
#pragma unroll
                for (int j = 1; j <= PD; ++j) {
                    const int tj = j - off_c;
                    pk[j & (PD - 1)] = make_ulonglong2(0ull, 0ull);
                    if ((unsigned)(tj - pk_lo) < in_span) pk[j & (PD - 1)] = kb_pkt_load(mail_in + (size_t)tj * FACES);
                }
            }
            __syncthreads();                              // pairs with the compute warps' initial barrier
            for (int t0 = 0; t0 < nsteps_cta; t0 += PD) {
#pragma unroll
                for (int u = 0; u < PD; ++u) {
                    const int t = t0 + u;
                    // incoming value of compute step t + 1 -> parity t & 1, then request the packet of step t + 1 + PD
                    {
                        const int tl = t + 1 - off_c;
                        double v = 0.0;
                        if ((unsigned)(tl - pk_lo) < in_span) {
                            ulonglong2 q = pk[(u + 1) & (PD - 1)];
                            if (!kb_pkt_ok(q, tag)) { q = kb_pkt_wait(mail_in + (size_t)tl * FACES, tag, a.err); ++tr_stalls; }
                            v = kb_pkt_value(q);
                        }
                        if (live) in_slot[(u & 1) * fstride] = v;
                        pk[(u + 1) & (PD - 1)] = make_ulonglong2(0ull, 0ull);
                        if ((unsigned)(tl + PD - pk_lo) < in_span) pk[(u + 1) & (PD - 1)] = kb_pkt_load(mail_in + (size_t)(tl + PD) * FACES);
                    }
                    // outgoing value computed in step t - 1 (parity (t - 1) & 1)
                    {
                        const int tlp = t - 1 - off_p;
                        if ((unsigned)(tlp - tp_lo) < out_span) kb_pkt_store(mail_out + (ptrdiff_t)(tlp - depth) * FACES, out_slot[((u & 1) ^ 1) * fstride], tag);
                    }
                    __syncthreads();
                }
            }
            if (a.trace) {      // blocking packet waits of this group's helper lanes -> slot 3 of the group's first pencil
                const unsigned st_all = __reduce_add_sync(0xffffffffu, tr_stalls);
                if (lane == 0 && Pa0 < a.px && Pb0 < a.py) atomicAdd(a.trace + ((size_t)(UPPER ? a.npencils : 0) + Pa0 + a.px * Pb0) * 4 + 3, (unsigned long long)st_all);
            }
        }
        // (the next group's face clearing is ordered after every read of this group by the barrier above)
    }
    // the last CTA to finish publishes the epoch: every CTA has read it by then
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(a.sync + 2 + (UPPER ? 1 : 0), 1u);
        if (t == gridDim.x - 1u) {
            a.sync[2 + (UPPER ? 1 : 0)] = 0u;
            a.sync[UPPER ? 1 : 0] = epoch;
            __threadfence();
        }
    }
}

// ---- host ------------------------------------------------------------------------------------------------------
void kb_march_free(KbMarch* m) {
    if (!m) return;
    KB_FREE(m->coef[0]); KB_FREE(m->coef[1]);
    KB_FREE(m->order); KB_FREE(m->mail); KB_FREE(m->sync); KB_FREE(m->trace);
    delete m;
}

typedef void (*kb_march_fn)(KbMarchArgs);
// group shapes: 3-D pencils of 8 x 4 lines in groups of g x g (g = 1, 2, 4); 2-D pencils of 32 lines in groups of g x 1
static kb_march_fn march_kernel(bool upper, bool two_d, int g) {
    if (two_d) {
        if (g == 1) return upper ? kb_trsv_march<true, 32, 1, 1, 1> : kb_trsv_march<false, 32, 1, 1, 1>;
        if (g == 2) return upper ? kb_trsv_march<true, 32, 1, 2, 1> : kb_trsv_march<false, 32, 1, 2, 1>;
        return upper ? kb_trsv_march<true, 32, 1, 4, 1> : kb_trsv_march<false, 32, 1, 4, 1>;
    }
    if (g == 1) return upper ? kb_trsv_march<true, 8, 4, 1, 1> : kb_trsv_march<false, 8, 4, 1, 1>;
    if (g == 2) return upper ? kb_trsv_march<true, 8, 4, 2, 2> : kb_trsv_march<false, 8, 4, 2, 2>;
    return upper ? kb_trsv_march<true, 8, 4, 4, 4> : kb_trsv_march<false, 8, 4, 4, 4>;
}
static size_t march_smem(bool upper, bool two_d, int g) {
    const int warps = two_d ? g : g * g;
    if (two_d) return ((size_t)warps * KbMarchShape<32, 1>::warp_doubles(upper) + KbMarchShape<32, 1>::face_doubles(warps)) * sizeof(double);
    return ((size_t)warps * KbMarchShape<8, 4>::warp_doubles(upper) + KbMarchShape<8, 4>::face_doubles(warps)) * sizeof(double);
}

// lean kernel: (ga, gb) group shapes.  3-D: 1x1, 2x2 (rhs ring 32), 2x4, 4x4 (ring 16: shared memory); 2-D: g x 1 (ring 64)
struct KmKernel { kb_march_fn fn; size_t smem; int threads; };
template <bool UPPER, int LX, int LY, int GA, int GB, int RD>
static KmKernel km_make() { typedef KmShape<UPPER, LX, LY, GA, GB, RD> SH; return KmKernel{kb_trsv_lean<UPPER, LX, LY, GA, GB, RD>, SH::SMEM, SH::THREADS}; }
static KmKernel lean_kernel(bool upper, bool two_d, int ga, int gb) {
    if (two_d) {
        if (ga == 1) return upper ? km_make<true, 32, 1, 1, 1, 64>() : km_make<false, 32, 1, 1, 1, 64>();
        if (ga == 2) return upper ? km_make<true, 32, 1, 2, 1, 64>() : km_make<false, 32, 1, 2, 1, 64>();
        return upper ? km_make<true, 32, 1, 4, 1, 64>() : km_make<false, 32, 1, 4, 1, 64>();
    }
    if (ga == 1) return upper ? km_make<true, 8, 4, 1, 1, 32>() : km_make<false, 8, 4, 1, 1, 32>();
    if (ga == 2 && gb == 2) return upper ? km_make<true, 8, 4, 2, 2, 32>() : km_make<false, 8, 4, 2, 2, 32>();
    if (ga == 2) return upper ? km_make<true, 8, 4, 2, 4, 16>() : km_make<false, 8, 4, 2, 4, 16>();
    return upper ? km_make<true, 8, 4, 4, 4, 16>() : km_make<false, 8, 4, 4, 4, 16>();
}

// gx, gy, gz: the box grid detected from the factor's pattern (kb_trsv_tiles.cu).  *out stays nullptr (KB_OK) when the
// pattern is not a full stencil.
int kb_march_build(kb_pc_s* pc, int gx, int gy, int gz, unsigned* d_err, KbMarch** out) {
    *out = nullptr;
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    const int n = (int)A->n;
    const bool two_d = gz == 1;
    KbMarch* m = new KbMarch;
    m->n = n; m->err = d_err;
    if (two_d) { m->nx = gx; m->ny = 1; m->nz = gy; m->lx = 32; m->ly = 1; m->faces = 1; }
    else { m->nx = gx; m->ny = gy; m->nz = gz; m->lx = 8; m->ly = 4; m->faces = 12; }
    m->px = (m->nx + m->lx - 1) / m->lx; m->py = (m->ny + m->ly - 1) / m->ly;
    m->nsteps = (m->nz + m->lx + m->ly - 2 + 3) & ~3;      // (padded steps: zero coefficients, no rows)
    const long long np = (long long)m->px * m->py;
    const long long slots = np * m->nsteps * m->faces;
    const long long skewed = np * m->nsteps * 32;
    if (np > (1 << 24) || slots * 16 > (4ll << 30) || skewed * 4 > (1ll << 33)) { delete m; return KB_OK; }
    m->npencils = (int)np;
    int st = KB_OK;
    int* d_bad = nullptr;
    do {
        for (int u = 0; u < 2 && st == KB_OK; ++u) {
            const size_t cnt = (size_t)skewed * (u ? 4 : 3) + 64;
            st = kb_alloc(&m->coef[u], cnt);
            if (st == KB_OK) cudaMemsetAsync(m->coef[u], 0, cnt * sizeof(double), c->stream);
        }
        if (st != KB_OK || (st = kb_alloc(&d_bad, 1)) != KB_OK) break;
        cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream);
        const int sB = two_d ? 0 : gx, sC = two_d ? gx : gx * gy;
        {
            KbLaunch L(c, KB_K_OTHER);
            k_march_skew<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(pc->l_rp, pc->l_col, pc->diag_ptr, pc->lu, pc->inv_diag, n, gx, gy, sB, sC, m->nx, m->ny,
                                                                           m->nz, m->lx, m->ly, m->px, m->nsteps, m->coef[0], m->coef[1], d_bad);
        }
        int h_bad = 0;
        if (cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        if (h_bad) { cudaFree(d_bad); kb_march_free(m); return KB_OK; }      // not a full stencil: other kernels take it
        // groups of ga x gb pencils in level order
        m->lean = getenv("KB_MARCH_LEAN") ? atoi(getenv("KB_MARCH_LEAN")) : 1;
        if (getenv("KB_MARCH_LAG")) m->lag = std::max(KM_PD + 1, atoi(getenv("KB_MARCH_LAG")));
        int g = m->lean && !two_d ? 2 : 4;
        int g2 = 0;                                   // lean 3-D only: "24" = 2 x 4 pencils
        if (getenv("KB_MARCH_GROUP")) { const int e = atoi(getenv("KB_MARCH_GROUP")); if (e == 1 || e == 2 || e == 4) g = e; else if (e == 24 && m->lean && !two_d) { g = 2; g2 = 4; } }
        m->ga = g; m->gb = two_d ? 1 : (g2 ? g2 : g);
        m->gpx = (m->px + m->ga - 1) / m->ga; m->gpy = (m->py + m->gb - 1) / m->gb;
        m->ngroups = m->gpx * m->gpy;
        std::vector<int> ord((size_t)m->ngroups), lev((size_t)m->ngroups);
        for (int b = 0, id = 0; b < m->gpy; ++b) for (int a2 = 0; a2 < m->gpx; ++a2, ++id) { ord[id] = id; lev[id] = a2 + b; }
        std::stable_sort(ord.begin(), ord.end(), [&](int p, int q) { return lev[p] < lev[q]; });
        if ((st = kb_alloc(&m->order, (size_t)m->ngroups)) != KB_OK) break;
        if (cudaMemcpyAsync(m->order, ord.data(), (size_t)m->ngroups * sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        if ((st = kb_alloc(&m->mail, (size_t)(slots + (long long)(m->px + 2) * m->nsteps * m->faces))) != KB_OK) break;
        cudaMemsetAsync(m->mail, 0, (size_t)slots * sizeof(ulonglong2), c->stream);     // tag 0 is never used
        if ((st = kb_alloc(&m->sync, 4)) != KB_OK) break;
        if (getenv("KB_MARCH_TRACE") && (st = kb_alloc(&m->trace, (size_t)np * 8)) != KB_OK) break;
        if (m->trace) cudaMemsetAsync(m->trace, 0, (size_t)np * 8 * sizeof(unsigned long long), c->stream);
        cudaMemsetAsync(m->sync, 0, 4 * sizeof(unsigned), c->stream);
        // one CTA per group; all CTAs co-resident (a group may wait on a group of any other CTA)
        m->warps = m->ga * m->gb;
        int cap = 1 << 30;
        for (int u = 0; u < 2 && st == KB_OK; ++u) {
            kb_march_fn f = march_kernel(u == 1, two_d, m->ga);
            size_t sh = march_smem(u == 1, two_d, m->ga);
            int threads = m->warps * 32;
            if (m->lean) { const KmKernel k = lean_kernel(u == 1, two_d, m->ga, m->gb); f = k.fn; sh = k.smem; threads = k.threads; }
            m->smem[u] = sh; m->threads = threads;
            if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, f, threads, sh) != cudaSuccess || occ < 1) { st = KB_SOLVE_ERROR; break; }
            cap = std::min(cap, occ * c->sm_count);
        }
        if (st != KB_OK) break;
        m->grid = std::max(1, std::min(cap, m->ngroups));
        if (getenv("KB_MARCH_GRID")) m->grid = std::max(1, std::min(m->grid, atoi(getenv("KB_MARCH_GRID"))));
        if (m->trace) fprintf(stderr, "[kb march] lean=%d group %dx%d threads %d smem %zu/%zu grid %d (cap %d) groups %d nsteps %d lag %d\n", m->lean, m->ga, m->gb, m->threads, m->smem[0], m->smem[1], m->grid, cap, m->ngroups, m->nsteps, m->lag);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
    } while (0);
    if (d_bad) cudaFree(d_bad);
    if (st != KB_OK) { cudaGetLastError(); kb_set_error("ilu0: pencil-march schedule setup failed"); kb_march_free(m); return st; }
    *out = m;
    return KB_OK;
}

int kb_march_apply(kb_pc_s* pc, KbMarch* m, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask) {
    kb_ctx_s* c = pc->a->ctx;
    const bool two_d = m->ly == 1;
    KbMarchArgs a{};
    a.n = m->n; a.nx = m->nx; a.ny = m->ny; a.nz = m->nz; a.px = m->px; a.py = m->py; a.npencils = m->npencils;
    a.gpx = m->gpx; a.gpy = m->gpy; a.ngroups = m->ngroups; a.nsteps = m->nsteps; a.lag = m->lag;
    a.order = m->order; a.mail = m->mail; a.sync = m->sync; a.err = m->err; a.skip_ctl = skip_ctl; a.skip_mask = skip_mask; a.trace = m->trace;
    if (m->trace) cudaMemsetAsync(m->trace, 0, (size_t)m->npencils * 8 * sizeof(unsigned long long), c->stream);
    for (int u = 0; u < 2; ++u) {
        a.coef = m->coef[u];
        a.rhs = u == 0 ? d_r : pc->tmp; a.out = u == 0 ? pc->tmp : d_z;
        KbLaunch L(c, KB_K_TRSV);
        if (m->lean) lean_kernel(u == 1, two_d, m->ga, m->gb).fn<<<m->grid, m->threads, m->smem[u], c->stream>>>(a);
        else march_kernel(u == 1, two_d, m->ga)<<<m->grid, m->warps * 32, march_smem(u == 1, two_d, m->ga), c->stream>>>(a);
    }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// diagnostics (not part of the ABI header): copy the per-pencil timeline of the last apply; returns pencils per solve
int kb_march_trace_get(KbMarch* m, unsigned long long* out, int* px, int* py) {
    if (!m || !m->trace) return 0;
    cudaMemcpy(out, m->trace, (size_t)m->npencils * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    *px = m->px; *py = m->py;
    return m->npencils;
}
