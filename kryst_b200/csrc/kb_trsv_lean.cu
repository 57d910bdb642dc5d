// kb_trsv_lean.cu — pencil-marching triangular solves for grid-structured ILU(0) factors, warp-specialised.
//
// A lower solve on a lexicographic box grid is y(a,b,c) = r - cC*y(a,b,c-1) - cB*y(a,b-1,c) - cA*y(a-1,b,c): a
// hyperplane wavefront of nx+ny+nz-2 levels.  ONE WARP owns a pencil of LX x (LY*R) grid lines and marches along c:
// lane (la,lb') holds R rows (b = R*lb' + r); the row with skew sk = la + b handles c = t - sk at local step t, so
//   * y(a,b,c-1) is the lane's own previous value, y(a,b-1,c) the lane's previous row or one shuffle away,
//     y(a-1,b,c) one shuffle away: no shared-memory exchange between lanes, no fence anywhere in the loop;
//   * GA x GB pencils form a group marched by one CTA in lockstep (one bar.sync per step), pencil (wa,wb) running
//     LX*wa + LY*R*wb steps behind pencil (0,0): a face value written to shared memory in step t is consumed in t+1;
//   * faces between groups travel as 16-byte packets {lo32|tag, hi32|tag} through an L2-resident mailbox indexed by
//     the consumer's step (the data is its own flag: no release fence, no acquire; tag = epoch of the apply).
// The kernel this replaces (kb_trsv_march.cu, v2) spent ~235 dependent SASS instructions per step in every warp.
// Here the CTA is specialised so that every warp's per-step instruction stream is short:
//   * compute warps: per row 4-5 LDS (coefficients, rhs), the 6-7 FP64 operations, one predicated store; per step
//     the shuffles, the predicated face loads / stores and one mbarrier poll.  No global loads, no address upkeep;
//   * the factor is stored a second time in exactly the order a GROUP consumes it:
//     coef[group][step][pencil][stream][lane] - ONE cp.async.bulk (TMA engine, mbarrier complete_tx) per CTA and
//     step, issued by the loader warp into an 8-stage ring; the loader also runs the rhs delay line of the L solve
//     (un-skewed, coalesced slabs by cp.async, consumed D + skew steps later);
//   * the L solve stores its result straight into the U solve's stream (stream 4 of the U array, pre-skewed and
//     mirrored), so the U solve has no rhs staging at all; the U solve stores the solution row-major (skewed 8-byte
//     stores: the 32-byte sectors are completed in L2 by the neighbouring lanes within a few steps);
//   * helper warps own ALL L2 packet traffic: lane f owns one line of the group's outer faces, validates / prefetches
//     incoming packets (register ring, PD steps ahead) into the shared-memory face slots the compute lanes read like
//     any in-group face, and publishes the outgoing face values one step after they were computed.
// Per-row operation order is the oracle's (ascending column: L = c-1, b-1, a-1; U = a+1, b+1, c+1; mul then sub; U
// multiplies by 1/u_ii last); absent neighbours contribute "- 0.0 * 0.0": results are bit-identical to the
// level-scheduled solves.  Requirement checked at setup from the factor's pattern alone: every in-grid neighbour of
// a row is stored (full stencil).  Anything else keeps the tile / level-scheduled kernels.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kb_objects.h"

#define KL_STAGE_LOG 4
#define KL_STAGES (1 << KL_STAGE_LOG) // coefficient ring (steps): the march needs its stage within KL_STAGES-1 steps of the request, i.e.
                                     // step time >= (DRAM + TMA latency) / (KL_STAGES - 1); 8 stages pinned the step at ~0.22 us
#define KL_PD 4                      // packet prefetch distance (steps) = unroll of the helpers' step loop

struct KlArgs {
    const double* __restrict__ coef; // [group][step][pencil][stream][32]
    const double* __restrict__ rhs;  // L: row-major right-hand side
    double* out;                     // L: the U solve's array (stream 4 receives y); U: row-major solution
    int n, nx, ny, nz, px, py, npencils;
    int nsteps, nsteps_cta, lag;
    int dbg;                         // timing experiments only (KB_LEAN_DBG bit mask; results are wrong when set)
    int gpx, gpy, ngroups;
    const int* __restrict__ order;   // group ids by level (ascending)
    ulonglong2* mail;
    unsigned* sync; unsigned* err;
    const KbCtl* skip_ctl; int skip_mask;
    unsigned long long* trace;       // diagnostics (nullptr normally): per pencil {entry, first step, end, packet stalls}
};
typedef void (*kl_fn)(KlArgs);

struct KbLean {
    int nx = 0, ny = 0, nz = 0;      // march-space grid: a (lanes), b (lanes x rows), c (march axis)
    int lx = 0, lyr = 0, r = 1;      // pencil cross-section lx x lyr, r rows per lane
    int px = 0, py = 0, npencils = 0, nsteps = 0, faces = 0;
    int ga = 2, gb = 2, gpx = 0, gpy = 0, ngroups = 0, nsteps_cta = 0;
    int n = 0;
    double* coef[2] = {nullptr, nullptr};
    int* order = nullptr;
    ulonglong2* mail = nullptr;
    unsigned* sync = nullptr;        // [0],[1] epoch of L / U ; [2],[3] finish tickets
    unsigned* err = nullptr;         // borrowed: the preconditioner's error word
    unsigned long long* trace = nullptr;
    int grid = 1, threads = 0, lag = KL_PD + 2, dbg = 0;
    size_t smem[2] = {0, 0};
    kl_fn fn[2] = {nullptr, nullptr};
};

// ---- device helpers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned kl_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kl_cp8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(kl_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void kl_cp8s(unsigned smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void kl_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void kl_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void kl_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kl_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kl_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(kl_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void kl_bulk(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kl_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(kl_u32(dst)),
                 "l"(src), "r"(bytes), "r"(kl_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ ulonglong2 kl_pkt_load(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void kl_pkt_store(ulonglong2* p, double val, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(val);
    const unsigned long long t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ __forceinline__ bool kl_pkt_ok(const ulonglong2& v, unsigned tag) { return (unsigned)(v.x >> 32) == tag && (unsigned)(v.y >> 32) == tag; }
__device__ __forceinline__ double kl_pkt_value(const ulonglong2& v) { return __longlong_as_double((long long)((v.x & 0xffffffffull) | (v.y << 32))); }
// wait for a packet that the prefetch found stale: bounded, bails out when anybody raised the error flag
__device__ __noinline__ ulonglong2 kl_pkt_wait(const ulonglong2* p, unsigned tag, unsigned* err) {
    unsigned spins = 0;
    ulonglong2 v = kl_pkt_load(p);
    while (!kl_pkt_ok(v, tag)) {
        if (++spins > 64u) __nanosleep(spins > 8192u ? 1000 : 50);
        v = kl_pkt_load(p);
        if ((spins & 1023u) == 0u) {
            if (spins > (1u << 22)) atomicExch(err, 1u);
            if (*reinterpret_cast<volatile unsigned*>(err)) break;
        }
    }
    return v;
}
__device__ __forceinline__ unsigned long long kl_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) :: "memory"); return t; }

// ---- geometry shared by the setup kernel and the solve -----------------------------------------------------------
// position of march-space row (a,b,c) in a group-ordered array of NS streams: [group][t][w][stream*R + r][lane]
struct KlGeo { int LX, LYR, R, GA, GB, gpx, nsteps_cta; };
__host__ __device__ __forceinline__ size_t kl_slot(const KlGeo& g, int NS, int a, int b, int c) {
    const int Pa = a / g.LX, Pb = b / g.LYR, qa = a % g.LX, qb = b % g.LYR;
    const int lane = qa + g.LX * (qb / g.R), rr = qb % g.R;
    const int wa = Pa % g.GA, wb = Pb % g.GB;
    const size_t group = (size_t)(Pa / g.GA) + (size_t)g.gpx * (Pb / g.GB);
    const int w = wa + g.GA * wb;
    const int t = c + qa + qb + g.LX * wa + g.LYR * wb;
    return ((group * g.nsteps_cta + t) * (g.GA * g.GB) + w) * ((size_t)NS * g.R * 32) + rr * 32 + lane;      // + e * R * 32 for stream e
}

// ---- setup: group-ordered copies of the factor + full-stencil check ----------------------------------------------
// gx, gy: natural grid strides (sB == 0: two-dimensional grid, marched along j).  (nx,ny,nz): march space.
__global__ void k_lean_skew(const int* __restrict__ rp, const int* __restrict__ col, const int* __restrict__ dp, const double* __restrict__ lu,
                            const double* __restrict__ inv_ud, int n, int gx, int gy, int sB, int sC, int nx, int ny, int nz, KlGeo geo,
                            double* __restrict__ cL, double* __restrict__ cU, int* __restrict__ bad) {
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
    const int a0 = i, b0 = j, c0 = r / sC;
    const int R32 = geo.R * 32;
    {
        const size_t s = kl_slot(geo, 3, a0, b0, c0);
        cL[s] = la; cL[s + R32] = lb; cL[s + 2 * R32] = lc;
    }
    {
        const size_t s = kl_slot(geo, 5, nx - 1 - a0, ny - 1 - b0, nz - 1 - c0);
        cU[s] = ua; cU[s + R32] = ub; cU[s + 2 * R32] = uc; cU[s + 3 * R32] = inv_ud[r];
    }
}

// ---- the solve ---------------------------------------------------------------------------------------------------
template <bool UPPER, int LX, int LY, int R, int GA, int GB, int RD>
struct KlShape {
    static constexpr int NS = UPPER ? 5 : 3;             // streams per row: coefficients A,B,C (, 1/u_ii, rhs)
    static constexpr int LYR = LY * R;
    static constexpr int SK = LX + LYR - 2;              // largest skew
    static constexpr int D = RD - SK - 1;                // rhs prefetch distance (slabs), L only
    static constexpr int NCW = GA * GB;                  // compute warps
    static constexpr int NA = GB * LYR;                  // lines of the group's A faces (in and out)
    static constexpr int NB = LY > 1 ? GA * LX : 0;      // lines of the B faces
    static constexpr int NHW = (NA + NB + 31) / 32;      // helper warps
    static constexpr int THREADS = (NCW + 2 * NHW + 2) * 32; // + the stage warp (TMA, mbarriers) + the mover warp (L: rhs slabs in, U: solution slabs out)
    static constexpr int STEP_DOUBLES = NCW * NS * R * 32;
    static constexpr int OD = SK + 2 <= 16 ? 16 : (SK + 2 <= 32 ? 32 : 64);      // U: output delay line (slabs), power of two >= SK + 2
    static constexpr int RR_DOUBLES = UPPER ? NCW * R * OD * 32 : NCW * R * RD * 32;
    static constexpr int FA = (GA + 1) * GB * LYR;       // faceA[par][wa' = 0..GA][wb][qb]: slot 0 = from L2, slot wa+1 = written by pencil (wa,wb)
    static constexpr int FB = LY > 1 ? (GB + 1) * GA * LX : 0;
    static constexpr size_t SMEM = ((size_t)KL_STAGES * STEP_DOUBLES + RR_DOUBLES + 2 * (FA + FB)) * sizeof(double) + KL_STAGES * sizeof(unsigned long long);
    static_assert((RD & (RD - 1)) == 0 && (UPPER || D >= 3), "rhs delay line: power of two, deep enough");
    static_assert(LX * LY == 32, "a pencil is one warp");
    static_assert(LX % 4 == 0 && (LY == 1 || LYR % 4 == 0), "pencil delays inside a group must be multiples of the unroll");
};

struct KlRow { int c_lo, c_hi, cA_lo, cB_lo; long long rbase, rstep; int gi, gj; };
// march-space windows of row (la,qb) of pencil (Pa,Pb): c for which the row exists / its a-1 (b-1) neighbour exists
template <bool UPPER, int LX, int LYR>
__device__ __forceinline__ KlRow kl_row(const KlArgs& a, int Pa, int Pb, int la, int qb) {
    KlRow g;
    const long long plane = (long long)a.nx * a.ny;
    const int ca = Pa * LX + la, cb = Pb * LYR + qb;
    const bool in_ab = Pa >= 0 && Pb >= 0 && Pa < a.px && Pb < a.py && ca < a.nx && cb < a.ny;
    g.gi = UPPER ? a.nx - 1 - ca : ca; g.gj = UPPER ? a.ny - 1 - cb : cb;
    const long long row0 = (long long)g.gi + (long long)a.nx * g.gj;
    auto kcount = [&](long long x) -> int {      // planes k with row0 + x + plane*k < n
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

template <bool UPPER, int LX, int LY, int R, int GA, int GB, int RD>
__global__ void __launch_bounds__(KlShape<UPPER, LX, LY, R, GA, GB, RD>::THREADS, 1) kb_trsv_lean(KlArgs a) {
    if (kb_skip(a.skip_ctl, a.skip_mask)) return;
    typedef KlShape<UPPER, LX, LY, R, GA, GB, RD> SH;
    constexpr int NS = SH::NS, LYR = SH::LYR, SK = SH::SK, D = SH::D, NCW = SH::NCW, NA = SH::NA, NB = SH::NB, NHW = SH::NHW, FA = SH::FA, FB = SH::FB;
    constexpr int RM = RD - 1, PD = KL_PD, FACES = LY == 1 ? 1 : LX + LYR, THREADS = SH::THREADS, STEP = SH::STEP_DOUBLES;
    constexpr unsigned STEP_BYTES = (unsigned)STEP * sizeof(double);
    constexpr int R32 = R * 32;
    (void)SK;
    extern __shared__ __align__(128) double kl_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* cst = kl_smem;                                            // [KL_STAGES][NCW][NS*R][32]
    double* rr = cst + (size_t)KL_STAGES * STEP;                      // L: [NCW][R][RD][32]
    double* faceA = rr + SH::RR_DOUBLES;                              // [2][GA+1][GB][LYR]
    double* faceB = faceA + 2 * FA;                                   // [2][GB+1][GA][LX]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(faceB + 2 * FB);     // [KL_STAGES]

    // L and U share the mailbox: their tags never coincide (even / odd), and both advance once per apply
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(a.sync + (UPPER ? 1 : 0)) + 1u;
    const unsigned tag = 2u * epoch + (UPPER ? 1u : 0u);
    const int nsteps = a.nsteps;               // local steps of one pencil (multiple of 4)
    const int nsteps_cta = a.nsteps_cta;       // nsteps + largest pencil delay + 4 (last outgoing faces), rounded up to a multiple of 8
    if (threadIdx.x < KL_STAGES) kl_mbar_init(full + threadIdx.x, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned gbase = 0;               // steps marched by this CTA in earlier groups (multiple of 8): ring slot = (gbase + t) & (KL_STAGES - 1)
    const KlGeo geoU{LX, LYR, R, GA, GB, a.gpx, nsteps_cta};

    for (int gi_ = blockIdx.x; gi_ < a.ngroups; gi_ += gridDim.x, gbase += (unsigned)nsteps_cta) {
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
            const int la = lane % LX, lbp = lane / LX;
            const int offset = LX * wa + LYR * wb;                     // this pencil's delay inside the group
            int t_lo[R]; unsigned span[R]; int rr_off[R];
            double* outp[R]; long long ostep;
            ostep = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int qb = R * lbp + r, sk = la + qb;
                const KlRow g = kl_row<UPPER, LX, LYR>(a, Pa, Pb, la, qb);
                t_lo[r] = g.c_lo + sk;
                span[r] = (unsigned)(g.c_hi - g.c_lo);
                rr_off[r] = (0 - sk) & RM;                             // delay-line slot of this row at local step 0
                if (!UPPER) {
                    // this row's place in the U solve's stream 4 at c = -sk (one step = one slot of the group's step sequence)
                    const long long s0 = (long long)kl_slot(geoU, 5, a.nx - 1 - g.gi, a.ny - 1 - g.gj, a.nz - 1);    // c = 0
                    ostep = -(long long)NCW * 5 * R32;                 // c + 1  ->  U step - 1
                    outp[r] = a.out + s0 + 4 * R32 + ostep * (long long)(0 - sk);
                } else {
                    ostep = 0;
                    outp[r] = nullptr;                                 // U: the mover warp stores the solution (un-skewed, coalesced)
                }
            }
            double* oring = rr + ((size_t)warp * R) * (SH::OD * 32) + lane;    // U: [R][OD][32], slot = local step & (OD-1)
            const double* fa_in = faceA + (wa * GB + wb) * LYR + R * lbp;        // read by la == 0 (+ parity * FA)
            double* fa_out = faceA + ((wa + 1) * GB + wb) * LYR + R * lbp;       // written by la == LX-1
            const double* fb_in = faceB + (wb * GA + wa) * LX + la;              // read by lbp == 0, row 0 (+ parity * FB)
            double* fb_out = faceB + ((wb + 1) * GA + wa) * LX + la;             // written by lbp == LY-1, row R-1
            const double* rr_w = rr + ((size_t)warp * R) * (RD * 32) + lane;     // + r * RD*32 + slot*32
            double s[R], vA[R], vB[R], vC[R], dg[R], rh[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { s[r] = 0.0; vA[r] = vB[r] = vC[r] = rh[r] = 0.0; dg[r] = 1.0; }
            __syncthreads();                              // the helpers' step-0 face values and the loader's first slabs are in place
            unsigned long long tr_entry = 0ull, tr_first = 0ull;
            if (a.trace) tr_entry = kl_gtime();
            for (int t0 = 0; t0 < nsteps_cta; t0 += 4) {
                const int tl0 = t0 - offset;
                const bool in_range = valid && tl0 >= 0 && tl0 < nsteps;      // warp-uniform, same for the 4 steps
                if (a.trace && tl0 == 0) tr_first = kl_gtime();
                const unsigned gt = gbase + (unsigned)t0;
                const unsigned quad = (gt >> 2) & (KL_STAGES / 4 - 1), ph = (gt >> KL_STAGE_LOG) & 1u;
                (void)ph;
                const double* cw = cst + warp * (NS * R32) + lane;
                const double* cb = cw + (size_t)quad * (4 * STEP);
                const double* cb_next = cw + (size_t)((quad + 1) & (KL_STAGES / 4 - 1)) * (4 * STEP);
                // operands of local step tl out of the stage at cp / the rhs delay line.  The loader warp has seen the stage's
                // mbarrier complete (and the slab's cp.async group land) before the previous step barrier.
                auto load_ops = [&](const double* cp, int tl) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        vA[r] = cp[r * 32]; vB[r] = LY > 1 ? cp[R32 + r * 32] : 0.0; vC[r] = cp[2 * R32 + r * 32];
                        dg[r] = UPPER ? cp[3 * R32 + r * 32] : 1.0;
                        if (UPPER) rh[r] = cp[4 * R32 + r * 32];
                        else rh[r] = rr_w[r * (RD * 32) + ((rr_off[r] + tl) & RM) * 32];
                    }
                };
                if (in_range && tl0 == 0) load_ops(cb, 0);                      // first step of this pencil (one exposed shared-memory latency)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (in_range) {
                        const int tl = tl0 + u;
                        const int par_r = (u & 1) ^ 1, par_w = u & 1;          // t0 is a multiple of 4
                        double ya[R], yb[R], v[R];
#pragma unroll
                        for (int r = 0; r < R; ++r) ya[r] = __shfl_up_sync(0xffffffffu, s[r], 1);
                        if (LY > 1) {
                            yb[0] = __shfl_up_sync(0xffffffffu, s[R - 1], LX);
#pragma unroll
                            for (int r = 1; r < R; ++r) yb[r] = s[r - 1];
                        }
                        if (!(a.dbg & 8)) {
                        if (la == 0) {
#pragma unroll
                            for (int r = 0; r < R; ++r) ya[r] = fa_in[par_r * FA + r];
                        }
                        if (LY > 1 && lbp == 0) yb[0] = fb_in[par_r * FB];
                        }
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            if (!UPPER) { v[r] = rh[r] - vC[r] * s[r]; if (LY > 1) v[r] = v[r] - vB[r] * yb[r]; v[r] = v[r] - vA[r] * ya[r]; }          // ascending column: c-1, b-1, a-1
                            else { v[r] = rh[r] - vA[r] * ya[r]; if (LY > 1) v[r] = v[r] - vB[r] * yb[r]; v[r] = v[r] - vC[r] * s[r]; v[r] = v[r] * dg[r]; }   // a+1, b+1, c+1, then 1/u_ii
                        }
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const bool act = (unsigned)(tl - t_lo[r]) < span[r];
                            s[r] = act ? v[r] : 0.0;
                            if (!UPPER) { if (act && !(a.dbg & 1)) *outp[r] = s[r]; outp[r] += ostep; }
                            else oring[r * (SH::OD * 32) + (tl & (SH::OD - 1)) * 32] = s[r];
                        }
                        if (!(a.dbg & 8)) {
                        if (la == LX - 1) {
#pragma unroll
                            for (int r = 0; r < R; ++r) fa_out[par_w * FA + r] = s[r];
                        }
                        if (LY > 1 && lbp == LY - 1) fb_out[par_w * FB] = s[R - 1];
                        }
                        // operands of the next step: their latency hides behind the barrier
                        if (tl + 1 < nsteps) load_ops(u < 3 ? cb + (u + 1) * STEP : cb_next, tl + 1);
                    }
                    __syncthreads();                      // step barrier: faces written in step t are read in step t + 1
                }
            }
            if (a.trace && valid && lane == 0) {
                unsigned long long* q = a.trace + ((size_t)(UPPER ? a.npencils : 0) + pencil) * 4;
                q[0] = tr_entry; q[1] = tr_first; q[2] = kl_gtime();
            }
        } else if (warp < NCW + 2 * NHW) {
            // ================= helper warps: the group's L2 faces (first NHW warps: incoming packets, next NHW: outgoing) =================
            const bool inbound = warp < NCW + NHW;
            const int line = (warp - NCW - (inbound ? 0 : NHW)) * 32 + lane;
            const bool isA = line < NA, live = line < NA + NB;
            int wa_c = 0, wb_c = 0, la_c = 0, qb_c = 0, wa_p = 0, wb_p = 0, la_p = 0, qb_p = 0, fidx = 0;
            if (isA) { wb_c = wb_p = line / LYR; qb_c = qb_p = line % LYR; wa_p = GA - 1; la_p = LX - 1; fidx = qb_c; }
            else if (live) { const int q = line - NA; wa_c = wa_p = q / LX; la_c = la_p = q % LX; wb_p = GB - 1; qb_p = LYR - 1; fidx = LYR + la_c; }
            const int fstride = isA ? FA : FB;
            if (inbound) {
                const int Pa_c = Pa0 + wa_c, Pb_c = Pb0 + wb_c;
                const KlRow gc = kl_row<UPPER, LX, LYR>(a, Pa_c, Pb_c, la_c, qb_c);
                const bool has_in = live && Pa_c < a.px && Pb_c < a.py && (isA ? Pa_c > 0 : Pb_c > 0);
                const int skew_c = isA ? qb_c : la_c;
                int pk_lo = 0, pk_hi = 0;                      // consumer-local steps at which an incoming packet exists
                if (has_in) { pk_lo = (isA ? gc.cA_lo : gc.cB_lo) + skew_c; pk_hi = gc.c_hi + skew_c; if (pk_hi < pk_lo) pk_hi = pk_lo; }
                const unsigned in_span = (unsigned)(pk_hi - pk_lo);
                const int off_c = LX * wa_c + LYR * wb_c;
                const ulonglong2* mail_in = a.mail + (size_t)(Pa_c + a.px * Pb_c) * nsteps * FACES + fidx;      // + tl * FACES
                double* in_slot = isA ? faceA + wb_c * LYR + qb_c : faceB + wa_c * LX + la_c;                     // slot 0 (+ parity * FA/FB)
                // keep the distance: start only when the packets of the group's step `lag` are there (lines whose first
                // packet is needed later than that do not hold the group back)
                {
                    const int tw = a.lag - off_c;
                    if ((unsigned)(tw - pk_lo) < in_span) (void)kl_pkt_wait(mail_in + (ptrdiff_t)tw * FACES, tag, a.err);
                }
                // packet of compute step t' lives at consumer-local step t' - off_c; ring slot t' & (PD - 1)
                ulonglong2 pk[PD];
                unsigned tr_stalls = 0u;
                {
                    double v0 = 0.0;
                    const int tl = 0 - off_c;
                    if ((unsigned)(tl - pk_lo) < in_span) v0 = kl_pkt_value(kl_pkt_wait(mail_in + (ptrdiff_t)tl * FACES, tag, a.err));
                    if (live) in_slot[1 * fstride] = v0;       // step 0 reads parity (0 - 1) & 1
#pragma unroll
                    for (int j = 1; j <= PD; ++j) {
                        const int tj = j - off_c;
                        pk[j & (PD - 1)] = make_ulonglong2(0ull, 0ull);
                        if ((unsigned)(tj - pk_lo) < in_span) pk[j & (PD - 1)] = kl_pkt_load(mail_in + (ptrdiff_t)tj * FACES);
                    }
                }
                int din = 1 - off_c - pk_lo;                                       // (tl - pk_lo) of the packet of compute step t + 1, at t = 0
                const ulonglong2* pin = mail_in + (ptrdiff_t)(1 - off_c) * FACES;   // that packet
                __syncthreads();                              // pairs with the compute warps' initial barrier
                for (int t0 = 0; t0 < nsteps_cta; t0 += PD) {
#pragma unroll
                    for (int u = 0; u < PD; ++u) {
                        // incoming value of compute step t + 1 -> parity t & 1, then request the packet of step t + 1 + PD
                        double v = 0.0;
                        if ((unsigned)din < in_span) {
                            ulonglong2 q = pk[(u + 1) & (PD - 1)];
                            if (!kl_pkt_ok(q, tag)) {
                                // the group caught up with its producer: wait, then request the following packets again (the
                                // copies in the ring were loaded before their packets existed; the wait has restored the distance)
                                q = kl_pkt_wait(pin, tag, a.err);
                                ++tr_stalls;
#pragma unroll
                                for (int j = 1; j < PD; ++j)
                                    if ((unsigned)(din + j) < in_span) pk[(u + 1 + j) & (PD - 1)] = kl_pkt_load(pin + j * FACES);
                            }
                            v = kl_pkt_value(q);
                        }
                        if (live) in_slot[(u & 1) * fstride] = v;
                        if ((unsigned)(din + PD) < in_span) pk[(u + 1) & (PD - 1)] = kl_pkt_load(pin + PD * FACES);
                        ++din; pin += FACES;
                        __syncthreads();
                    }
                }
                if (a.trace) {      // blocking packet waits of this group's helper lanes -> slot 3 of the group's first pencil
                    const unsigned st_all = __reduce_add_sync(0xffffffffu, tr_stalls);
                    if (lane == 0 && Pa0 < a.px && Pb0 < a.py) atomicAdd(a.trace + ((size_t)(UPPER ? a.npencils : 0) + Pa0 + a.px * Pb0) * 4 + 3, (unsigned long long)st_all);
                }
            } else {
                const int Pa_p = Pa0 + wa_p, Pb_p = Pb0 + wb_p;
                const KlRow gp = kl_row<UPPER, LX, LYR>(a, Pa_p, Pb_p, la_p, qb_p);
                const bool has_out = live && Pa_p < a.px && Pb_p < a.py && (isA ? Pa_p + 1 < a.px : Pb_p + 1 < a.py);
                int tp_lo = 0, tp_hi = 0;                      // producer-local steps at which an outgoing value exists
                if (has_out) { tp_lo = gp.c_lo + la_p + qb_p; tp_hi = gp.c_hi + la_p + qb_p; if (tp_hi < tp_lo) tp_hi = tp_lo; }
                const unsigned out_span = (unsigned)(tp_hi - tp_lo);
                const int off_p = LX * wa_p + LYR * wb_p;
                const int depth = isA ? LX - 1 : LYR - 1;
                // the value computed in step t - 1 (producer-local step tlp = t - 1 - off_p) is the consumer's packet of step tlp - depth
                ulonglong2* pout = a.mail + ((size_t)(Pa_p + a.px * Pb_p + (isA ? 1 : a.px)) * nsteps) * FACES + fidx + (ptrdiff_t)(-1 - off_p - depth) * FACES;
                const double* out_slot = isA ? faceA + (GA * GB + wb_p) * LYR + qb_p : faceB + (GB * GA + wa_p) * LX + la_p;   // slot GA / GB
                int dout = -1 - off_p - tp_lo;                 // (tlp - tp_lo) at t = 0
                __syncthreads();                              // initial barrier
                for (int t0 = 0; t0 < nsteps_cta; t0 += 2) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if ((unsigned)dout < out_span) kl_pkt_store(pout, out_slot[(u ^ 1) * fstride], tag);      // parity (t - 1) & 1
                        ++dout; pout += FACES;
                        __syncthreads();
                    }
                }
            }
        } else if (warp == NCW + 2 * NHW) {
            // ================= stage warp: coefficient stages by TMA, mbarrier waits on behalf of everybody =================
            unsigned long long pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            const double* gsrc = a.coef + (size_t)group * nsteps_cta * STEP;
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < KL_STAGES; ++j)
                    if (j < nsteps_cta) {
                        const unsigned sl = (gbase + (unsigned)j) & (KL_STAGES - 1);
                        kl_mbar_wait(full + sl, ((gbase + (unsigned)j - KL_STAGES) >> KL_STAGE_LOG) & 1u);      // the slot's previous copy has landed (nobody may have waited for it)
                        kl_bulk(cst + (size_t)sl * STEP, gsrc + (size_t)j * STEP, STEP_BYTES, full + sl, pol);
                    }
            }
            // the compute warps read a stage one step before they use it and never touch the mbarriers: this warp sees the
            // stage of step t + 2 complete before it joins the barrier of step t
            kl_mbar_wait(full + (gbase & (KL_STAGES - 1)), (gbase >> KL_STAGE_LOG) & 1u);
            kl_mbar_wait(full + ((gbase + 1u) & (KL_STAGES - 1)), ((gbase + 1u) >> KL_STAGE_LOG) & 1u);
            __syncthreads();                              // initial barrier
            for (int t = 0; t < nsteps_cta; ++t) {
                // the slot of step t - 1 is free: stage step t - 1 + KL_STAGES
                if (lane == 0 && t >= 1 && t - 1 + KL_STAGES < nsteps_cta && !(a.dbg & 16)) {
                    const unsigned sl = (gbase + (unsigned)(t - 1)) & (KL_STAGES - 1);      // (this warp saw its phase complete three steps ago)
                    kl_bulk(cst + (size_t)sl * STEP, gsrc + (size_t)(t - 1 + KL_STAGES) * STEP, STEP_BYTES, full + sl, pol);
                }
                if (t + 2 < nsteps_cta && !(a.dbg & 16)) kl_mbar_wait(full + ((gbase + (unsigned)(t + 2)) & (KL_STAGES - 1)), ((gbase + (unsigned)(t + 2)) >> KL_STAGE_LOG) & 1u);
                __syncthreads();
            }
        } else {
            // ================= mover warp: L - rhs slabs into the delay line; U - finished solution slabs out, both un-skewed =================
            // lane (la,lb') moves the elements of lane (la,lb') of every pencil: one slab c per pencil and step, 64-byte segments
            const int la = lane % LX, lbp = lane / LX;
            constexpr int NK = NCW * R;
            const double* gp[NK]; int cw[NK]; unsigned cs[NK]; unsigned sbase[NK]; long long rstep = 0;
#pragma unroll
            for (int w = 0; w < NCW; ++w)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int k = w * R + r;
                    const KlRow g = kl_row<UPPER, LX, LYR>(a, Pa0 + w % GA, Pb0 + w / GA, la, R * lbp + r);
                    const int off_w = LX * (w % GA) + LYR * (w / GA);
                    // the slab of pencil w handled at step t:  L: c = t - off_w + D (prefetch);  U: c = t - 1 - off_w - SK (complete since step t - 1)
                    const int cshift = UPPER ? -1 - off_w - SK : D - off_w;
                    cw[k] = cshift - g.c_lo;                             // c - c_lo at t = 0
                    cs[k] = (unsigned)(g.c_hi - g.c_lo);
                    gp[k] = (UPPER ? (const double*)a.out : a.rhs) + (g.rbase + g.rstep * (long long)cshift);
                    sbase[k] = kl_u32(rr + (size_t)k * ((UPPER ? SH::OD : RD) * 32) + lane);
                    rstep = g.rstep;
                }
            if (!UPPER) {
                // prologue: steps -D .. -1
                for (int t = -D; t < 0; ++t) {
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        const int c = t + D - (LX * ((k / R) % GA) + LYR * ((k / R) / GA));
                        if ((unsigned)(cw[k] + t) < cs[k]) kl_cp8s(sbase[k] + (unsigned)(c & RM) * 256u, gp[k] + rstep * t);
                    }
                    kl_cp_commit();
                }
                kl_cp_wait<D - 2 < 0 ? 0 : D - 2>();      // the slabs of steps 0 and 1 have landed
            }
            __syncthreads();                              // initial barrier
            for (int t = 0; t < nsteps_cta; ++t) {
                if (!UPPER) {
                    if (!(a.dbg & 2)) {
#pragma unroll
                        for (int k = 0; k < NK; ++k) {
                            const int c = t + D - (LX * ((k / R) % GA) + LYR * ((k / R) / GA));
                            if ((unsigned)(cw[k] + t) < cs[k]) kl_cp8s(sbase[k] + (unsigned)(c & RM) * 256u, gp[k]);
                            gp[k] += rstep;
                        }
                        kl_cp_commit();
                        kl_cp_wait<D - 2 < 0 ? 0 : D - 2>();      // the slabs consumed in step t + 2 have landed
                    }
                } else if (!(a.dbg & 1)) {
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        // element of this lane's row in slab c: written at the row's local step c + sk
                        const int tlw = t - 1 - (LX * ((k / R) % GA) + LYR * ((k / R) / GA)) - SK + la + R * lbp + (k % R);
                        if ((unsigned)(cw[k] + t) < cs[k]) {
                            double v;
                            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sbase[k] + (unsigned)(tlw & (SH::OD - 1)) * 256u) : "memory");
                            *const_cast<double*>(gp[k]) = v;
                        }
                        gp[k] += rstep;
                    }
                }
                __syncthreads();
            }
            if (!UPPER) kl_cp_wait<0>();
        }
        // (the next group's face clearing is ordered after every read of this group by the last step barrier)
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
void kb_lean_free(KbLean* m) {
    if (!m) return;
    KB_FREE(m->coef[0]); KB_FREE(m->coef[1]);
    KB_FREE(m->order); KB_FREE(m->mail); KB_FREE(m->sync); KB_FREE(m->trace);
    delete m;
}

struct KlKernel { kl_fn fn; size_t smem; int threads; };
template <bool UPPER, int LX, int LY, int R, int GA, int GB, int RD>
static KlKernel kl_make() { typedef KlShape<UPPER, LX, LY, R, GA, GB, RD> SH; return KlKernel{kb_trsv_lean<UPPER, LX, LY, R, GA, GB, RD>, SH::SMEM, SH::THREADS}; }
// shapes: 3-D pencils of 8 x 4 lanes with R = 1, 2 rows per lane in groups of g x g; 2-D pencils of 32 lanes in groups of g x 1
static bool kl_kernel(bool upper, bool two_d, int r, int g, KlKernel* k) {
    if (two_d) {
        if (g == 1) *k = upper ? kl_make<true, 32, 1, 1, 1, 1, 64>() : kl_make<false, 32, 1, 1, 1, 1, 64>();
        else if (g == 2) *k = upper ? kl_make<true, 32, 1, 1, 2, 1, 64>() : kl_make<false, 32, 1, 1, 2, 1, 64>();
        else *k = upper ? kl_make<true, 32, 1, 1, 4, 1, 64>() : kl_make<false, 32, 1, 1, 4, 1, 64>();
        return true;
    }
    if (r == 1) {
        if (g == 1) *k = upper ? kl_make<true, 8, 4, 1, 1, 1, 32>() : kl_make<false, 8, 4, 1, 1, 1, 32>();
        else *k = upper ? kl_make<true, 8, 4, 1, 2, 2, 32>() : kl_make<false, 8, 4, 1, 2, 2, 32>();
        return true;
    }
    if (g == 1) *k = upper ? kl_make<true, 8, 4, 2, 1, 1, 32>() : kl_make<false, 8, 4, 2, 1, 1, 32>();
    else *k = upper ? kl_make<true, 8, 4, 2, 2, 2, 32>() : kl_make<false, 8, 4, 2, 2, 2, 32>();
    return true;
}

// gx, gy, gz: the box grid detected from the factor's pattern (kb_trsv_tiles.cu).  *out stays nullptr (KB_OK) when the
// pattern is not a full stencil.
int kb_lean_build(kb_pc_s* pc, int gx, int gy, int gz, unsigned* d_err, KbLean** out) {
    *out = nullptr;
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    const int n = (int)A->n;
    const bool two_d = gz == 1;
    KbLean* m = new KbLean;
    m->n = n; m->err = d_err;
    int r = two_d ? 1 : 2, g = two_d ? 4 : 2;
    if (getenv("KB_LEAN_ROWS")) { const int e = atoi(getenv("KB_LEAN_ROWS")); if (!two_d && (e == 1 || e == 2)) r = e; }
    if (getenv("KB_MARCH_GROUP")) { const int e = atoi(getenv("KB_MARCH_GROUP")); if (e == 1 || e == 2 || (e == 4 && two_d)) g = e; }
    if (getenv("KB_MARCH_LAG")) m->lag = std::max(KL_PD + 1, atoi(getenv("KB_MARCH_LAG")));
    if (getenv("KB_LEAN_DBG")) m->dbg = atoi(getenv("KB_LEAN_DBG"));
    m->r = r;
    if (two_d) { m->nx = gx; m->ny = 1; m->nz = gy; m->lx = 32; m->lyr = 1; m->faces = 1; }
    else { m->nx = gx; m->ny = gy; m->nz = gz; m->lx = 8; m->lyr = 4 * r; m->faces = m->lx + m->lyr; }
    m->ga = g; m->gb = two_d ? 1 : g;
    m->px = (m->nx + m->lx - 1) / m->lx; m->py = (m->ny + m->lyr - 1) / m->lyr;
    m->nsteps = (m->nz + m->lx + m->lyr - 2 + 3) & ~3;          // (padded steps: zero coefficients, no rows)
    m->nsteps_cta = (m->nsteps + m->lx * (m->ga - 1) + m->lyr * (m->gb - 1) + 4 + 7) & ~7;      // (+ one block for the last outgoing faces)
    m->gpx = (m->px + m->ga - 1) / m->ga; m->gpy = (m->py + m->gb - 1) / m->gb;
    m->ngroups = m->gpx * m->gpy;
    const long long np = (long long)m->px * m->py;
    const long long slots = np * m->nsteps * m->faces;
    const long long per_stream = (long long)m->ngroups * m->nsteps_cta * (m->ga * m->gb) * r * 32;
    if (np > (1 << 24) || slots * 16 > (4ll << 30) || per_stream * 5 * 8 > (24ll << 30)) { delete m; return KB_OK; }
    m->npencils = (int)np;
    int st = KB_OK;
    int* d_bad = nullptr;
    do {
        for (int u = 0; u < 2 && st == KB_OK; ++u) {
            const size_t cnt = (size_t)per_stream * (u ? 5 : 3) + 64;
            st = kb_alloc(&m->coef[u], cnt);
            if (st == KB_OK) cudaMemsetAsync(m->coef[u], 0, cnt * sizeof(double), c->stream);
        }
        if (st != KB_OK || (st = kb_alloc(&d_bad, 1)) != KB_OK) break;
        cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream);
        const int sB = two_d ? 0 : gx, sC = two_d ? gx : gx * gy;
        const KlGeo geo{m->lx, m->lyr, r, m->ga, m->gb, m->gpx, m->nsteps_cta};
        {
            KbLaunch L(c, KB_K_OTHER);
            k_lean_skew<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(pc->l_rp, pc->l_col, pc->diag_ptr, pc->lu, pc->inv_diag, n, gx, gy, sB, sC, m->nx, m->ny,
                                                                          m->nz, geo, m->coef[0], m->coef[1], d_bad);
        }
        int h_bad = 0;
        if (cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        if (h_bad) { cudaFree(d_bad); kb_lean_free(m); return KB_OK; }      // not a full stencil: other kernels take it
        std::vector<int> ord((size_t)m->ngroups), lev((size_t)m->ngroups);
        for (int b = 0, id = 0; b < m->gpy; ++b) for (int a2 = 0; a2 < m->gpx; ++a2, ++id) { ord[id] = id; lev[id] = a2 * m->lx * m->ga + b * m->lyr * m->gb; }
        std::stable_sort(ord.begin(), ord.end(), [&](int p, int q) { return lev[p] < lev[q]; });
        if ((st = kb_alloc(&m->order, (size_t)m->ngroups)) != KB_OK) break;
        if (cudaMemcpyAsync(m->order, ord.data(), (size_t)m->ngroups * sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        if ((st = kb_alloc(&m->mail, (size_t)(slots + (long long)(m->px + 2) * m->nsteps * m->faces))) != KB_OK) break;
        cudaMemsetAsync(m->mail, 0, (size_t)slots * sizeof(ulonglong2), c->stream);     // tag 0 is never used
        if ((st = kb_alloc(&m->sync, 4)) != KB_OK) break;
        if (getenv("KB_MARCH_TRACE") && (st = kb_alloc(&m->trace, (size_t)np * 8)) != KB_OK) break;
        cudaMemsetAsync(m->sync, 0, 4 * sizeof(unsigned), c->stream);
        // all CTAs co-resident (a group may wait on a group of any other CTA); groups beyond the grid are picked up in level order
        int cap = 1 << 30;
        for (int u = 0; u < 2 && st == KB_OK; ++u) {
            KlKernel k{};
            kl_kernel(u == 1, two_d, r, g, &k);
            m->fn[u] = k.fn; m->smem[u] = k.smem; m->threads = k.threads;
            if (cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k.fn, k.threads, k.smem) != cudaSuccess || occ < 1) { st = KB_SOLVE_ERROR; break; }
            cap = std::min(cap, occ * c->sm_count);
        }
        if (st != KB_OK) break;
        m->grid = std::max(1, std::min(cap, m->ngroups));
        if (getenv("KB_MARCH_GRID")) m->grid = std::max(1, std::min(m->grid, atoi(getenv("KB_MARCH_GRID"))));
        if (m->trace) fprintf(stderr, "[kb lean] rows/lane %d group %dx%d threads %d smem %zu/%zu grid %d (cap %d) groups %d nsteps %d (cta %d) lag %d\n", r, m->ga, m->gb,
                              m->threads, m->smem[0], m->smem[1], m->grid, cap, m->ngroups, m->nsteps, m->nsteps_cta, m->lag);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
    } while (0);
    if (d_bad) cudaFree(d_bad);
    if (st != KB_OK) { cudaGetLastError(); kb_set_error("ilu0: pencil-march schedule setup failed"); kb_lean_free(m); return st; }
    *out = m;
    return KB_OK;
}

int kb_lean_apply(kb_pc_s* pc, KbLean* m, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask) {
    kb_ctx_s* c = pc->a->ctx;
    KlArgs a{};
    a.n = m->n; a.nx = m->nx; a.ny = m->ny; a.nz = m->nz; a.px = m->px; a.py = m->py; a.npencils = m->npencils;
    a.nsteps = m->nsteps; a.nsteps_cta = m->nsteps_cta; a.lag = m->lag; a.dbg = m->dbg;
    a.gpx = m->gpx; a.gpy = m->gpy; a.ngroups = m->ngroups;
    a.order = m->order; a.mail = m->mail; a.sync = m->sync; a.err = m->err; a.skip_ctl = skip_ctl; a.skip_mask = skip_mask; a.trace = m->trace;
    if (m->trace) cudaMemsetAsync(m->trace, 0, (size_t)m->npencils * 8 * sizeof(unsigned long long), c->stream);
    for (int u = 0; u < 2; ++u) {
        a.coef = m->coef[u];
        a.rhs = u == 0 ? d_r : nullptr; a.out = u == 0 ? m->coef[1] : d_z;
        KbLaunch L(c, KB_K_TRSV);
        m->fn[u]<<<m->grid, m->threads, m->smem[u], c->stream>>>(a);
    }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// diagnostics (not part of the ABI header): copy the per-pencil timeline of the last apply; returns pencils per solve
int kb_lean_trace_get(KbLean* m, unsigned long long* out, int* px, int* py) {
    if (!m || !m->trace) return 0;
    cudaMemcpy(out, m->trace, (size_t)m->npencils * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    *px = m->px; *py = m->py;
    return m->npencils;
}
