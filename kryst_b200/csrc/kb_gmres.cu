// kb_gmres.cu — restarted GMRES(m), device-resident (replaces src/solver/gmres.rs:216-402).
//
// Orthogonalisation is classical Gram-Schmidt with reorthogonalisation done as block GEMVs over the
// Krylov basis (north_star item 3) instead of the reference's 2(j+1) sequential dot+axpy pairs
// (gmres.rs:83-96):   h1 = V^T w ; w -= V h1 ; h2 = V^T w ; w -= V h2 (fused with ||w||^2) ; H[:,j] = h1 + h2.
// The basis is column-major V[ld x (m+1)] (each column contiguous, with ghost space on a shard).
// Givens rotations (gmres.rs:154-176), the stop test (Convergence::check, :348-354), back-substitution
// (:180-192) and the per-cycle true-residual test (:388-398, strict <) run on the device in single-block
// epilogues, so a whole restart cycle is one CUDA-graph replay with no host round trip.
// Preconditioned modes use the textbook formulation (Tier T, SURVEY §8c: the reference's Left/Right
// branches are mathematically inconsistent, F7): Left = Arnoldi on M^-1 A started from M^-1 r0 with the
// inner test on ||M^-1 r||; Right = Arnoldi on A M^-1, x += M^-1 (V y).
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include "kb_objects.h"
#include "kb_epilogue.cuh"
#include "kb_driver.cuh"

#define KB_GM_EPS 1e-14     // gmres.rs:233

struct KbGmresDev {       // device pointers shared by the kernels
    KbCtl* ctl;
    double* V; size_t ld;
    const double* h1src; const double* h2src;   // where the (all-reduced) CGS coefficients live
    int flex;                                   // FGMRES (fgmres.rs): single-pass CGS, z_j = M^-1 v_j kept, literal quirks
    double* Z;                                  // flexible basis (== V when there is no preconditioner)
    int block;                                  // KB_FLAG_BLOCK_ORTH: one block classical GS pass, ||w'||^2 = w.w - sum h^2, ONE reduction per step
};
#define KB_SIDE_FLEX 3

// ---- multi-column dot: partial[c][tile] = canonical tile sum of V_c . w, c = 0..j ----------------------
// self_col != 0: one more "column" w . w (block orthogonalisation: the norm rides the same reduction)
__global__ void __launch_bounds__(KB_THREADS) kb_gs_dot(KbGmresDev g, const double* __restrict__ w, long long n, double* partials, size_t pstride,
                                                        int ncols, int self_col = 0) {
    KbCtl* ctl = g.ctl;
    if (ctl->done || ctl->cycle_break) return;
    __shared__ double sm[(KB_MAX_RESTART + 1) * 8];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const long long i = (long long)blockIdx.x * KB_TILE + 2 * tid;
    double w0 = 0.0, w1 = 0.0;
    const bool h0 = i < n, h1 = i + 1 < n;
    if (h1) { double2 t = kb_ld2(w + i); w0 = t.x; w1 = t.y; }
    else if (h0) w0 = w[i];
#pragma unroll 4
    for (int c = 0; c < ncols; ++c) {
        const double* vc = g.V + (size_t)c * g.ld;
        double e0 = 0.0, e1 = 0.0;
        if (h1) { double2 t = kb_ld2(vc + i); e0 = t.x * w0; e1 = t.y * w1; }
        else if (h0) e0 = vc[i] * w0;
        double v = kb_warp_butterfly(e0 + e1);
        if (lane == 0) sm[c * 8 + wp] = v;
    }
    if (self_col) {
        double v = kb_warp_butterfly(w0 * w0 + w1 * w1);
        if (lane == 0) sm[ncols * 8 + wp] = v;
        ++ncols;
    }
    __syncthreads();
    for (int c = tid; c < ncols; c += KB_THREADS) {
        double s = sm[c * 8];
#pragma unroll
        for (int k = 1; k < 8; ++k) s = s + sm[c * 8 + k];
        partials[(size_t)c * pstride + blockIdx.x] = s;
    }
}
// ---- CGS2 sweep 2 as ONE pass over the basis: w -= V h1 and the tile sums of h2 = V^T w_new ----------------------
// (north_star kernel 3 / SURVEY K6 "update_dotV".)  A register-tile version could not overlap its loads with its
// arithmetic (1 CTA/SM, every warp in the same phase; removed); a first shared-memory version staged half tiles with one bulk copy
// per 2-KB column slice and paid more for ~32 bulk-copy issues per stage than for the stage's HBM time (0.89 ms per
// sweep against 0.34 + 0.37 ms for the two plain sweeps).  This version stages with cp.async instead:
//   * thread l of a 128-thread team owns rows 2l, 2l+1 of a HALF tile (256 rows = canonical lanes) and copies exactly
//     its own 16 bytes of w and of every basis column into its own shared-memory slots (cp.async.cg, 16 B) - the data a
//     thread consumes is the data it copied, so cp.async.wait_group is the only synchronisation of the pipeline;
//   * stages form a ring NST deep (NST-1 half tiles in flight per team, 2 teams per CTA working on alternate tiles),
//     sized at launch to the shared memory of the SM;
//   * the thread runs  t -= V[c]*h1[c]  for c = 0..j, stores w, then forms the products V[c]*t for all columns and
//     reduces them with a TRANSPOSED butterfly: at offset 16/8/4/2/1 a lane keeps half of its columns and trades the
//     other half with its xor partner, so 32 columns cost 31 shuffles instead of 160, and lane k ends up with column
//     k's warp sum.  Each addition pairs the same two lanes' values as the plain xor butterfly (IEEE addition is
//     commutative), hence the tile sums are bit-identical to kb_gs_dot's;
//   * the two halves' 4 + 4 warp sums are added in canonical order (warp 0..7) by one thread per column.
// Traffic per Arnoldi step drops from 4 to 3 sweeps over V (SURVEY 8d byte model).
#define KB_GSF_ROWS 256
#define KB_GSF_TEAM 128
#define KB_GSF_TEAMS 1        // one 128-thread team per CTA; two CTAs share an SM when their rings fit
#define KB_GSF_THREADS (KB_GSF_TEAM * KB_GSF_TEAMS)
#define KB_GSF_MAXC 64
#define KB_GSF_MAX_STAGES 6
struct KbGsfHead {
    double h[KB_GSF_MAXC];
    double wsum[KB_GSF_TEAMS][KB_GSF_MAXC * 8];       // warp sums of the team's current tile: [column][canonical warp 0..7]
};
__device__ __forceinline__ void kb_gsf_cp16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
template <int N>
__device__ __forceinline__ void kb_gsf_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Products of W consecutive basis columns with this thread's two updated rows, reduced over the warp: returns, in lane l, the
// warp sum of column g0 + (l & (W-1)).  log2(32/W) plain xor-butterfly rounds over all W columns, then the transposed rounds
// (a lane keeps half of its columns and trades the other half with its xor partner): every addition pairs the same two lanes'
// values as the canonical 16-8-4-2-1 butterfly, so the sums are bit-identical to kb_gs_dot's whatever W is.  W is chosen from
// the number of columns left, so that short bases do not pay for 32 columns.
template <int W>
__device__ __forceinline__ double kb_gsf_group(const double* st, int g0, int ncols, bool h0, bool h1, double t0, double t1, int lane) {
    double x[W];
    if (h1) {                 // both rows exist (every thread but the tail of the last tile)
#pragma unroll
        for (int k = 0; k < W; ++k) {
            x[k] = 0.0;
            if (g0 + k < ncols) {         // warp-uniform
                const double2 v = *reinterpret_cast<const double2*>(st + (size_t)(g0 + k + 1) * KB_GSF_ROWS);
                x[k] = v.x * t0 + v.y * t1;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < W; ++k) {
            double e0 = 0.0;
            if (g0 + k < ncols && h0) e0 = st[(size_t)(g0 + k + 1) * KB_GSF_ROWS] * t0;
            x[k] = e0 + 0.0;
        }
    }
#pragma unroll
    for (int off = 16; off >= W; off >>= 1) {
#pragma unroll
        for (int k = 0; k < W; ++k) x[k] = x[k] + __shfl_xor_sync(0xffffffffu, x[k], off);
    }
#pragma unroll
    for (int off = W / 2; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
            const double send = hi ? x[k] : x[k + off];
            const double keep = hi ? x[k + off] : x[k];
            x[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return x[0];
}

template <int NST>
__global__ void __launch_bounds__(KB_GSF_THREADS) kb_gs_fused(KbGmresDev g, double* __restrict__ w, const double* __restrict__ hsrc, int n, int ntiles,
                                                                double* partials, size_t pstride, int ncols) {
    if (g.ctl->done || g.ctl->cycle_break) return;
    extern __shared__ __align__(128) unsigned char kb_gsf_raw[];
    KbGsfHead& H = *reinterpret_cast<KbGsfHead*>(kb_gsf_raw);
    const int tid = threadIdx.x, team = tid / KB_GSF_TEAM, lt = tid % KB_GSF_TEAM, lane = tid & 31, wp = lt >> 5;
    const size_t stage_doubles = (size_t)(ncols + 1) * KB_GSF_ROWS;       // [w | V_0 .. V_{ncols-1}] x 256 rows
    double* ring = reinterpret_cast<double*>(kb_gsf_raw + ((sizeof(KbGsfHead) + 127) & ~(size_t)127)) + (size_t)team * NST * stage_doubles + 2 * lt;
    for (int c = tid; c < ncols; c += KB_GSF_THREADS) H.h[c] = hsrc[c];
    __syncthreads();
    // team t walks tiles blockIdx.x*TEAMS + t, + gridDim.x*TEAMS, ... ; two halves per tile
    const int first = (int)blockIdx.x * KB_GSF_TEAMS + team, stride = (int)gridDim.x * KB_GSF_TEAMS;
    const int nhalves = first < ntiles ? 2 * ((ntiles - first + stride - 1) / stride) : 0;
    auto issue = [&](int it) {        // stage half `it` (this thread's 16 bytes of w and of every column)
        if (it < nhalves) {
            const int tile = first + (it >> 1) * stride;
            const long long i = (long long)tile * KB_TILE + (it & 1) * KB_GSF_ROWS + 2 * lt;
            if (i < n) {              // rows i, i+1 (vectors are padded by >= 2, so the pair is always readable)
                double* st = ring + (size_t)(it % NST) * stage_doubles;
                kb_gsf_cp16(st, w + i);
                for (int c = 0; c < ncols; ++c) kb_gsf_cp16(st + (size_t)(c + 1) * KB_GSF_ROWS, g.V + (size_t)c * g.ld + i);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int it = 0; it < NST - 1; ++it) issue(it);
    for (int it = 0; it < nhalves; ++it) {
        const int tile = first + (it >> 1) * stride;
        const int half = it & 1;
        const long long i = (long long)tile * KB_TILE + half * KB_GSF_ROWS + 2 * lt;
        const bool h0 = i < n, h1 = i + 1 < n;
        issue(it + NST - 1);                      // keep NST-1 halves in flight (its slot was consumed in iteration it-1)
        kb_gsf_wait<NST - 1>();                   // half `it` has landed
        const double* st = ring + (size_t)(it % NST) * stage_doubles;
        double t0 = 0.0, t1 = 0.0;
        if (h0) { const double2 t = *reinterpret_cast<const double2*>(st); t0 = t.x; t1 = t.y; }
        // phase 1: w -= V h1 (sequential in c, as GsUpdateOp)
        if (h0) {
#pragma unroll 4
            for (int c = 0; c < ncols; ++c) {
                const double2 v = *reinterpret_cast<const double2*>(st + (size_t)(c + 1) * KB_GSF_ROWS);
                const double h = H.h[c];
                t0 = t0 - v.x * h; t1 = t1 - v.y * h;
            }
        }
        if (h1) kb_st2(w + i, make_double2(t0, t1));
        else if (h0) w[i] = t0;
        // phase 2: products for every column, transposed butterfly per group of 32 / 16 / 8 columns
        for (int g0 = 0; g0 < ncols;) {
            const int rem = ncols - g0;
            if (rem > 24) {
                const double v = kb_gsf_group<32>(st, g0, ncols, h0, h1, t0, t1, lane);
                if (g0 + lane < ncols) H.wsum[team][(g0 + lane) * 8 + half * 4 + wp] = v;
                g0 += 32;
            } else if (rem > 8) {
                const double v = kb_gsf_group<16>(st, g0, ncols, h0, h1, t0, t1, lane);
                if (lane < 16 && g0 + lane < ncols) H.wsum[team][(g0 + lane) * 8 + half * 4 + wp] = v;
                g0 += 16;
            } else {
                const double v = kb_gsf_group<8>(st, g0, ncols, h0, h1, t0, t1, lane);
                if (lane < 8 && g0 + lane < ncols) H.wsum[team][(g0 + lane) * 8 + half * 4 + wp] = v;
                g0 += 8;
            }
        }
        if (half == 1) {
            asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(KB_GSF_TEAM) : "memory");
            for (int c = lt; c < ncols; c += KB_GSF_TEAM) {
                double sum = H.wsum[team][c * 8];
#pragma unroll
                for (int k = 1; k < 8; ++k) sum = sum + H.wsum[team][c * 8 + k];
                partials[(size_t)c * pstride + tile] = sum;
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(KB_GSF_TEAM) : "memory");
        }
    }
    kb_gsf_wait<0>();
}
static size_t kb_gsf_smem(int ncols, int nstages) {
    return ((sizeof(KbGsfHead) + 127) & ~(size_t)127) + (size_t)KB_GSF_TEAMS * nstages * (size_t)(ncols + 1) * KB_GSF_ROWS * sizeof(double);
}
// ring depth: small stages -> up to 4 deep and two CTAs per SM; large stages -> whatever fits one CTA per SM (0: does not fit)
static int kb_gsf_stages(int ncols) {
    const size_t head = (sizeof(KbGsfHead) + 127) & ~(size_t)127;
    const size_t stage = (size_t)(ncols + 1) * KB_GSF_ROWS * sizeof(double);
    if (ncols > KB_GSF_MAXC) return 0;
    const size_t half_sm = 110 * 1024, full_sm = 225 * 1024;
    // the kernel is issue-latency bound at 4 warps per CTA: prefer MORE resident CTAs (two stages each) over deeper rings
    static const int occ_env = getenv("KB_GSF_OCC") ? atoi(getenv("KB_GSF_OCC")) : 1;
    if (occ_env && 3 * (2 * stage + head + 1024) <= full_sm + 2048) return 2;
    if (2 * stage + head <= half_sm) { const size_t k = (half_sm - head) / stage; return (int)(k > 4 ? 4 : k); }
    const size_t k = (full_sm - head) / stage;
    return k < 2 ? 0 : (int)(k > KB_GSF_MAX_STAGES ? KB_GSF_MAX_STAGES : k);
}
typedef void (*kb_gsf_fn)(KbGmresDev, double*, const double*, int, int, double*, size_t, int);
static kb_gsf_fn kb_gsf_kernel(int nst) {
    switch (nst) {
    case 2: return kb_gs_fused<2>;
    case 3: return kb_gs_fused<3>;
    case 4: return kb_gs_fused<4>;
    case 5: return kb_gs_fused<5>;
    default: return kb_gs_fused<6>;
    }
}

// level 2 for every column in parallel: block c reduces partial[c][0..P)
__global__ void __launch_bounds__(KB_THREADS) kb_gs_level2(KbCtl* ctl, const double* partials, size_t pstride, int P, double* dst) {
    if (ctl->done || ctl->cycle_break) return;
    __shared__ double sm[8];
    double s = kb_level2(partials + (size_t)blockIdx.x * pstride, P, sm);
    if (threadIdx.x == 0) dst[blockIdx.x] = s;
}

// ---- Arnoldi step epilogue: H column, Givens, stop test (all threads of one block) ----------------------
__device__ void kb_arnoldi_fin(const KbGmresDev& g, double ww, int j) {
    KbCtl* c = g.ctl;
    __shared__ double hcol[KB_MAX_RESTART + 2], scs[KB_MAX_RESTART], ssn[KB_MAX_RESTART];
    if (g.flex || g.block) { for (int k = threadIdx.x; k <= j; k += blockDim.x) hcol[k] = g.h1src[k]; }      // one classical GS pass (fgmres.rs:219-228)
    else { for (int k = threadIdx.x; k <= j; k += blockDim.x) hcol[k] = g.h1src[k] + g.h2src[k]; }
    for (int k = threadIdx.x; k < j; k += blockDim.x) { scs[k] = c->cs[k]; ssn[k] = c->sn[k]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ldh = c->restart + 1;
        const double hn = sqrt(ww);
        hcol[j + 1] = hn;
        c->hnorm = hn;
        // happy breakdown: gmres.rs:97-101 (|h| < 1e-14) ; fgmres.rs:253-262 (|h| < haptol*|s_j|, v_{j+1} := 0, no break)
        const int happy = g.flex ? (fabs(hn) < 1e-12 * fabs(c->g[j])) : (fabs(hn) < KB_GM_EPS);
        for (int i = 0; i < j; ++i) {                       // gmres.rs:155-159
            const double temp = scs[i] * hcol[i] + ssn[i] * hcol[i + 1];
            hcol[i + 1] = -ssn[i] * hcol[i] + scs[i] * hcol[i + 1];
            hcol[i] = temp;
        }
        const double hkk = hcol[j], hk1k = hcol[j + 1];
        const double r = sqrt(hkk * hkk + hk1k * hk1k);
        double cj, sj;
        if (g.flex ? (r == 0.0) : (fabs(r) < KB_GM_EPS)) { cj = 1.0; sj = 0.0; } else { cj = hkk / r; sj = hk1k / r; }
        hcol[j] = cj * hkk + sj * hk1k;
        hcol[j + 1] = 0.0;
        c->cs[j] = cj; c->sn[j] = sj;
        const double gj = c->g[j], gj1 = c->g[j + 1];
        const double temp = cj * gj + sj * gj1;
        c->g[j + 1] = -sj * gj + cj * gj1;
        c->g[j] = temp;
        for (int i = 0; i <= j + 1; ++i) c->h[i + (size_t)ldh * j] = hcol[i];
        const double res_norm = fabs(c->g[j + 1]);
        if (c->hist_len < c->hist_cap) c->hist[c->hist_len] = res_norm;      // residual_history.push(res_norm) (fgmres.rs:290)
        c->hist_len += 1;
        const unsigned long long it = c->iter + 1;
        c->iter = it;
        const double rel = res_norm / (g.flex ? c->beta_g : c->res0);   // Convergence::check; FGMRES divides by this cycle's s[0] (fgmres.rs:292)
        const int stop = (rel <= c->tol) || (it >= c->max_iters);
        c->res = res_norm; c->converged = stop;
        c->m = j + 1; c->happy = g.flex ? (c->happy | happy) : happy;
        if (g.flex) { c->hflag = happy; if (stop) { c->cycle_break = 1; c->early = 1; } }   // early: "converged" of fgmres.rs:299
        else if (stop || happy) c->cycle_break = 1;
        c->j = j + 1;
    }
}
__global__ void kb_arnoldi_fin_kernel(KbGmresDev g, const double* sums, int j) {
    if (g.ctl->done || g.ctl->cycle_break) return;
    kb_arnoldi_fin(g, sums[0], j);
}
// Block orthogonalisation (the idea of pca_gmres.rs:172-229: every inner product of the step is gathered into one
// array and reduced ONCE): sums = {V_0.w, ..., V_j.w, w.w}; ||w - V h||^2 = w.w - sum_k h_k^2 for an orthonormal V,
// evaluated sequentially in k (mul, then sub) and clamped at 0; then the usual H column / Givens / stop test.
__global__ void kb_block_fin_kernel(KbGmresDev g, const double* sums, int j) {
    if (g.ctl->done || g.ctl->cycle_break) return;
    double hn2 = sums[j + 1];
    for (int k = 0; k <= j; ++k) hn2 = hn2 - sums[k] * sums[k];
    if (hn2 < 0.0) hn2 = 0.0;
    kb_arnoldi_fin(g, hn2, j);
}

// w -= V h (sequential in c: t = t - V[c][i]*h[c]); NORM: fused ||w||^2 and the Arnoldi epilogue
template <bool NORM>
struct GsUpdateOp : KbRedBase {
    static constexpr int NRED = NORM ? 1 : 0;
    KbGmresDev g; double* w; const double* hsrc; double* slots; int ncols; const KbP2PDev* p2p;
    __device__ bool skip() const { return g.ctl->done != 0 || g.ctl->cycle_break != 0; }
    __device__ void pair(long long i, bool has1, double* red) const {
        if (has1) {
            double2 t = kb_ld2(w + i);
#pragma unroll 4
            for (int c = 0; c < ncols; ++c) {
                const double2 v = kb_ld2(g.V + (size_t)c * g.ld + i);
                const double h = hsrc[c];
                t.x = t.x - v.x * h; t.y = t.y - v.y * h;
            }
            kb_st2(w + i, t);
            if (NORM) red[0] = t.x * t.x + t.y * t.y;
        } else {
            double t = w[i];
            for (int c = 0; c < ncols; ++c) t = t - g.V[(size_t)c * g.ld + i] * hsrc[c];
            w[i] = t;
            if (NORM) red[0] = t * t + 0.0;
        }
    }
    __device__ void finish_block(double* sums) const {
        if (p2p) { kb_p2p_allreduce_block<0>(*p2p, sums, 1); kb_arnoldi_fin(g, sums[0], ncols - 1); }
        else if (slots) { if (threadIdx.x == 0) slots[0] = sums[0]; }
        else kb_arnoldi_fin(g, sums[0], ncols - 1);
    }
};

// block orthogonalisation: v_{j+1} = (w - V h) / h_{j+1,j} in one pass (the epilogue ran before this sweep)
struct GsUpdateScaleOp : KbRedBase {
    static constexpr int NRED = 0;
    KbGmresDev g; const double* w; const double* hsrc; double* dst; int ncols;
    __device__ bool skip() const { return g.ctl->done != 0 || g.ctl->cycle_break != 0; }
    __device__ void pair(long long i, bool has1, double*) const {
        const double d = g.ctl->hnorm;
        if (has1) {
            double2 t = kb_ld2(w + i);
#pragma unroll 4
            for (int c = 0; c < ncols; ++c) {
                const double2 v = kb_ld2(g.V + (size_t)c * g.ld + i);
                const double h = hsrc[c];
                t.x = t.x - v.x * h; t.y = t.y - v.y * h;
            }
            kb_st2(dst + i, make_double2(t.x / d, t.y / d));
        } else {
            double t = w[i];
            for (int c = 0; c < ncols; ++c) t = t - g.V[(size_t)c * g.ld + i] * hsrc[c];
            dst[i] = t / d;
        }
    }
    __device__ void finish_block(double*) const {}
};

// dst = src / *div  (v_{j+1} = w / h_{j+1,j}: gmres.rs:102-103 ; v_0 = r / beta)
struct ScaleOp : KbRedBase {
    static constexpr int NRED = 0;
    KbGmresDev g; const double* src; double* dst; int start_vec;   // start_vec: divide by beta_g (column 0), else by hnorm (column j+1)
    __device__ bool skip() const {
        const KbCtl* c = g.ctl;
        if (c->done) return true;
        if (g.flex) return start_vec ? false : (c->cycle_break != 0);
        return start_vec ? false : (c->happy != 0 || c->cycle_break != 0);
    }
    __device__ void pair(long long i, bool has1, double*) const {
        const KbCtl* c = g.ctl;
        if (g.flex && !start_vec && c->hflag) {           // fgmres.rs:261: v_{j+1} := 0
            if (has1) kb_st2(dst + i, make_double2(0.0, 0.0)); else dst[i] = 0.0;
            return;
        }
        const double d = start_vec ? c->beta_g : c->hnorm;
        if (has1) { double2 t = kb_ld2(src + i); kb_st2(dst + i, make_double2(t.x / d, t.y / d)); }
        else dst[i] = src[i] / d;
    }
    __device__ void finish_block(double*) const {}
};

// back-substitution (gmres.rs:180-192), one block; y -> ctl->y
__global__ void kb_gmres_backsubst(KbCtl* c) {
    if (c->done) return;
    extern __shared__ double sh[];   // h: m*m (row-major copy), g: m, y: m
    const int m = c->m, ldh = c->restart + 1;
    double* H = sh; double* G = sh + (size_t)m * m; double* Y = G + m;
    for (int k = threadIdx.x; k < m * m; k += blockDim.x) { int i = k / m, j = k % m; H[k] = c->h[i + (size_t)ldh * j]; }
    for (int k = threadIdx.x; k < m; k += blockDim.x) G[k] = c->g[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = m - 1; i >= 0; --i) {
            double yi = G[i];
            for (int j = i + 1; j < m; ++j) yi = yi - H[i * m + j] * Y[j];
            if (c->side == KB_SIDE_FLEX) yi = yi / H[i * m + i];                      // fgmres.rs:309-315
            else if (fabs(H[i * m + i]) > KB_GM_EPS) yi = yi / H[i * m + i]; else yi = 0.0;
            Y[i] = yi; c->y[i] = yi;
        }
    }
}

// x += sum_j y_j V_j (gmres.rs:362-386); RIGHT: out = sum_j y_j V_j (then x += M^-1 out)
template <bool RIGHT>
struct UpdateXOp : KbRedBase {
    static constexpr int NRED = 0;
    KbGmresDev g; double* x; double* out; const double* basis;
    __device__ bool skip() const { return g.ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double*) const {
        const KbCtl* c = g.ctl;
        const int m = c->m;
        if (has1) {
            double2 t = RIGHT ? make_double2(0.0, 0.0) : kb_ld2(x + i);
#pragma unroll 4
            for (int j = 0; j < m; ++j) {
                const double2 v = kb_ld2(basis + (size_t)j * g.ld + i);
                const double y = c->y[j];
                t.x = t.x + y * v.x; t.y = t.y + y * v.y;
            }
            kb_st2((RIGHT ? out : x) + i, t);
        } else {
            double t = RIGHT ? 0.0 : x[i];
            for (int j = 0; j < m; ++j) t = t + c->y[j] * basis[(size_t)j * g.ld + i];
            (RIGHT ? out : x)[i] = t;
        }
    }
    __device__ void finish_block(double*) const {}
};
struct AddOp : KbRedBase {            // x += z
    static constexpr int NRED = 0;
    KbCtl* ctl; double* x; const double* z;
    __device__ bool skip() const { return ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double*) const {
        if (has1) { double2 a = kb_ld2(x + i), b = kb_ld2(z + i); kb_st2(x + i, make_double2(a.x + b.x, a.y + b.y)); }
        else x[i] = x[i] + z[i];
    }
    __device__ void finish_block(double*) const {}
};

// ---- scalar epilogues of the residual / norm kernels -------------------------------------------------------
__device__ __forceinline__ void gm_reset_cycle(KbCtl* c, double r0_norm) {
    c->beta_g = r0_norm;
    c->j = 0; c->m = 0; c->cycle_break = 0; c->happy = 0;
    c->g[0] = r0_norm;
    for (int k = 1; k <= c->restart; ++k) c->g[k] = 0.0;
}
struct GmInitFin {     // gmres.rs:221-233
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double beta = sqrt(s[0]);
        c->res0_true = beta; c->res0 = beta; c->res = beta; c->converged = 0; c->iter = 0; c->outer = 0;
        if (c->n_outer == 0) { c->done = 1; return; }
        gm_reset_cycle(c, beta);
    }
};
struct GmCycleFin {    // gmres.rs:388-398
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double beta = sqrt(s[0]);
        c->res = beta;
        c->converged = (beta < c->tol * c->res0_true) ? 1 : 0;
        c->outer = c->outer + 1;
        if (c->converged || c->iter >= c->max_iters || c->outer >= c->n_outer) { c->done = 1; return; }
        gm_reset_cycle(c, beta);
    }
};
struct FgInitFin {     // fgmres.rs:146-158
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double beta = sqrt(s[0]);
        c->res0_true = beta; c->res0 = beta; c->res = beta; c->converged = 0; c->iter = 0; c->outer = 0; c->early = 0; c->hflag = 0;
        if (beta == 0.0) { c->res0_true = 0.0; c->converged = 1; c->done = 1; return; }
        if (c->max_iters == 0) { c->done = 1; return; }
        gm_reset_cycle(c, beta);
    }
};
struct FgCycleFin {    // fgmres.rs:317-334
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double rn = sqrt(s[0]);
        if (rn < c->tol || c->early) { c->converged = 1; c->done = 1; return; }   // ABSOLUTE test, or the inner stop fired
        if (c->iter >= c->max_iters) { c->done = 1; return; }
        gm_reset_cycle(c, rn);
    }
};
struct GmLeftNormFin { // textbook left: r0_norm = ||M^-1 r||, inner denominator from the first cycle
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double r0n = sqrt(s[0]);
        c->beta_g = r0n; c->g[0] = r0n;
        if (c->outer == 0) c->res0 = r0n;
    }
};
template <class Fin>
struct NormOp : KbRedBase {
    static constexpr int NRED = 1;
    const double* z; KbCtl* ctl; KbFinish<Fin> fin;
    __device__ bool skip() const { return ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double* red) const {
        if (has1) { double2 t = kb_ld2(z + i); red[0] = t.x * t.x + t.y * t.y; }
        else red[0] = z[i] * z[i] + 0.0;
    }
    __device__ void finish_block(double* s) const { fin.template coop<0>(s); }
};

// ---- workspace --------------------------------------------------------------------------------------------
struct KbGmresWs {
    uint64_t n = 0, nx = 0; int restart = 0; size_t ld = 0;
    double *V = nullptr, *Z = nullptr, *x = nullptr, *b = nullptr, *r = nullptr, *w = nullptr, *t = nullptr, *z = nullptr;
    double* partials = nullptr; size_t pstride = 0;
    double* slots = nullptr;
    KbCtl* ctl = nullptr; KbCtl* h_ctl = nullptr;
    KbGraphCache gc;
};
void kb_gmres_ws_free(KbGmresWs* w) {
    if (!w) return;
    w->gc.reset();
    KB_FREE(w->V); KB_FREE(w->Z); KB_FREE(w->x); KB_FREE(w->b); KB_FREE(w->r); KB_FREE(w->w); KB_FREE(w->t); KB_FREE(w->z);
    KB_FREE(w->partials); KB_FREE(w->slots); KB_FREE(w->ctl);
    if (w->h_ctl) cudaFreeHost(w->h_ctl);
    delete w;
}
static int gmres_ws_get(kb_csr_s* A, int restart, KbGmresWs** out) {
    KbGmresWs* w = A->gmres_ws;
    if (w && w->restart < restart) { kb_gmres_ws_free(w); w = nullptr; A->gmres_ws = nullptr; }
    if (!w) {
        w = new KbGmresWs; A->gmres_ws = w;
        w->n = A->n; w->nx = A->ncols_local; w->restart = restart;
        w->ld = ((size_t)w->nx + 2 + 31) & ~(size_t)31;
        KB_TRY(kb_alloc(&w->V, w->ld * (size_t)(restart + 1)));
        KB_TRY(kb_alloc(&w->x, w->nx + 2)); KB_TRY(kb_alloc(&w->b, w->n + 2)); KB_TRY(kb_alloc(&w->r, w->n + 2));
        KB_TRY(kb_alloc(&w->w, w->n + 2)); KB_TRY(kb_alloc(&w->t, w->nx + 2)); KB_TRY(kb_alloc(&w->z, w->n + 2));
        w->pstride = (size_t)A->ntiles + 1;
        KB_TRY(kb_alloc(&w->partials, (size_t)(restart + 2) * w->pstride));
        KB_TRY(kb_alloc(&w->slots, 3 * (KB_MAX_RESTART + 8)));
        KB_TRY(kb_alloc(&w->ctl, 1));
        KB_CUDA(cudaMallocHost((void**)&w->h_ctl, sizeof(KbCtl)));
    }
    *out = w;
    return KB_OK;
}

template <class Op>
static int gm_tile(kb_csr_s* A, Op& op, int cls) {
    kb_ctx_s* c = A->ctx;
    op.n = (long long)A->n; op.ticket = c->ticket;
    { KbLaunch L(c, cls); kb_tile_kernel<<<A->ntiles, KB_THREADS, 0, c->stream>>>(op); }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

struct GmPlan { kb_csr_s* A; kb_pc_s* pc; KbGmresWs* w; int side; bool dist; int restart; KbGmresDev g; };
struct GmNoFin { __device__ void operator()(const double*) const {} };

// start vector of a cycle: (Left: z = M^-1 r, ||z||) ; V_0 = (z | r) / r0_norm
static int gm_start_vector(GmPlan& P) {
    kb_csr_s* A = P.A; KbGmresWs* w = P.w; kb_ctx_s* c = A->ctx;
    const double* src = w->r;
    if (P.side == KB_SIDE_LEFT) {
        KB_TRY(kb_pc_apply_dev(P.pc, w->r, w->z, w->ctl, 0));
        NormOp<GmLeftNormFin> op; op.partials = w->partials; op.pstride = w->pstride; op.z = w->z; op.ctl = w->ctl;
        op.fin = kb_make_fin(c, GmLeftNormFin{w->ctl}, P.dist, w->slots, 1);
        KB_TRY(gm_tile(A, op, KB_K_SMALL));
        if (P.dist) KB_TRY((kb_finish_dist<GmLeftNormFin>(c, GmLeftNormFin{w->ctl}, w->ctl, w->slots, 1)));
        src = w->z;
    }
    ScaleOp op; op.partials = nullptr; op.pstride = 0; op.g = P.g; op.src = src; op.dst = w->V; op.start_vec = 1;
    return gm_tile(A, op, KB_K_SMALL);
}

// inner Arnoldi step j.  j is a host constant: a replayed cycle always starts at j = 0, and once the
// device sets cycle_break every later step of the replay is skipped, so step k of the graph has j == k.
static int gm_inner_iteration(GmPlan& P, int j) {
    kb_csr_s* A = P.A; KbGmresWs* w = P.w; kb_ctx_s* c = A->ctx;
    KbCtl* ctl = w->ctl;
    double* vj = w->V + (size_t)j * w->ld;
    const int ncols = j + 1;
    // default: the shared-memory-staged 3-sweep CGS2 (kb_gs_fused, cp.async ring): measured on B200 (C4g) one fused sweep
    // costs 0.58 ms against 0.34 + 0.37 ms for the separate dot and update sweeps it replaces (291 -> 309 it/s).
    // KB_GS_FUSE=0 selects the 4-sweep form.
    static const bool fuse_smem_env = !getenv("KB_GS_FUSE") || atoi(getenv("KB_GS_FUSE")) != 0;
    const bool fuse_smem = fuse_smem_env && !P.g.flex && kb_gsf_stages(ncols) >= 2;
    typedef KbSpmvEpi<GmNoFin, false, false> Epi;
    Epi epi; epi.ctl = ctl; epi.skip_mask = 2; epi.fin = kb_make_fin(c, GmNoFin{}, false, nullptr, 0);
    if (P.g.flex) {                          // z_j = M^-1 v_j kept ; w = A z_j ; one classical GS pass (fgmres.rs:205-251)
        double* zj = P.pc ? P.g.Z + (size_t)j * w->ld : vj;
        if (P.pc) KB_TRY(kb_pc_apply_dev(P.pc, vj, zj, ctl, 2));
        KB_TRY((kb_launch_spmv<Epi, false>(A, zj, w->w, nullptr, nullptr, nullptr, 0, epi, P.dist ? zj : nullptr)));
        { KbLaunch L(c, KB_K_GS_DOT); kb_gs_dot<<<A->ntiles, KB_THREADS, 0, c->stream>>>(P.g, w->w, (long long)A->n, w->partials, w->pstride, ncols); }
        double* dst = P.dist ? w->slots : &ctl->h1[0];
        { KbLaunch L(c, KB_K_SMALL); kb_gs_level2<<<ncols, KB_THREADS, 0, c->stream>>>(ctl, w->partials, w->pstride, A->ntiles, dst); }
        if (P.dist) KB_TRY(kb_allreduce_slots(c, w->slots, ncols));
        double* s3 = w->slots + 2 * (KB_MAX_RESTART + 8);
        GsUpdateOp<true> op; op.partials = w->partials; op.pstride = w->pstride; op.g = P.g; op.w = w->w; op.hsrc = P.g.h1src;
        op.p2p = P.dist ? kb_p2p_dev_ptr(c) : nullptr;
        op.slots = (P.dist && !op.p2p) ? s3 : nullptr; op.ncols = ncols;
        KB_TRY(gm_tile(A, op, KB_K_GS_UPDATE));
        if (P.dist && !op.p2p) {
            KB_TRY(kb_allreduce_slots(c, s3, 1));
            KbLaunch L(c, KB_K_SMALL);
            kb_arnoldi_fin_kernel<<<1, KB_THREADS, 0, c->stream>>>(P.g, s3, j);
            KB_CUDA(cudaGetLastError());
        }
        ScaleOp sc; sc.partials = nullptr; sc.pstride = 0; sc.g = P.g; sc.src = w->w; sc.dst = w->V + (size_t)(j + 1) * w->ld; sc.start_vec = 0;
        return gm_tile(A, sc, KB_K_SMALL);
    }
    if (P.side == KB_SIDE_LEFT) {            // w = M^-1 (A v_j)
        KB_TRY((kb_launch_spmv<Epi, false>(A, vj, w->z, nullptr, nullptr, nullptr, 0, epi, P.dist ? vj : nullptr)));
        KB_TRY(kb_pc_apply_dev(P.pc, w->z, w->w, ctl, 2));
    } else if (P.side == KB_SIDE_RIGHT) {    // w = A (M^-1 v_j)
        KB_TRY(kb_pc_apply_dev(P.pc, vj, w->t, ctl, 2));
        KB_TRY((kb_launch_spmv<Epi, false>(A, w->t, w->w, nullptr, nullptr, nullptr, 0, epi, P.dist ? w->t : nullptr)));
    } else {                                 // w = A v_j   (gmres.rs:80-81)
        KB_TRY((kb_launch_spmv<Epi, false>(A, vj, w->w, nullptr, nullptr, nullptr, 0, epi, P.dist ? vj : nullptr)));
    }
    if (P.g.block) {   // {h, w.w} in ONE sweep and ONE reduction, epilogue, then v_{j+1} = (w - V h) / h_{j+1,j}: 2 sweeps over V per step
        { KbLaunch L(c, KB_K_GS_DOT); kb_gs_dot<<<A->ntiles, KB_THREADS, 0, c->stream>>>(P.g, w->w, (long long)A->n, w->partials, w->pstride, ncols, 1); }
        double* dst = P.dist ? w->slots : &ctl->h1[0];
        { KbLaunch L(c, KB_K_SMALL); kb_gs_level2<<<ncols + 1, KB_THREADS, 0, c->stream>>>(ctl, w->partials, w->pstride, A->ntiles, dst); }
        if (P.dist) KB_TRY(kb_allreduce_slots(c, w->slots, ncols + 1));
        { KbLaunch L(c, KB_K_SMALL); kb_block_fin_kernel<<<1, KB_THREADS, 0, c->stream>>>(P.g, P.g.h1src, j); }
        KB_CUDA(cudaGetLastError());
        GsUpdateScaleOp op; op.partials = nullptr; op.pstride = 0; op.g = P.g; op.w = w->w; op.hsrc = P.g.h1src; op.dst = w->V + (size_t)(j + 1) * w->ld; op.ncols = ncols;
        return gm_tile(A, op, KB_K_GS_UPDATE);
    }
    {   // h1 = V^T w ; w -= V h1
        { KbLaunch L(c, KB_K_GS_DOT); kb_gs_dot<<<A->ntiles, KB_THREADS, 0, c->stream>>>(P.g, w->w, (long long)A->n, w->partials, w->pstride, ncols); }
        double* dst = P.dist ? w->slots : &ctl->h1[0];
        { KbLaunch L(c, KB_K_SMALL); kb_gs_level2<<<ncols, KB_THREADS, 0, c->stream>>>(ctl, w->partials, w->pstride, A->ntiles, dst); }
        if (P.dist) KB_TRY(kb_allreduce_slots(c, w->slots, ncols));
        if (fuse_smem) {   // one pass over V: w -= V h1 and the tile sums of h2 = V^T w (shared-memory ring, cp.async)
            const int nst = kb_gsf_stages(ncols);
            kb_gsf_fn kfn = kb_gsf_kernel(nst);
            if (!c->configured.count((const void*)kfn)) {
                KB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
                c->configured.insert((const void*)kfn);
            }
            KbLaunch L(c, KB_K_GS_UPDATE);
            const size_t sh = kb_gsf_smem(ncols, nst);
            const int per_sm = std::max(1, std::min(6, (int)((227 * 1024) / (sh + 1024))));
            kfn<<<std::min(per_sm * c->sm_count, A->ntiles), KB_GSF_THREADS, sh, c->stream>>>(P.g, w->w, P.g.h1src, (int)A->n, A->ntiles, w->partials, w->pstride, ncols);
            KB_CUDA(cudaGetLastError());
        } else {
            GsUpdateOp<false> op; op.partials = nullptr; op.pstride = 0; op.g = P.g; op.w = w->w; op.hsrc = P.g.h1src; op.slots = nullptr; op.ncols = ncols; op.p2p = nullptr;
            KB_TRY(gm_tile(A, op, KB_K_GS_UPDATE));
        }
    }
    {   // h2 = V^T w ; w -= V h2 fused with ||w||^2 ; Arnoldi epilogue (H column, Givens, stop test)
        if (!fuse_smem) { KbLaunch L(c, KB_K_GS_DOT); kb_gs_dot<<<A->ntiles, KB_THREADS, 0, c->stream>>>(P.g, w->w, (long long)A->n, w->partials, w->pstride, ncols); }
        double* s2 = w->slots + (KB_MAX_RESTART + 8);
        double* dst = P.dist ? s2 : &ctl->h2[0];
        { KbLaunch L(c, KB_K_SMALL); kb_gs_level2<<<ncols, KB_THREADS, 0, c->stream>>>(ctl, w->partials, w->pstride, A->ntiles, dst); }
        if (P.dist) KB_TRY(kb_allreduce_slots(c, s2, ncols));
        double* s3 = w->slots + 2 * (KB_MAX_RESTART + 8);
        GsUpdateOp<true> op; op.partials = w->partials; op.pstride = w->pstride; op.g = P.g; op.w = w->w; op.hsrc = P.g.h2src;
        op.p2p = P.dist ? kb_p2p_dev_ptr(c) : nullptr;
        op.slots = (P.dist && !op.p2p) ? s3 : nullptr; op.ncols = ncols;
        KB_TRY(gm_tile(A, op, KB_K_GS_UPDATE));
        if (P.dist && !op.p2p) {
            KB_TRY(kb_allreduce_slots(c, s3, 1));
            KbLaunch L(c, KB_K_SMALL);
            kb_arnoldi_fin_kernel<<<1, KB_THREADS, 0, c->stream>>>(P.g, s3, j);
            KB_CUDA(cudaGetLastError());
        }
    }
    {   // v_{j+1} = w / h_{j+1,j}   (skipped on happy breakdown / stop)
        ScaleOp op; op.partials = nullptr; op.pstride = 0; op.g = P.g; op.src = w->w; op.dst = w->V + (size_t)(j + 1) * w->ld; op.start_vec = 0;
        KB_TRY(gm_tile(A, op, KB_K_SMALL));
    }
    return KB_OK;
}

// one restart cycle = `restart` inner steps, least-squares solve, x update, true residual + cycle test,
// and the next cycle's start vector
static int gm_cycle(GmPlan& P) {
    kb_csr_s* A = P.A; KbGmresWs* w = P.w; kb_ctx_s* c = A->ctx;
    for (int j = 0; j < P.restart; ++j) KB_TRY(gm_inner_iteration(P, j));
    {
        KbLaunch L(c, KB_K_SMALL);
        const size_t sh = ((size_t)P.restart * P.restart + 2 * (size_t)P.restart) * sizeof(double);
        if (sh > 48 * 1024 && !c->configured.count((const void*)kb_gmres_backsubst)) {
            KB_CUDA(cudaFuncSetAttribute(kb_gmres_backsubst, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((KB_MAX_RESTART * KB_MAX_RESTART + 2 * KB_MAX_RESTART) * sizeof(double))));
            c->configured.insert((const void*)kb_gmres_backsubst);
        }
        kb_gmres_backsubst<<<1, KB_THREADS, sh, c->stream>>>(w->ctl);
        KB_CUDA(cudaGetLastError());
    }
    if (P.side == KB_SIDE_RIGHT) {           // x += M^-1 (V y)
        UpdateXOp<true> op; op.partials = nullptr; op.pstride = 0; op.g = P.g; op.x = w->x; op.out = w->w; op.basis = w->V;
        KB_TRY(gm_tile(A, op, KB_K_GS_UPDATE));
        KB_TRY(kb_pc_apply_dev(P.pc, w->w, w->z, w->ctl, 0));
        AddOp add; add.partials = nullptr; add.pstride = 0; add.ctl = w->ctl; add.x = w->x; add.z = w->z;
        KB_TRY(gm_tile(A, add, KB_K_SMALL));
    } else {
        UpdateXOp<false> op; op.partials = nullptr; op.pstride = 0; op.g = P.g; op.x = w->x; op.out = nullptr; op.basis = P.g.flex ? P.g.Z : w->V;
        KB_TRY(gm_tile(A, op, KB_K_GS_UPDATE));
    }
    {   // r = b - A x ; beta = ||r|| ; converged = beta < tol * res0 (gmres.rs:388-398)
        if (P.g.flex) {
            typedef KbSpmvEpi<FgCycleFin, false, true> Epi;
            Epi epi; epi.ctl = w->ctl; epi.fin = kb_make_fin(c, FgCycleFin{w->ctl}, P.dist, w->slots, 1);
            KB_TRY((kb_launch_spmv<Epi, true>(A, w->x, w->r, w->b, nullptr, w->partials, w->pstride, epi, P.dist ? w->x : nullptr)));
            if (P.dist) KB_TRY((kb_finish_dist<FgCycleFin>(c, FgCycleFin{w->ctl}, w->ctl, w->slots, 1)));
        } else {
        typedef KbSpmvEpi<GmCycleFin, false, true> Epi;
        Epi epi; epi.ctl = w->ctl; epi.fin = kb_make_fin(c, GmCycleFin{w->ctl}, P.dist, w->slots, 1);
        KB_TRY((kb_launch_spmv<Epi, true>(A, w->x, w->r, w->b, nullptr, w->partials, w->pstride, epi, P.dist ? w->x : nullptr)));
        if (P.dist) KB_TRY((kb_finish_dist<GmCycleFin>(c, GmCycleFin{w->ctl}, w->ctl, w->slots, 1)));
        }
    }
    return gm_start_vector(P);
}

static int gmres_solve_impl(kb_csr A, kb_pc pc, const double* b, double* x, uint64_t restart, double tol, uint64_t max_iters, int side,
                            uint32_t flags, kb_stats* stats);
extern "C" int kb_gmres_solve(kb_csr A, kb_pc pc, const double* b, double* x, uint64_t restart, double tol, uint64_t max_iters, int side,
                              uint32_t flags, kb_stats* stats) {
    if (side < 0 || side > 2) { kb_set_error("bad preconditioning side"); return KB_SOLVE_ERROR; }
    return gmres_solve_impl(A, pc, b, x, restart, tol, max_iters, side, flags, stats);
}
// FgmresSolver::new(tol, max_iters, restart).solve_flex(&a, pc, &b, &mut x)  (src/solver/fgmres.rs:52-340)
extern "C" int kb_fgmres_solve(kb_csr A, kb_pc pc, const double* b, double* x, uint64_t restart, double tol, uint64_t max_iters,
                               uint32_t flags, kb_stats* stats) {
    return gmres_solve_impl(A, pc, b, x, restart, tol, max_iters, KB_SIDE_FLEX, flags, stats);
}
static int gmres_solve_impl(kb_csr A, kb_pc pc, const double* b, double* x, uint64_t restart, double tol, uint64_t max_iters, int side,
                            uint32_t flags, kb_stats* stats) {
    if (!A || !b || !x || !stats) { kb_set_error("kb_gmres_solve: null argument"); return KB_SOLVE_ERROR; }
    if (pc && pc->a != A) { kb_set_error("preconditioner was set up for a different operator"); return KB_SOLVE_ERROR; }
    if (restart < 1 || restart > KB_MAX_RESTART) { kb_set_error("restart must be in [1,%d]", KB_MAX_RESTART); return KB_UNSUPPORTED; }
    const bool flex = side == KB_SIDE_FLEX;
    if (!pc && !flex) side = KB_SIDE_NONE;      // gmres.rs:262: `_ =>` branch when pc is None
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    const bool dev = (flags & KB_FLAG_DEVICE_PTRS) != 0;
    const bool dist = A->dist && c->size > 1;
    KbGmresWs* w = nullptr;
    KB_TRY(gmres_ws_get(A, (int)restart, &w));
    memset(stats, 0, sizeof(*stats));
    if (A->n == 0 && !dist) return KB_OK;
    KB_TRY(kb_upload_or_alias(c, b, w->b, w->n, dev));
    KB_TRY(kb_upload_or_alias(c, x, w->x, w->n, dev));
    KbCtl* h = w->h_ctl;
    memset(h, 0, offsetof(KbCtl, h));
    h->max_iters = max_iters; h->tol = tol; h->restart = (int)restart; h->side = side;
    h->n_outer = (int)std::min<uint64_t>((max_iters + restart - 1) / restart, 0x7fffffffull);
    if (flex) h->n_outer = (int)std::min<uint64_t>(max_iters, 0x7fffffffull);   // `while total_iters < max_iters` (fgmres.rs:162): every cycle makes >= 1 step
    if (flex && pc && !w->Z) KB_TRY(kb_alloc(&w->Z, w->ld * (size_t)w->restart));
    KB_TRY(kb_hist_prepare(A, flags, max_iters, h));
    KbMonitor mon;
    if ((flags & KB_FLAG_MONITOR) && A->monitor) { mon.fn = A->monitor; mon.user = A->monitor_user; mon.d_hist = h->hist; mon.cap = h->hist_cap; mon.index_offset = 1; }
    KB_CUDA(cudaMemcpyAsync(w->ctl, h, offsetof(KbCtl, h), cudaMemcpyHostToDevice, c->stream));
    GmPlan P{A, pc, w, side, dist, (int)restart, {}};
    P.g.ctl = w->ctl; P.g.V = w->V; P.g.ld = w->ld;
    P.g.h1src = dist ? w->slots : &w->ctl->h1[0];
    P.g.h2src = dist ? w->slots + (KB_MAX_RESTART + 8) : &w->ctl->h2[0];
    P.g.flex = flex ? 1 : 0; P.g.Z = (flex && pc) ? w->Z : w->V;
    const bool block = (flags & KB_FLAG_BLOCK_ORTH) != 0;
    if (block && flex) { kb_set_error("KB_FLAG_BLOCK_ORTH applies to kb_gmres_solve (FGMRES already makes a single pass)"); return KB_UNSUPPORTED; }
    P.g.block = block ? 1 : 0;
    const bool profile = (flags & KB_FLAG_PROFILE) != 0;
    const bool use_graph = !(flags & (KB_FLAG_NO_GRAPH | KB_FLAG_PROFILE));
    const bool was_prof = c->profiling;
    c->profiling = profile;
    int st = KB_OK;
    do {
        {   // r0 = b - A x ; beta = ||r0|| (gmres.rs:221-229)
            if (flex) {
                typedef KbSpmvEpi<FgInitFin, false, true> Epi;
                Epi epi; epi.ctl = nullptr; epi.fin = kb_make_fin(c, FgInitFin{w->ctl}, dist, w->slots, 1);
                if ((st = kb_launch_spmv<Epi, true>(A, w->x, w->r, w->b, nullptr, w->partials, w->pstride, epi, dist ? w->x : nullptr)) != KB_OK) break;
                if (dist && (st = kb_finish_dist<FgInitFin>(c, FgInitFin{w->ctl}, w->ctl, w->slots, 1)) != KB_OK) break;
            } else {
            typedef KbSpmvEpi<GmInitFin, false, true> Epi;
            Epi epi; epi.ctl = nullptr; epi.fin = kb_make_fin(c, GmInitFin{w->ctl}, dist, w->slots, 1);
            if ((st = kb_launch_spmv<Epi, true>(A, w->x, w->r, w->b, nullptr, w->partials, w->pstride, epi, dist ? w->x : nullptr)) != KB_OK) break;
            if (dist && (st = kb_finish_dist<GmInitFin>(c, GmInitFin{w->ctl}, w->ctl, w->slots, 1)) != KB_OK) break;
            }
        }
        if ((st = gm_start_vector(P)) != KB_OK) break;
        const uint64_t key = (((kb_pc_serial(pc) + 1) * 4 + (uint64_t)side) * 256 + restart) * 2 + (block ? 1 : 0);
        st = kb_run_iterations(c, &w->gc, key, 1, (uint64_t)h->n_outer, use_graph, w->ctl, h, [&]() { return gm_cycle(P); }, &mon);
        if (st != KB_OK) break;
        if (cudaMemcpyAsync(h, w->ctl, offsetof(KbCtl, h), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("gmres: readback failed"); st = KB_SOLVE_ERROR; break; }
        stats->iterations = h->iter; stats->final_residual = flex ? h->res0_true : h->res; stats->converged = h->converged; stats->breakdown = h->happy;
        A->hist_len = std::min<uint64_t>(h->hist_len, h->hist_cap);
        st = h->status;
        if (dist && kb_p2p_error(c)) { kb_set_error("%s: peer-memory collective timed out", "gmres"); st = KB_SOLVE_ERROR; break; }
        if (pc && kb_ilu0_error(const_cast<kb_pc_s*>(pc))) { kb_set_error("%s: a triangular-solve dependency wait timed out", "gmres"); st = KB_SOLVE_ERROR; break; }
        if (st == KB_OK) {
            if (cudaMemcpyAsync(x, w->x, w->n * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("gmres: copy-out of x failed"); st = KB_SOLVE_ERROR; }
        }
    } while (0);
    c->profiling = was_prof;
    if (profile) kb_prof_collect(c);
    return st;
}
