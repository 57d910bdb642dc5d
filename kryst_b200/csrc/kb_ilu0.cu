// kb_ilu0.cu — textbook ILU(0) on the CSR pattern, all on the device (K9-K11 of SURVEY §7.2).
//
// Replaces src/preconditioner/ilu.rs:59-122, which is dense, O(n^3) and numerically not an ILU
// (SURVEY F5).  Specification = Saad Alg. 10.4 (IKJ) on the sorted pattern, unit-lower L, U with stored
// inverse diagonal; on a row-block shard couplings to ghost columns are dropped (block-Jacobi ILU(0) ==
// AdditiveSchwarz with overlap 0 over the chunk partition, asm.rs:34-57).
//   symbolic : diag_ptr; level sets lev_L[i] = 1 + max lev_L[k<i in pattern] (lev_U mirrored) by monotone
//              relaxation sweeps on the device; rows bucketed per level with a stable radix sort so each
//              level lists its rows ascending -> integer structures bit-exact vs the oracle.
//   numeric  : one launch per level, a thread per row runs the row's IKJ updates in the oracle's order
//              (k ascending, j ascending; mul then sub) -> factors bit-exact vs the oracle.
//   apply    : two sync-free triangular solves (forward unit-L, backward U) without grid-wide barriers.
//              Rows are laid out in level order (levels padded to warps) and cut into 256-slot chunks; the
//              factors are copied once into a chunk-local slot-major (ELL) layout so loads coalesce.  A
//              persistent, fully co-resident grid walks the chunks in order: a CTA prefetches its chunk's
//              static data, one thread gates it on a per-level completion counter until the wavefront is
//              three levels behind (so only a few levels of rows ever poll L2), then every thread polls its
//              dependencies in parallel.  "Not ready" is the all-ones NaN pattern stored in the solution
//              vector itself (value == flag: one 64-bit load per dependency, no fences on the critical path).
//              Measured on B200, 256^3 7-pt: 11.1 ms -> 3.0 ms per apply vs the first ticketed version.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cub/cub.cuh>
#include "kb_objects.h"

#define KB_SENTINEL 0xFFFFFFFFFFFFFFFFull

__global__ void k_ilu_count_local(const int* __restrict__ rp, const int* __restrict__ col, int n, int* __restrict__ cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = 0;
    for (int p = rp[i]; p < rp[i + 1]; ++p) c += (col[p] < n);
    cnt[i] = c;
}
__global__ void k_ilu_fill_local(const int* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ vals, int n,
                                 const int* __restrict__ lrp, int* __restrict__ lcol, double* __restrict__ lu) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int q = lrp[i];
    for (int p = rp[i]; p < rp[i + 1]; ++p)
        if (col[p] < n) { lcol[q] = col[p]; lu[q] = vals[p]; ++q; }
}
__global__ void k_ilu_diag_ptr(const int* __restrict__ lrp, const int* __restrict__ lcol, int n, int* __restrict__ dp,
                               unsigned long long* bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = -1;
    for (int p = lrp[i]; p < lrp[i + 1]; ++p)
        if (lcol[p] == i) { d = p; break; }
    dp[i] = d < 0 ? lrp[i + 1] : d;
    if (d < 0) atomicMin(bad, (unsigned long long)i);
}
// one relaxation sweep of lev[i] = 1 + max(lev[k]) over the strictly lower (upper) pattern
__global__ void k_ilu_level_sweep(const int* __restrict__ lrp, const int* __restrict__ lcol, const int* __restrict__ dp, int n,
                                  int upper, int* lev, int* changed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lv = 0;
    const int a = upper ? dp[i] + 1 : lrp[i], b = upper ? lrp[i + 1] : dp[i];
    for (int p = a; p < b; ++p) {
        int k = lcol[p];
        int t = ((volatile int*)lev)[k] + 1;
        lv = t > lv ? t : lv;
    }
    if (lv > lev[i]) { lev[i] = lv; *changed = 1; }
}
__global__ void k_iota(int* a, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
__global__ void k_level_ptr(const int* __restrict__ sorted_lev, int n, int nlev, int* __restrict__ level_ptr) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > nlev) return;
    int lo = 0, hi = n;   // first position with sorted_lev >= l
    while (lo < hi) { int mid = (lo + hi) >> 1; if (sorted_lev[mid] < l) lo = mid + 1; else hi = mid; }
    level_ptr[l] = lo;
}
// warp-padded schedule: level l occupies [spad[l], spad[l+1]) with -1 fill
__global__ void k_sched_fill(const int* __restrict__ order, const int* __restrict__ level_ptr, const int* __restrict__ spad,
                             int nlev, int* __restrict__ sched) {
    int l = blockIdx.x;
    if (l >= nlev) return;
    const int a = level_ptr[l], cnt = level_ptr[l + 1] - a, o = spad[l], e = spad[l + 1];
    for (int t = threadIdx.x; t < e - o; t += blockDim.x) sched[o + t] = t < cnt ? order[a + t] : -1;
}

// numeric factorisation of the rows of one level (thread per row) — Saad Alg. 10.4, IKJ
__global__ void k_ilu_factor_level(const int* __restrict__ rows, int nrows, const int* __restrict__ lrp, const int* __restrict__ lcol,
                                   const int* __restrict__ dp, double* lu, double* __restrict__ inv_ud, unsigned long long* bad) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrows) return;
    const int i = rows[t];
    const int rb = lrp[i], re = lrp[i + 1], di = dp[i];
    for (int p = rb; p < di; ++p) {
        const int k = lcol[p];
        const double lik = lu[p] / lu[dp[k]];
        lu[p] = lik;
        int pos = p + 1;
        const int ke = lrp[k + 1];
        for (int q = dp[k] + 1; q < ke; ++q) {
            const int j = lcol[q];
            while (pos < re && lcol[pos] < j) ++pos;
            if (pos < re && lcol[pos] == j) lu[pos] = lu[pos] - lik * lu[q];
        }
    }
    const double piv = lu[di];
    if (piv == 0.0 || piv != piv) atomicMin(bad, (unsigned long long)i);
    inv_ud[i] = 1.0 / piv;
}

// ---- sync-free triangular solves ------------------------------------------------------------------------
struct KbTrsvArgs {
    const int* __restrict__ sched; int sched_len;
    const int* __restrict__ lrp; const int* __restrict__ lcol; const int* __restrict__ dp;
    const double* __restrict__ lu; const double* __restrict__ inv_ud;
    const double* __restrict__ rhs;   // forward: r ; backward: y
    double* out;                      // forward: y ; backward: z   (pre-filled with the NaN sentinel)
    unsigned* counters;               // [0] block ticket, [1] finished blocks, [2] error flag
    const KbCtl* skip_ctl; int skip_mask;
    double* fill;                     // forward solve also marks the backward solve's output "not ready"
};

__device__ __forceinline__ double kb_wait_value(const double* p, unsigned* err) {
    const volatile unsigned long long* q = reinterpret_cast<const volatile unsigned long long*>(p);
    unsigned long long v = *q;
    unsigned spins = 0;
    while (v == KB_SENTINEL) {
        v = *q;
        if (++spins > (1u << 26)) { atomicExch(err, 1u); break; }   // never hang the GPU on a broken dependency graph
    }
    return __longlong_as_double((long long)v);
}

// mark both solution vectors "not ready" (replaces two memsets; skippable inside solver graphs)
__global__ void kb_trsv_fill(double* a, double* b, long long n, const KbCtl* skip_ctl, int skip_mask, int* done0, int n0, int* done1, int n1) {
    if (kb_skip(skip_ctl, skip_mask)) return;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (done0 && i < n0) done0[i] = 0;
    if (done1 && i < n1) done1[i] = 0;
    if (i < n) {
        reinterpret_cast<unsigned long long*>(a)[i] = KB_SENTINEL;
        reinterpret_cast<unsigned long long*>(b)[i] = KB_SENTINEL;
    }
}

template <bool UPPER>
__global__ void __launch_bounds__(KB_THREADS) kb_trsv_syncfree(KbTrsvArgs a) {
    if (kb_skip(a.skip_ctl, a.skip_mask)) return;
    __shared__ unsigned s_blk;
    if (threadIdx.x == 0) s_blk = atomicAdd(&a.counters[0], 1u);
    __syncthreads();
    const int t = (int)s_blk * KB_THREADS + threadIdx.x;
    if (t < a.sched_len) {
        const int i = a.sched[t];
        if (i >= 0) {
            double s = a.rhs[i];
            if (!UPPER) {
                const int pe = a.dp[i];
                for (int p = a.lrp[i]; p < pe; ++p) s = s - a.lu[p] * kb_wait_value(a.out + a.lcol[p], &a.counters[2]);
            } else {
                const int pe = a.lrp[i + 1];
                for (int p = a.dp[i] + 1; p < pe; ++p) s = s - a.lu[p] * kb_wait_value(a.out + a.lcol[p], &a.counters[2]);
                s = s * a.inv_ud[i];
            }
            // publish: a single 64-bit store is the data and the ready flag at once
            unsigned long long sb = (unsigned long long)__double_as_longlong(s);
            if (sb == KB_SENTINEL) sb = 0x7FF8000000000000ull;      // never publish the "not ready" pattern as a value
            *reinterpret_cast<volatile unsigned long long*>(a.out + i) = sb;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned f = atomicAdd(&a.counters[1], 1u);
        if (f == gridDim.x - 1u) { a.counters[0] = 0u; a.counters[1] = 0u; __threadfence(); }
    }
}

// ---- level-ordered chunk-ELL copy of the factors + persistent sync-free solve -----------------------------
// The schedule (rows in level order, levels padded to warps, total padded to 256) is cut into chunks of 256
// slots.  Chunk c stores its rows' strictly-lower (upper) entries slot-major: entry e of slot t lives at
// off[c] + e*256 + t, so a warp's loads of values / column ids are fully coalesced.  Entry order inside a
// row is unchanged (ascending column), hence the solve performs the oracle's exact operation sequence.
__global__ void k_ell_width(const int* __restrict__ sched, int nchunks, const int* __restrict__ lrp, const int* __restrict__ dp,
                            int upper, int* __restrict__ width) {
    __shared__ int smax;
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
    const int row = sched[blockIdx.x * KB_THREADS + threadIdx.x];
    int len = 0;
    if (row >= 0) len = upper ? (lrp[row + 1] - dp[row] - 1) : (dp[row] - lrp[row]);
    atomicMax(&smax, len);
    __syncthreads();
    if (threadIdx.x == 0) width[blockIdx.x] = smax;
}
__global__ void k_ell_fill(const int* __restrict__ sched, const int* __restrict__ lrp, const int* __restrict__ lcol, const int* __restrict__ dp,
                           const double* __restrict__ lu, const double* __restrict__ inv_ud, int upper, const int* __restrict__ width,
                           const long long* __restrict__ off, int* __restrict__ ecol, double* __restrict__ eval, double* __restrict__ ediag) {
    const int c = blockIdx.x, t = threadIdx.x;
    const int row = sched[c * KB_THREADS + t];
    const int w = width[c];
    const long long o = off[c];
    int a = 0, b = 0;
    if (row >= 0) { a = upper ? dp[row] + 1 : lrp[row]; b = upper ? lrp[row + 1] : dp[row]; }
    for (int e = 0; e < w; ++e) {
        const bool has = a + e < b;
        ecol[o + (long long)e * KB_THREADS + t] = has ? lcol[a + e] : -1;
        eval[o + (long long)e * KB_THREADS + t] = has ? lu[a + e] : 0.0;
    }
    if (ediag) ediag[c * KB_THREADS + t] = row >= 0 ? inv_ud[row] : 0.0;
}

struct KbTrsvEll {
    const int* __restrict__ sched; int nchunks;
    const int* __restrict__ width; const long long* __restrict__ off;
    const int* __restrict__ ecol; const double* __restrict__ eval; const double* __restrict__ ediag;
    const double* __restrict__ rhs; double* out;
    unsigned* counters;
    const KbCtl* skip_ctl; int skip_mask;
    int sleep_ns;
    const int* __restrict__ spad;       // [nlev+1] padded schedule offset of every level
    const int* __restrict__ chunk_lev;  // [nchunks] level of the chunk's first slot
    int* done;                          // [nlev] completed slots per level (zeroed by the fill kernel)
    int nlev; int gate;                 // gate: wait until level (first - gate) is complete before value-polling
};

__device__ __forceinline__ double kb_wait_value_bo(const double* p, unsigned* err, int sleep_ns) {
    const volatile unsigned long long* q = reinterpret_cast<const volatile unsigned long long*>(p);
    unsigned long long v = *q;
    unsigned spins = 0;
    while (v == KB_SENTINEL) {
        if (sleep_ns > 0) __nanosleep(sleep_ns);
        v = *q;
        if (++spins > (1u << 24)) { atomicExch(err, 1u); break; }
    }
    return __longlong_as_double((long long)v);
}

__device__ __forceinline__ unsigned long long kb_ld_relaxed_gpu(const double* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void kb_st_relaxed_gpu(double* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Persistent grid: CTA b handles chunks b, b+G, b+2G, ... in order.  All G CTAs are co-resident (G <= resident
// capacity), and a chunk only waits on rows of earlier chunks, so progress is guaranteed without tickets; G also
// bounds the window of in-flight rows to a few levels, so few threads spin at any time.
template <bool UPPER>
__global__ void __launch_bounds__(KB_THREADS) kb_trsv_persistent(KbTrsvEll a) {
    if (kb_skip(a.skip_ctl, a.skip_mask)) return;
    const int tid = threadIdx.x;
    for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
        // ---- prefetch everything that does not depend on other rows
        const int row = a.sched[c * KB_THREADS + tid];
        const int w = a.width[c];
        const long long o = a.off[c] + tid;
        const int lev0 = a.chunk_lev[c];
        double s = 0.0, dg = 1.0;
        int cc[4]; double vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { cc[u] = -1; vv[u] = 0.0; }
        if (row >= 0) {
            s = a.rhs[row];
            if (UPPER) dg = a.ediag[c * KB_THREADS + tid];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (u < w) { cc[u] = __ldcs(a.ecol + o + (long long)u * KB_THREADS); vv[u] = __ldcs(a.eval + o + (long long)u * KB_THREADS); }
        }
        // ---- gate: one thread waits until level lev0-gate is complete, so that only ~gate levels of rows
        //      value-poll at any time (keeps L2 free of a poll storm)
        const int gl = lev0 - a.gate;
        if (gl >= 0) {
            if (tid == 0) {
                const int need = a.spad[gl + 1] - a.spad[gl];
                const volatile int* dq = a.done + gl;
                unsigned spins = 0;
                while (*dq < need) {
                    __nanosleep(100);
                    if (++spins > (1u << 24)) { atomicExch(&a.counters[2], 1u); break; }
                }
            }
            __syncthreads();
        }
        if (row >= 0) {
            for (int e0 = 0; e0 < w; e0 += 4) {
                if (e0 > 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        cc[u] = -1; vv[u] = 0.0;
                        if (e0 + u < w) { cc[u] = __ldcs(a.ecol + o + (long long)(e0 + u) * KB_THREADS); vv[u] = __ldcs(a.eval + o + (long long)(e0 + u) * KB_THREADS); }
                    }
                }
                // poll the (up to 4) dependencies of this group in parallel
                unsigned long long dv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) dv[u] = cc[u] >= 0 ? kb_ld_relaxed_gpu(a.out + cc[u]) : 0ull;
                unsigned spins = 0;
                while (true) {
                    bool pending = false;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (cc[u] >= 0 && dv[u] == KB_SENTINEL) {
                            dv[u] = kb_ld_relaxed_gpu(a.out + cc[u]);
                            pending = pending || (dv[u] == KB_SENTINEL);
                        }
                    if (!pending) break;
                    if (a.sleep_ns > 0) __nanosleep(a.sleep_ns);
                    if (++spins > (1u << 24)) { atomicExch(&a.counters[2], 1u); break; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (cc[u] >= 0) s = s - vv[u] * __longlong_as_double((long long)dv[u]);
            }
            if (UPPER) s = s * dg;
            unsigned long long sb = (unsigned long long)__double_as_longlong(s);
            if (sb == KB_SENTINEL) sb = 0x7FF8000000000000ull;      // never publish the "not ready" pattern as a value
            kb_st_relaxed_gpu(a.out + row, sb);
        }
        // ---- publish per-level completion (slots of every level present in this chunk)
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const int cb = c * KB_THREADS, ce = cb + KB_THREADS;
            for (int l = lev0; l < a.nlev; ++l) {
                const int lb = a.spad[l], le = a.spad[l + 1];
                if (lb >= ce) break;
                const int cnt = min(le, ce) - max(lb, cb);
                if (cnt > 0) atomicAdd(a.done + l, cnt);
            }
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------------
struct KbTileSolve;                                  // kb_trsv_tiles.cu: block-wavefront solves for grid-structured factors
int kb_tiles_build(kb_pc_s* pc, unsigned* d_err, KbTileSolve** out);
int kb_tiles_apply(kb_pc_s* pc, KbTileSolve* t, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask);
void kb_tiles_free(KbTileSolve* t);
void kb_tiles_grid(const KbTileSolve* t, int* nx, int* ny, int* nz);
struct KbLean;                                       // kb_trsv_lean.cu: warp-specialised pencil march (full-stencil box grids; the default)
int kb_lean_build(kb_pc_s* pc, int gx, int gy, int gz, unsigned* d_err, KbLean** out);
int kb_lean_apply(kb_pc_s* pc, KbLean* m, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask);
void kb_lean_free(KbLean* m);
int kb_lean_trace_get(KbLean* m, unsigned long long* out, int* px, int* py);
struct KbIluExtra {
    int* level[2] = {nullptr, nullptr};
    unsigned* counters = nullptr;
    bool owns_pattern = false;
    // level-ordered chunk-ELL copies (index 0: L, 1: U)
    int nchunks[2] = {0, 0};
    int* width[2] = {nullptr, nullptr};
    long long* off[2] = {nullptr, nullptr};
    int* ecol[2] = {nullptr, nullptr};
    double* eval[2] = {nullptr, nullptr};
    double* ediag = nullptr;
    int grid[2] = {0, 0};
    int* spad[2] = {nullptr, nullptr};
    int* chunk_lev[2] = {nullptr, nullptr};
    int* done[2] = {nullptr, nullptr};
    int gate = 3;
    int kind = 1;            // 1 persistent chunk-ELL solve, 0 ticketed CSR solve
    KbTileSolve* tiles = nullptr;   // non-null: the factor's pattern is a 5-/7-point box grid -> block-wavefront solves
    KbLean* lean = nullptr;         // non-null: ... and a full stencil -> warp-specialised pencil march (preferred)
    int sleep_ns = 0;
};
static KbIluExtra* extra_of(kb_pc_s* pc) { return reinterpret_cast<KbIluExtra*>(pc->extra); }

void kb_ilu0_free(kb_pc_s* pc) {
    if (!pc || pc->kind != KB_PC_ILU0) return;
    KbIluExtra* x = extra_of(pc);
    if (x) {
        if (x->owns_pattern) { KB_FREE(pc->l_rp); KB_FREE(pc->l_col); }
        kb_tiles_free(x->tiles); x->tiles = nullptr;
        kb_lean_free(x->lean); x->lean = nullptr;
        KB_FREE(x->level[0]); KB_FREE(x->level[1]); KB_FREE(x->counters); KB_FREE(x->ediag);
        for (int u = 0; u < 2; ++u) { KB_FREE(x->width[u]); KB_FREE(x->off[u]); KB_FREE(x->ecol[u]); KB_FREE(x->eval[u]); KB_FREE(x->spad[u]); KB_FREE(x->chunk_lev[u]); KB_FREE(x->done[u]); }
        delete x;
        pc->extra = nullptr;
    }
    KB_FREE(pc->lu); KB_FREE(pc->diag_ptr); KB_FREE(pc->tmp);
    for (int u = 0; u < 2; ++u) { KB_FREE(pc->level_ptr[u]); KB_FREE(pc->order[u]); KB_FREE(pc->sched[u]); }
}

static int build_levels(kb_pc_s* pc, int upper) {
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    KbIluExtra* x = extra_of(pc);
    const int n = (int)A->n;
    const unsigned nb = (unsigned)((n + 255) / 256);
    int* lev = nullptr;
    KB_TRY(kb_alloc(&lev, (size_t)n + 1));
    x->level[upper] = lev;
    KB_CUDA(cudaMemsetAsync(lev, 0, ((size_t)n + 1) * sizeof(int), c->stream));
    int* d_changed = nullptr;
    KB_TRY(kb_alloc(&d_changed, 1));
    // monotone relaxation to the fixpoint (<= nlevels sweeps; in-place reads make it converge faster)
    for (int round = 0;; ++round) {
        KB_CUDA(cudaMemsetAsync(d_changed, 0, sizeof(int), c->stream));
        for (int k = 0; k < 32; ++k) {
            KbLaunch L(c, KB_K_OTHER);
            k_ilu_level_sweep<<<nb, 256, 0, c->stream>>>(pc->l_rp, pc->l_col, pc->diag_ptr, n, upper, lev, d_changed);
        }
        int h = 0;
        KB_CUDA(cudaMemcpyAsync(&h, d_changed, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        KB_CUDA(cudaStreamSynchronize(c->stream));
        if (!h) break;
        if (round > (n / 32) + 8) { cudaFree(d_changed); kb_set_error("ilu0: level construction did not converge"); return KB_FACTOR_ERROR; }
    }
    cudaFree(d_changed);
    // stable sort rows by level (ascending row inside a level)
    int *keys_out = nullptr, *rows_in = nullptr, *rows_out = nullptr;
    KB_TRY(kb_alloc(&keys_out, (size_t)n)); KB_TRY(kb_alloc(&rows_in, (size_t)n)); KB_TRY(kb_alloc(&rows_out, (size_t)n));
    { KbLaunch L(c, KB_K_OTHER); k_iota<<<nb, 256, 0, c->stream>>>(rows_in, n); }
    int* d_max = nullptr;
    KB_TRY(kb_alloc(&d_max, 1));
    size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, lev, keys_out, rows_in, rows_out, n, 0, 32, c->stream);
    cub::DeviceReduce::Max(nullptr, tb2, lev, d_max, n, c->stream);
    void* tmp = nullptr;
    KB_CUDA(cudaMalloc(&tmp, std::max(tb, tb2) + 16));
    cub::DeviceReduce::Max(tmp, tb2, lev, d_max, n, c->stream);
    int hmax = 0;
    KB_CUDA(cudaMemcpyAsync(&hmax, d_max, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    int end_bit = 1;
    while ((1ll << end_bit) <= (long long)hmax) ++end_bit;
    cub::DeviceRadixSort::SortPairs(tmp, tb, lev, keys_out, rows_in, rows_out, n, 0, end_bit, c->stream);
    c->launches += 4;
    const int nlev = hmax + 1;
    pc->nlev[upper] = nlev;
    pc->order[upper] = rows_out;
    KB_TRY(kb_alloc(&pc->level_ptr[upper], (size_t)nlev + 1));
    { KbLaunch L(c, KB_K_OTHER); k_level_ptr<<<(unsigned)((nlev + 256) / 256), 256, 0, c->stream>>>(keys_out, n, nlev, pc->level_ptr[upper]); }
    // warp-padded schedule for the sync-free solves
    std::vector<int> lp((size_t)nlev + 1), sp((size_t)nlev + 1);
    KB_CUDA(cudaMemcpyAsync(lp.data(), pc->level_ptr[upper], ((size_t)nlev + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    sp[0] = 0;
    for (int l = 0; l < nlev; ++l) sp[l + 1] = sp[l] + ((lp[l + 1] - lp[l] + 31) / 32) * 32;
    pc->sched_len[upper] = sp[nlev];
    int* d_sp = nullptr;
    KB_TRY(kb_alloc(&d_sp, (size_t)nlev + 1));
    KB_CUDA(cudaMemcpyAsync(d_sp, sp.data(), ((size_t)nlev + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const size_t padded = ((size_t)sp[nlev] + KB_THREADS - 1) / KB_THREADS * KB_THREADS;
    KB_TRY(kb_alloc(&pc->sched[upper], padded + 1));
    KB_CUDA(cudaMemsetAsync(pc->sched[upper], 0xFF, (padded + 1) * sizeof(int), c->stream));   // -1 = empty slot
    { KbLaunch L(c, KB_K_OTHER); k_sched_fill<<<nlev, 256, 0, c->stream>>>(rows_out, pc->level_ptr[upper], d_sp, nlev, pc->sched[upper]); }
    KB_CUDA(cudaStreamSynchronize(c->stream));
    x->spad[upper] = d_sp;
    {
        const int nch = (sp[nlev] + KB_THREADS - 1) / KB_THREADS;
        std::vector<int> cl((size_t)nch + 1, 0);
        int l = 0;
        for (int k = 0; k < nch; ++k) { while (l + 1 < nlev && sp[l + 1] <= k * KB_THREADS) ++l; cl[k] = l; }
        KB_TRY(kb_alloc(&x->chunk_lev[upper], (size_t)nch + 1));
        KB_CUDA(cudaMemcpyAsync(x->chunk_lev[upper], cl.data(), ((size_t)nch + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        KB_TRY(kb_alloc(&x->done[upper], (size_t)nlev + 1));
        KB_CUDA(cudaMemsetAsync(x->done[upper], 0, ((size_t)nlev + 1) * sizeof(int), c->stream));
        KB_CUDA(cudaStreamSynchronize(c->stream));
    }
    cudaFree(tmp); cudaFree(d_max); cudaFree(keys_out); cudaFree(rows_in);
    return KB_OK;
}

int kb_ilu0_build(kb_pc_s* pc) {
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    const int n = (int)A->n;
    const unsigned nb = (unsigned)((n + 255) / 256);
    KbIluExtra* x = new KbIluExtra;
    pc->extra = x;
    KB_TRY(kb_alloc(&x->counters, 4));
    KB_CUDA(cudaMemsetAsync(x->counters, 0, 4 * sizeof(unsigned), c->stream));
    if (n == 0) return KB_OK;
    // 1. block-local pattern: ghost couplings dropped on a shard, aliased otherwise
    if (A->dist && A->ncols_local > A->n) {
        x->owns_pattern = true;
        int* cnt = nullptr;
        KB_TRY(kb_alloc(&cnt, (size_t)n + 1));
        KB_TRY(kb_alloc(&pc->l_rp, (size_t)n + 1));
        KB_CUDA(cudaMemsetAsync(cnt, 0, ((size_t)n + 1) * sizeof(int), c->stream));
        { KbLaunch L(c, KB_K_OTHER); k_ilu_count_local<<<nb, 256, 0, c->stream>>>(A->row_ptr, A->col, n, cnt); }
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, pc->l_rp, n + 1, c->stream);
        void* tmp = nullptr;
        KB_CUDA(cudaMalloc(&tmp, tb + 16));
        cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, pc->l_rp, n + 1, c->stream);
        c->launches += 1;
        int lnnz = 0;
        KB_CUDA(cudaMemcpyAsync(&lnnz, pc->l_rp + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        KB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(tmp); cudaFree(cnt);
        pc->l_nnz = (uint64_t)lnnz;
        KB_TRY(kb_alloc(&pc->l_col, (size_t)lnnz + 8));
        KB_TRY(kb_alloc(&pc->lu, (size_t)lnnz + 8));
        { KbLaunch L(c, KB_K_OTHER); k_ilu_fill_local<<<nb, 256, 0, c->stream>>>(A->row_ptr, A->col, A->vals, n, pc->l_rp, pc->l_col, pc->lu); }
    } else {
        pc->l_rp = A->row_ptr; pc->l_col = A->col; pc->l_nnz = A->nnz;
        KB_TRY(kb_alloc(&pc->lu, (size_t)A->nnz + 8));
        KB_CUDA(cudaMemcpyAsync(pc->lu, A->vals, A->nnz * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    // 2. diagonal positions (a row without a stored diagonal is a FactorError)
    unsigned long long* d_bad = nullptr;
    KB_TRY(kb_alloc(&d_bad, 1));
    KB_CUDA(cudaMemsetAsync(d_bad, 0xFF, sizeof(unsigned long long), c->stream));
    KB_TRY(kb_alloc(&pc->diag_ptr, (size_t)n + 1));
    KB_TRY(kb_alloc(&pc->inv_diag, (size_t)n + 2));
    KB_TRY(kb_alloc(&pc->tmp, (size_t)n + 2));
    { KbLaunch L(c, KB_K_OTHER); k_ilu_diag_ptr<<<nb, 256, 0, c->stream>>>(pc->l_rp, pc->l_col, n, pc->diag_ptr, d_bad); }
    unsigned long long h_bad = ~0ull;
    KB_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    if (h_bad != ~0ull) {
        cudaFree(d_bad);
        pc->bad_row = A->row_lo + h_bad;
        kb_set_error("ilu0: row %llu has no stored diagonal entry", (unsigned long long)pc->bad_row);
        return KB_FACTOR_ERROR;
    }
    // 3. level sets, built on the device
    KB_TRY(build_levels(pc, 0));
    KB_TRY(build_levels(pc, 1));
    // 4. numeric factorisation, one launch per lower level (a row needs only finished rows k < i of its pattern)
    {
        std::vector<int> lp((size_t)pc->nlev[0] + 1);
        KB_CUDA(cudaMemcpyAsync(lp.data(), pc->level_ptr[0], lp.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        KB_CUDA(cudaStreamSynchronize(c->stream));
        for (int l = 0; l < pc->nlev[0]; ++l) {
            const int cnt = lp[l + 1] - lp[l];
            if (cnt <= 0) continue;
            KbLaunch L(c, KB_K_OTHER);
            k_ilu_factor_level<<<(unsigned)((cnt + 127) / 128), 128, 0, c->stream>>>(pc->order[0] + lp[l], cnt, pc->l_rp, pc->l_col, pc->diag_ptr,
                                                                                pc->lu, pc->inv_diag, d_bad);
        }
    }
    KB_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_bad);
    if (h_bad != ~0ull) {
        pc->bad_row = A->row_lo + h_bad;
        kb_set_error("zero pivot at row %llu", (unsigned long long)pc->bad_row);
        return KB_ZERO_PIVOT;
    }
    // 5. grid-structured pattern (5-/7-point box stencil, detected from the CSR): block-wavefront solves
    if (!(getenv("KB_TRSV_TILES") && atoi(getenv("KB_TRSV_TILES")) == 0)) KB_TRY(kb_tiles_build(pc, x->counters + 2, &x->tiles));
    // Pencil-marching warps (kb_trsv_lean.cu) when the stencil is full.  Measured on B200, per solve: 3-D 256^3 0.30 ms against
    // 0.72 ms for the tiles and 1.45 ms level-scheduled; 2-D 1024^2 0.35 ms against 0.98 / 2.3 ms.  KB_TRSV_MARCH=0 switches it off.
    if (x->tiles) {
        int gx = 0, gy = 0, gz = 0;
        kb_tiles_grid(x->tiles, &gx, &gy, &gz);
        const int want = getenv("KB_TRSV_MARCH") ? atoi(getenv("KB_TRSV_MARCH")) : -1;
        if (want != 0) KB_TRY(kb_lean_build(pc, gx, gy, gz, x->counters + 2, &x->lean));
    }
    // 6. level-ordered chunk-ELL copies of L and U for the persistent (general-pattern) solves
    if (getenv("KB_TRSV_KIND")) x->kind = atoi(getenv("KB_TRSV_KIND"));
    if (getenv("KB_TRSV_SLEEP")) x->sleep_ns = atoi(getenv("KB_TRSV_SLEEP"));
    for (int u = 0; u < 2 && x->kind == 1; ++u) {
        const int nch = (pc->sched_len[u] + KB_THREADS - 1) / KB_THREADS;
        x->nchunks[u] = nch;
        KB_TRY(kb_alloc(&x->width[u], (size_t)nch + 1));
        KB_TRY(kb_alloc(&x->off[u], (size_t)nch + 1));
        { KbLaunch L(c, KB_K_OTHER); k_ell_width<<<nch, KB_THREADS, 0, c->stream>>>(pc->sched[u], nch, pc->l_rp, pc->diag_ptr, u, x->width[u]); }
        std::vector<int> hw((size_t)nch);
        KB_CUDA(cudaMemcpyAsync(hw.data(), x->width[u], (size_t)nch * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        KB_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<long long> ho((size_t)nch + 1);
        long long acc = 0;
        for (int k = 0; k < nch; ++k) { ho[k] = acc; acc += (long long)hw[k] * KB_THREADS; }
        ho[nch] = acc;
        KB_CUDA(cudaMemcpyAsync(x->off[u], ho.data(), ((size_t)nch + 1) * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
        KB_TRY(kb_alloc(&x->ecol[u], (size_t)acc + 8));
        KB_TRY(kb_alloc(&x->eval[u], (size_t)acc + 8));
        if (u == 1) KB_TRY(kb_alloc(&x->ediag, (size_t)nch * KB_THREADS + 8));
        { KbLaunch L(c, KB_K_OTHER); k_ell_fill<<<nch, KB_THREADS, 0, c->stream>>>(pc->sched[u], pc->l_rp, pc->l_col, pc->diag_ptr, pc->lu, pc->inv_diag, u,
                                                                                  x->width[u], x->off[u], x->ecol[u], x->eval[u], u == 1 ? x->ediag : nullptr); }
        KB_CUDA(cudaStreamSynchronize(c->stream));
        // window of in-flight chunks ~ a few levels wide, between 1 and 8 CTAs per SM
        if (getenv("KB_TRSV_GATE")) x->gate = atoi(getenv("KB_TRSV_GATE"));
        int g = 3 * c->sm_count;   // measured plateau on B200: 3-5 CTAs/SM; more only adds gate pollers
        // every CTA of the grid must be co-resident (a chunk may wait on a chunk of any other CTA)
        int occ = 1;
        if (u == 0) KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb_trsv_persistent<false>, KB_THREADS, 0));
        else KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb_trsv_persistent<true>, KB_THREADS, 0));
        const int cap = std::max(1, occ) * c->sm_count;
        g = std::max(c->sm_count, g);
        if (getenv("KB_TRSV_GRID")) g = atoi(getenv("KB_TRSV_GRID"));
        x->grid[u] = std::max(1, std::min(std::min(nch, g), cap));
    }
    return KB_OK;
}

int kb_ilu0_apply_dev(kb_pc_s* pc, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask) {
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    if (A->n == 0) return KB_OK;
    KbIluExtra* x = extra_of(pc);
    if (x->lean) return kb_lean_apply(pc, x->lean, d_r, d_z, skip_ctl, skip_mask);
    if (x->tiles) return kb_tiles_apply(pc, x->tiles, d_r, d_z, skip_ctl, skip_mask);
    {
        KbLaunch L(c, KB_K_TRSV);
        kb_trsv_fill<<<(unsigned)((A->n + 255) / 256), 256, 0, c->stream>>>(pc->tmp, d_z, (long long)A->n, skip_ctl, skip_mask, x->done[0], pc->nlev[0], x->done[1], pc->nlev[1]);
    }
    if (x->kind == 1) {
        KbTrsvEll e{};
        e.counters = x->counters; e.skip_ctl = skip_ctl; e.skip_mask = skip_mask; e.sleep_ns = x->sleep_ns;
        {
            e.sched = pc->sched[0]; e.nchunks = x->nchunks[0]; e.width = x->width[0]; e.off = x->off[0]; e.ecol = x->ecol[0]; e.eval = x->eval[0];
            e.ediag = nullptr; e.rhs = d_r; e.out = pc->tmp;
            e.spad = x->spad[0]; e.chunk_lev = x->chunk_lev[0]; e.done = x->done[0]; e.nlev = pc->nlev[0]; e.gate = x->gate;
            KbLaunch L(c, KB_K_TRSV);
            kb_trsv_persistent<false><<<x->grid[0], KB_THREADS, 0, c->stream>>>(e);
        }
        {
            e.sched = pc->sched[1]; e.nchunks = x->nchunks[1]; e.width = x->width[1]; e.off = x->off[1]; e.ecol = x->ecol[1]; e.eval = x->eval[1];
            e.ediag = x->ediag; e.rhs = pc->tmp; e.out = d_z;
            e.spad = x->spad[1]; e.chunk_lev = x->chunk_lev[1]; e.done = x->done[1]; e.nlev = pc->nlev[1]; e.gate = x->gate;
            KbLaunch L(c, KB_K_TRSV);
            kb_trsv_persistent<true><<<x->grid[1], KB_THREADS, 0, c->stream>>>(e);
        }
        KB_CUDA(cudaGetLastError());
        return KB_OK;
    }
    KbTrsvArgs a{};
    a.skip_ctl = skip_ctl; a.skip_mask = skip_mask;
    a.lrp = pc->l_rp; a.lcol = pc->l_col; a.dp = pc->diag_ptr; a.lu = pc->lu; a.inv_ud = pc->inv_diag; a.counters = x->counters;
    {
        a.sched = pc->sched[0]; a.sched_len = pc->sched_len[0]; a.rhs = d_r; a.out = pc->tmp;
        KbLaunch L(c, KB_K_TRSV);
        kb_trsv_syncfree<false><<<(unsigned)((a.sched_len + KB_THREADS - 1) / KB_THREADS), KB_THREADS, 0, c->stream>>>(a);
    }
    {
        a.sched = pc->sched[1]; a.sched_len = pc->sched_len[1]; a.rhs = pc->tmp; a.out = d_z;
        KbLaunch L(c, KB_K_TRSV);
        kb_trsv_syncfree<true><<<(unsigned)((a.sched_len + KB_THREADS - 1) / KB_THREADS), KB_THREADS, 0, c->stream>>>(a);
    }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// 1 when a bounded spin of a triangular solve expired (a dependency never arrived): the solve's result is invalid
int kb_ilu0_error(kb_pc_s* pc) {
    if (pc && pc->kind == KB_PC_ASM) return kb_asm_error(pc);      // any inner block
    KbIluExtra* x = pc && pc->kind == KB_PC_ILU0 ? extra_of(pc) : nullptr;
    if (!x || !x->counters) return 0;
    unsigned e = 0;
    kb_ctx_s* c = pc->ctx;
    if (cudaMemcpyAsync(&e, x->counters + 2, sizeof(e), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) return 1;
    if (e) {    // reported once, then cleared: later applies on this preconditioner start clean
        cudaMemsetAsync(x->counters, 0, 4 * sizeof(unsigned), c->stream);
        cudaStreamSynchronize(c->stream);
    }
    return e != 0;
}

extern "C" int kb_pc_create_ilu0(kb_csr A, kb_pc* out) {
    *out = nullptr;
    if (!A) { kb_set_error("null operator"); return KB_SOLVE_ERROR; }
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    if (A->n != A->ncols_global && !A->dist) { kb_set_error("ilu0 needs a square operator"); return KB_FACTOR_ERROR; }
    kb_pc_s* pc = new kb_pc_s;
    pc->a = A; pc->ctx = c; pc->kind = KB_PC_ILU0;
    A->refs++;
    int st = kb_ilu0_build(pc);
    if (st != KB_OK) {
        if (st == KB_ZERO_PIVOT || st == KB_FACTOR_ERROR) { *out = pc; return st; }   // caller may query kb_pc_bad_row, then destroy
        kb_pc_destroy(pc);
        return st;
    }
    *out = pc;
    return KB_OK;
}

extern "C" int kb_pc_ilu0_get_factors(kb_pc pc, double* lu, uint64_t* diag_ptr) {
    if (!pc || pc->kind != KB_PC_ILU0) { kb_set_error("not an ILU(0) preconditioner"); return KB_SOLVE_ERROR; }
    kb_csr_s* A = pc->a;
    KB_CUDA(cudaSetDevice(A->ctx->device));
    if (lu && pc->l_nnz) KB_CUDA(cudaMemcpy(lu, pc->lu, pc->l_nnz * sizeof(double), cudaMemcpyDeviceToHost));
    if (diag_ptr && A->n) {
        std::vector<int> dp(A->n);
        KB_CUDA(cudaMemcpy(dp.data(), pc->diag_ptr, A->n * sizeof(int), cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < A->n; ++i) diag_ptr[i] = (uint64_t)dp[i];
    }
    return KB_OK;
}
extern "C" int kb_pc_ilu0_get_levels(kb_pc pc, int upper, uint64_t* nlevels, uint64_t* level_ptr, uint64_t* order) {
    if (!pc || pc->kind != KB_PC_ILU0) { kb_set_error("not an ILU(0) preconditioner"); return KB_SOLVE_ERROR; }
    kb_csr_s* A = pc->a;
    KB_CUDA(cudaSetDevice(A->ctx->device));
    const int u = upper ? 1 : 0;
    *nlevels = (uint64_t)pc->nlev[u];
    if (A->n == 0) return KB_OK;
    std::vector<int> lp((size_t)pc->nlev[u] + 1), od(A->n);
    KB_CUDA(cudaMemcpy(lp.data(), pc->level_ptr[u], lp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    KB_CUDA(cudaMemcpy(od.data(), pc->order[u], od.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t l = 0; l < lp.size(); ++l) level_ptr[l] = (uint64_t)lp[l];
    for (size_t i = 0; i < od.size(); ++i) order[i] = (uint64_t)od[i];
    return KB_OK;
}

// diagnostics for tuning scripts (deliberately not declared in include/kryst_b200.h)
extern "C" int kb_debug_march_trace(kb_pc pc, unsigned long long* out, int* px, int* py) {
    KbIluExtra* x = pc && pc->kind == KB_PC_ILU0 ? extra_of(pc) : nullptr;
    if (!x || !x->lean) return 0;
    cudaSetDevice(pc->ctx->device);
    cudaStreamSynchronize(pc->ctx->stream);
    return kb_lean_trace_get(x->lean, out, px, py);
}

