// kb_internal.cuh — shared internals of libkryst_b200 (not part of the ABI).
//
// Numerical contract: every kernel performs the *same sequence of IEEE f64 operations* as
// the CPU oracle (see oracle/): separate mul and add (this library MUST be
// compiled with -fmad=false), ascending-column row sums, and the canonical reduction tree R:
//   level 1: tiles of 512 elements, lane l of 256 holds e(2l)+e(2l+1); 32-lane xor
//            butterflies (16,8,4,2,1); the 8 warp sums added sequentially;
//   level 2: lane l = 0.0 + u[l] + u[l+256] + ... over the tile sums, same butterfly;
//   ranks  : sequential sum in rank order.
// Hence GPU results are bit-identical to the oracle, and independent of grid scheduling.
#pragma once
#ifndef KB_NO_FMA
#error "libkryst_b200 must be built with -fmad=false -DKB_NO_FMA (bit-parity with the oracle)"
#endif
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <string>
#include <vector>
#include <set>
#include "kryst_b200.h"

#define KB_TILE 512
#define KB_THREADS 256

void kb_set_error(const char* fmt, ...);
#define KB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            kb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
            return KB_SOLVE_ERROR;                                                                 \
        }                                                                                          \
    } while (0)
#define KB_TRY(call)                                                                               \
    do {                                                                                           \
        int s_ = (call);                                                                           \
        if (s_ != KB_OK) return s_;                                                                \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Device-resident control block: all per-iteration scalars live here so that a solve needs no
// host round trip per iteration (host polls `done` once per replayed batch).
// ---------------------------------------------------------------------------------------------
#define KB_MAX_RESTART 128
struct KbCtl {
    // --- common (mirrors SolveStats + KError)
    int done;                  // 1 => every later kernel of this solve is a no-op
    int status;                // kb_status
    int converged;
    int breakdown;
    unsigned long long iter;   // iterations completed / reported
    unsigned long long max_iters;
    double tol;
    double res0;               // denominator of Convergence::check
    double res;                // last residual norm
    unsigned long long hist_len, hist_cap;
    double* hist;              // device residual history (residual_history pushes)
    // --- PCG (pcg.rs:114-222)
    double rz, pAp, alpha, beta, rz_new;
    double sr_loc[3];          // single-reduction / pipelined PCG: this rank's sums of the update kernel, folded into the SpMV's all-reduce
    int norm_type;
    // --- BiCGStab (bicgstab.rs:69-293)
    double rho, rho_prev, omega, omega_prev, alpha_den, thr, rnorm, vv;
    int textbook;
    int early;                 // ||s|| <= tol: finish with x += alpha p (bicgstab.rs:189-206)
    // --- GMRES (gmres.rs:216-402)
    int j;                     // inner index of the current Arnoldi step
    int m;                     // columns accumulated in this cycle (for back-substitution)
    int restart;
    int cycle_break;           // inner loop of this cycle has exited
    int happy;
    int hflag;                 // FGMRES: happy breakdown of the current step only
    int side;
    int outer, n_outer;
    double res0_true, beta_g, hnorm;
    double h[(KB_MAX_RESTART + 1) * KB_MAX_RESTART];   // column-major: h[i + (restart+1)*j]
    double g[KB_MAX_RESTART + 1], cs[KB_MAX_RESTART], sn[KB_MAX_RESTART], y[KB_MAX_RESTART];
    double h1[KB_MAX_RESTART + 1], h2[KB_MAX_RESTART + 1];
};

// ---------------------------------------------------------------------------------------------
// host-side context
// ---------------------------------------------------------------------------------------------
struct KbProfEvent { int cls; cudaEvent_t a, b; };

struct kb_ctx_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    uint64_t launches = 0;
    // reduction scratch shared by all kernels on this stream
    unsigned* ticket = nullptr;          // device counter for last-block detection
    // profiling
    bool profiling = false;
    std::vector<KbProfEvent> prof_events;
    std::vector<cudaEvent_t> event_pool;
    kb_profile prof_acc{};
    // stream capture bookkeeping
    bool capturing = false;
    uint64_t captured_launches = 0;
    // communicator (one process per GPU)
    int rank = 0, size = 1;
    void* nccl = nullptr;                // ncclComm_t
    void* p2p = nullptr;                 // KbP2PHost: IPC-mapped all-reduce mailboxes (NVLink peer path)
    double* comm_buf = nullptr;          // device scratch for all-gathered partial scalars
    double* host_scalar = nullptr;       // pinned
    int refs = 1;                        // the user handle + one per operator created on this context
    std::set<const void*> configured;    // kernels whose dynamic-smem attribute has been raised
};

// RAII launch bookkeeping: counts the launch and (in profile mode) brackets it with events.
struct KbLaunch {
    kb_ctx_s* c; int cls; cudaEvent_t a = nullptr, b = nullptr;
    KbLaunch(kb_ctx_s* ctx, int k);
    ~KbLaunch();
};
int kb_prof_collect(kb_ctx_s* c);

// ---------------------------------------------------------------------------------------------
// device: canonical reduction
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double kb_warp_butterfly(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// 256-thread block: combine NRED lane values into tile sums (valid in thread 0 only).
template <int NRED>
__device__ __forceinline__ void kb_block_reduce(double (&v)[NRED], double* sm /* NRED*8 */, double (&out)[NRED]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < NRED; ++r) {
        double t = kb_warp_butterfly(v[r]);
        if (lane == 0) sm[r * 8 + w] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < NRED; ++r) {
            double s = sm[r * 8];
#pragma unroll
            for (int k = 1; k < 8; ++k) s = s + sm[r * 8 + k];
            out[r] = s;
        }
    }
    __syncthreads();
}

// level 2 over P tile sums (all 256 threads of one block participate; result in thread 0)
__device__ __forceinline__ double kb_level2(const double* __restrict__ part, int P, double* sm8) {
    double acc = 0.0;
#pragma unroll 8
    for (int k = threadIdx.x; k < P; k += KB_THREADS) acc = acc + __ldcg(part + k);
    double v[1] = {acc}, out[1] = {0.0};
    kb_block_reduce<1>(v, sm8, out);
    return out[0];
}

// Thread 0 has already stored this block's partial(s).  Returns true in every thread of the
// block that arrived last (all partials of the grid are then visible to it).
__device__ __forceinline__ bool kb_arrive_last(unsigned* ticket, unsigned nblocks, int* sflag) {
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        *sflag = (t == nblocks - 1u);
        if (*sflag) { *ticket = 0u; __threadfence(); }
    }
    __syncthreads();
    return *sflag != 0;
}

// Programmatic dependent launch (PDL): the kernels of an iteration are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs are scheduled (and run their prologue)
// while the current kernel drains, instead of after a full kernel boundary (~2-3 us per boundary inside a CUDA graph;
// an iteration of the 512^2 problem is three kernels of ~3 us of work each).  Every kernel on such a chain starts with
// kb_pdl_wait() - nothing written by the previous kernel (control block included) is read before it - and releases its
// own dependents right away.  For a kernel launched without the attribute both calls are no-ops.
__device__ __forceinline__ void kb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void kb_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Generic fused BLAS-1 "tile kernel": one 512-element tile per block, 2 adjacent elements per
// thread (128-bit accesses), NRED fused canonical dots, scalar epilogue by the last block.
//   Op::NRED                      number of fused reductions
//   bool Op::skip() const         early-out (e.g. ctl->done)
//   void Op::pair(i, has1, red)   process elements i (and i+1 if has1); red[r] = e_r(i) + e_r(i+1)
//   void Op::finish_block(ssum)   epilogue, called by ALL threads of the last block (sums in shared memory)
// ---------------------------------------------------------------------------------------------
template <class Op>
__global__ void __launch_bounds__(KB_THREADS) kb_tile_kernel(Op op) {
    kb_pdl_wait();
    kb_pdl_launch_dependents();
    if (op.skip()) return;
    constexpr int NR = Op::NRED > 0 ? Op::NRED : 1;
    __shared__ double sm[NR * 8];
    __shared__ int sflag;
    const long long i = (long long)blockIdx.x * KB_TILE + 2 * threadIdx.x;
    double red[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) red[r] = 0.0;
    if (i + 1 < op.n) op.pair(i, true, red);
    else if (i < op.n) op.pair(i, false, red);
    else {
#pragma unroll
        for (int r = 0; r < NR; ++r) red[r] = 0.0 + 0.0;
    }
    if constexpr (Op::NRED > 0) {
        double out[NR];
        kb_block_reduce<NR>(red, sm, out);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int r = 0; r < NR; ++r) op.partials[(size_t)r * op.pstride + blockIdx.x] = out[r];
        }
        if (kb_arrive_last(op.ticket, gridDim.x, &sflag)) {
            double sums[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) sums[r] = kb_level2(op.partials + (size_t)r * op.pstride, (int)gridDim.x, sm);
            __shared__ double ssum[NR];
            if (threadIdx.x == 0) {
#pragma unroll
                for (int r = 0; r < NR; ++r) ssum[r] = sums[r];
            }
            __syncthreads();
            op.finish_block(ssum);      // every thread of the last block (scalar epilogue / fused all-reduce)
        }
    }
}

// shared skip predicate of solver-embedded helper kernels
__device__ __forceinline__ bool kb_skip(const KbCtl* c, int mask) {
    return c != nullptr && (c->done != 0 || ((mask & 1) && c->early != 0) || ((mask & 2) && c->cycle_break != 0));
}

struct KbRedBase {
    long long n;
    double* partials;
    size_t pstride;
    unsigned* ticket;
};

__device__ __forceinline__ double2 kb_ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void kb_st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
#endif  // __CUDACC__

static inline int kb_num_tiles(uint64_t n) { return (int)((n + KB_TILE - 1) / KB_TILE); }

#ifdef __CUDACC__
// launch with (pdl = true) or without the programmatic-dependent-launch attribute
template <class... KArgs, class... Args>
static inline cudaError_t kb_launch_ex(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
// Measured on B200 (CUDA-graph replay): no gain at 512^2 (54.5k -> 53.7k it/s) and a 5 % loss at 256^3 (1835 -> 1747
// it/s: dependents that become resident early take shared memory and issue slots from the draining kernel), so the
// attribute is opt-in (KB_PDL=1); the griddepcontrol instructions stay in the kernels and are no-ops without it.
static inline bool kb_pdl_enabled() {
    static const bool on = getenv("KB_PDL") && atoi(getenv("KB_PDL")) == 1;
    return on;
}
#endif
