// kb_p2p.cuh — NVLink peer-memory collectives (one process per GPU, CUDA-IPC mapped mailboxes).
//
// The Krylov loop needs two exchanges: 1-51 doubles of all-reduce per reduction and one ghost
// plane per neighbour per SpMV.  Both are latency-bound, so instead of NCCL launches (all-gather + sum +
// epilogue = 3 launches, ~10 us each) they are single kernels that store straight into the peers' HBM through
// NVLink/NVSwitch and spin on sequence-numbered flags:
//   * kb_p2p_allreduce_block : every rank writes its `count` partial sums into slot [parity][rank] of EVERY
//     rank's mailbox, fences (system scope), writes the sequence number into the flag slot, waits until all p
//     flags of its own mailbox carry the number, then adds the p contributions in rank order — the same
//     deterministic order as the oracle's sharded reduction.  Double-buffered by parity: a rank can be at most
//     one reduction ahead of any other, because it needs everybody's contribution to pass one.
//   * halo push / receive     : see kb_dist.cu.
// All spins are bounded; on timeout an error flag is raised and the solve returns KB_SOLVE_ERROR instead of
// hanging the GPU.
#pragma once
#include "kb_internal.cuh"

#define KB_MAX_RANKS 16
#define KB_AR_MAX (KB_MAX_RESTART + 8)
#define KB_SPIN_LIMIT (1u << 25)

struct KbP2PDev {
    int rank, size;
    double* vals[KB_MAX_RANKS];                 // mailbox of rank q: vals[2][size][KB_AR_MAX]
    unsigned long long* flags[KB_MAX_RANKS];    // mailbox of rank q: flags[2][size]
    unsigned long long* seq;                    // local device counter of reductions performed
    unsigned* err;                              // local error flag
};

#ifdef __CUDACC__
// BAR == 0: __syncthreads(); BAR == 1: named barrier 1 over the 256 consumer threads of the bulk SpMV
template <int BAR>
__device__ __forceinline__ void kb_sync() {
    if (BAR == 0) __syncthreads();
    else asm volatile("bar.sync 1, 256;" ::: "memory");
}
// in place on `inout` (global or shared), threads 0..255 of ONE block must call it
template <int BAR = 0>
__device__ __forceinline__ void kb_p2p_allreduce_block(const KbP2PDev& p, double* inout, int count) {
    __shared__ unsigned long long s_seq;
    const int tid = threadIdx.x;
    if (tid == 0) { s_seq = *p.seq + 1ull; *p.seq = s_seq; }
    kb_sync<BAR>();
    const unsigned long long seq = s_seq;
    const size_t par = (size_t)(seq & 1ull);
    for (int idx = tid; idx < p.size * count; idx += KB_THREADS) {
        const int q = idx / count, r = idx - q * count;
        p.vals[q][(par * p.size + p.rank) * KB_AR_MAX + r] = inout[r];
    }
    kb_sync<BAR>();
    if (tid < p.size) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(p.flags[tid] + par * p.size + p.rank) = seq;
        const volatile unsigned long long* f = p.flags[p.rank] + par * p.size + tid;
        unsigned spins = 0;
        while (*f < seq) {
            if (++spins > KB_SPIN_LIMIT) { atomicExch(p.err, 1u); break; }
        }
        __threadfence_system();
    }
    kb_sync<BAR>();
    for (int r = tid; r < count; r += KB_THREADS) {
        const volatile double* v = p.vals[p.rank] + (par * p.size) * KB_AR_MAX + r;
        double s = v[0];
        for (int q = 1; q < p.size; ++q) s = s + v[(size_t)q * KB_AR_MAX];
        inout[r] = s;
    }
    kb_sync<BAR>();
}

__global__ void __launch_bounds__(KB_THREADS) kb_p2p_allreduce_kernel(KbP2PDev p, double* vals, int count);

// all-reduce + scalar epilogue of a solver in ONE launch (replaces all-gather, rank-ordered sum and epilogue kernels)
template <class Fin>
__global__ void __launch_bounds__(KB_THREADS) kb_p2p_allreduce_fin(KbP2PDev p, Fin fin, KbCtl* ctl, double* slots, int nred, int skip_early) {
    // the collective itself must run on every rank even when this rank's solve is `done`: all ranks take
    // identical decisions, so either all skip or none does.
    if (ctl->done || (skip_early && ctl->early)) return;
    kb_p2p_allreduce_block<0>(p, slots, nred);
    if (threadIdx.x == 0) fin(slots);
}
#endif
