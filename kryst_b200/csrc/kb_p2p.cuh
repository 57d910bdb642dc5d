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
    double* vals[KB_MAX_RANKS];                 // mailbox of rank q: 16-byte packets [2][size][KB_AR_MAX]
    unsigned long long* flags[KB_MAX_RANKS];    // mailbox of rank q: flags[2][size]
    unsigned long long* seq;                    // local device counter of reductions performed
    unsigned* err;                              // local error flag
};

struct KbHaloDev {                 // device view of the peer-memory halo exchange (kb_dist.cu)
    int rank, size, nsend, nghost, n_loc;
    long long gstride;                           // my ghost_in parity stride (doubles)
    long long peer_gstride[KB_MAX_RANKS];
    double* ghost[KB_MAX_RANKS];                 // rank q's ghost_in[2][gstride_q]
    unsigned long long* flags[KB_MAX_RANKS];     // rank q's flags[2][size]   : flags[q][par*size + src] = seq pushed by src
    unsigned long long* acks[KB_MAX_RANKS];      // rank q's acks[size]       : acks[q][dst] = last push of q consumed by dst
    int is_dest[KB_MAX_RANKS], is_src[KB_MAX_RANKS];
    const int* send_idx; const int* send_q; const int* send_pos;
    unsigned long long* seqs;                    // [0] pushes done, [1] receives done (device counters)
    unsigned* tickets;                           // [0] push, [1] recv last-block tickets
    unsigned* err;
    // the same send list sorted by canonical tile of the source row: lets the kernel that PRODUCES the SpMV operand
    // (PCG's p = z + beta p) store its boundary entries straight into the neighbours' mailboxes (fused push)
    const int* tile_send_ptr;                    // [ntiles + 1]
    const int* ts_idx; const int* ts_q; const int* ts_pos;
    int n_send_tiles;                            // tiles with at least one entry (the fused push's last-CTA ticket count)
    const int* tile_perm;                        // [ntiles] CTA -> tile, sending tiles first: their NVLink latency hides behind the rest of the grid
};

#ifdef __CUDACC__
// BAR == 0: __syncthreads(); BAR == 1: named barrier 1 over the 256 consumer threads of the bulk SpMV
template <int BAR>
__device__ __forceinline__ void kb_sync() {
    if (BAR == 0) __syncthreads();
    else asm volatile("bar.sync 1, 256;" ::: "memory");
}
// in place on `inout` (global or shared), threads 0..255 of ONE block must call it.
// Every partial sum travels as ONE 16-byte packet {lo32 | tag, hi32 | tag} (tag = low 32 bits of the reduction's
// sequence number): the data is its own flag, so there is no system-scope fence between "data" and "flag" stores and a
// reduction costs one NVLink one-way latency instead of store + fence round trip + flag (the NCCL-LL idea).
__device__ __forceinline__ void kb_ll_store(ulonglong2* p, double v, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ __forceinline__ bool kb_ll_load(const ulonglong2* p, unsigned tag, double* v) {
    unsigned long long x, y;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
    *v = __longlong_as_double((long long)((x & 0xffffffffull) | (y << 32)));
    return (unsigned)(x >> 32) == tag && (unsigned)(y >> 32) == tag;
}
template <int BAR = 0>
__device__ __forceinline__ void kb_p2p_allreduce_block(const KbP2PDev& p, double* inout, int count) {
    __shared__ unsigned long long s_seq;
    const int tid = threadIdx.x;
    if (tid == 0) { s_seq = *p.seq + 1ull; *p.seq = s_seq; }
    kb_sync<BAR>();
    const unsigned long long seq = s_seq;
    const unsigned tag = (unsigned)seq;                    // never 0 on its first use; a slot is reused every 2nd reduction
    const size_t par = (size_t)(seq & 1ull);
    for (int idx = tid; idx < p.size * count; idx += KB_THREADS) {
        const int q = idx / count, r = idx - q * count;
        kb_ll_store(reinterpret_cast<ulonglong2*>(p.vals[q]) + (par * p.size + p.rank) * KB_AR_MAX + r, inout[r], tag);
    }
    kb_sync<BAR>();                                        // all of inout[] has been read before anybody overwrites it
    for (int r = tid; r < count; r += KB_THREADS) {
        const ulonglong2* v = reinterpret_cast<const ulonglong2*>(p.vals[p.rank]) + (par * p.size) * KB_AR_MAX + r;
        double s = 0.0;
        for (int q = 0; q < p.size; ++q) {                 // rank order: the oracle's sharded reduction
            double x;
            unsigned spins = 0;
            while (!kb_ll_load(v + (size_t)q * KB_AR_MAX, tag, &x)) {
                if (++spins > KB_SPIN_LIMIT) { atomicExch(p.err, 1u); break; }
            }
            s = q == 0 ? x : s + x;
        }
        inout[r] = s;
    }
    kb_sync<BAR>();
}

// The two halves of kb_p2p_allreduce_block as separate calls (pipelined PCG): the kernel that produced the sums SENDS
// them, the last CTA of the following SpMV RECEIVES - the NVLink flight time passes while the SpMV runs.
// No other reduction may start between the two (the receive reads the sequence number the send advanced).
template <int BAR = 0>
__device__ __forceinline__ void kb_p2p_allreduce_send(const KbP2PDev& p, const double* in, int count) {
    __shared__ unsigned long long s_seq_s;
    const int tid = threadIdx.x;
    if (tid == 0) { s_seq_s = *p.seq + 1ull; *p.seq = s_seq_s; }
    kb_sync<BAR>();
    const unsigned long long seq = s_seq_s;
    const unsigned tag = (unsigned)seq;
    const size_t par = (size_t)(seq & 1ull);
    for (int idx = tid; idx < p.size * count; idx += KB_THREADS) {
        const int q = idx / count, r = idx - q * count;
        kb_ll_store(reinterpret_cast<ulonglong2*>(p.vals[q]) + (par * p.size + p.rank) * KB_AR_MAX + r, in[r], tag);
    }
    kb_sync<BAR>();
}
template <int BAR = 0>
__device__ __forceinline__ void kb_p2p_allreduce_recv(const KbP2PDev& p, double* out, int count) {
    const int tid = threadIdx.x;
    const unsigned long long seq = *reinterpret_cast<const volatile unsigned long long*>(p.seq);      // advanced by the send (an earlier kernel)
    const unsigned tag = (unsigned)seq;
    const size_t par = (size_t)(seq & 1ull);
    for (int r = tid; r < count; r += KB_THREADS) {
        const ulonglong2* v = reinterpret_cast<const ulonglong2*>(p.vals[p.rank]) + (par * p.size) * KB_AR_MAX + r;
        double s = 0.0;
        for (int q = 0; q < p.size; ++q) {                 // rank order: the oracle's sharded reduction
            double x;
            unsigned spins = 0;
            while (!kb_ll_load(v + (size_t)q * KB_AR_MAX, tag, &x)) {
                if (++spins > KB_SPIN_LIMIT) { atomicExch(p.err, 1u); break; }
            }
            s = q == 0 ? x : s + x;
        }
        out[r] = s;
    }
    kb_sync<BAR>();
}

__global__ void __launch_bounds__(KB_THREADS) kb_p2p_allreduce_kernel(KbP2PDev p, double* vals, int count);

// Halo push fused into the kernel that produces the SpMV operand: called by all 256 threads of the CTA that owns
// canonical tile `tile`, after they have stored x[tile].  Entries of this tile go straight into the destination
// GPUs' ghost_in[parity] over NVLink; the last sending CTA acknowledges the previous exchange to its sources and
// publishes the sequence number to the destinations (same protocol as kb_halo_push, kb_dist.cu).
__device__ __forceinline__ void kb_halo_push_tile(const KbHaloDev& h, const double* x, int tile) {
    const int k0 = h.tile_send_ptr[tile], k1 = h.tile_send_ptr[tile + 1];
    if (k1 <= k0) return;                                   // CTA-uniform
    __shared__ int s_last_push;
    const int tid = threadIdx.x;
    const unsigned long long seq = *reinterpret_cast<const volatile unsigned long long*>(h.seqs) + 1ull;   // stable until the last CTA below
    const size_t par = (size_t)(seq & 1ull);
    if (tid < h.size && h.is_dest[tid]) {                   // flow control: the parity buffer of push seq-2 has been consumed
        const volatile unsigned long long* a = h.acks[h.rank] + tid;
        unsigned spins = 0;
        while (*a + 2ull < seq) { if (++spins > KB_SPIN_LIMIT) { atomicExch(h.err, 1u); break; } }
    }
    __syncthreads();                                        // also orders this CTA's stores of x before the reads below
    for (int k = k0 + tid; k < k1; k += KB_THREADS) {
        const int q = h.ts_q[k];
        h.ghost[q][par * h.peer_gstride[q] + h.ts_pos[k]] = x[h.ts_idx[k]];
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        const unsigned t = atomicAdd(&h.tickets[0], 1u);
        s_last_push = (t == (unsigned)h.n_send_tiles - 1u);
    }
    __syncthreads();
    if (s_last_push) {
        if (tid < h.size) {
            __threadfence_system();
            // every kernel that read the ghosts of exchange seq-1 finished before this one started (stream order)
            if (h.is_src[tid] && seq > 1ull) *reinterpret_cast<volatile unsigned long long*>(h.acks[tid] + h.rank) = seq - 1ull;
            if (h.is_dest[tid]) *reinterpret_cast<volatile unsigned long long*>(h.flags[tid] + par * h.size + h.rank) = seq;
        }
        if (tid == 0) { *reinterpret_cast<volatile unsigned long long*>(h.seqs) = seq; h.tickets[0] = 0u; __threadfence(); }
    }
}

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
