// kb_pcg_mega.cuh — the whole PCG loop as ONE persistent cooperative kernel (single GPU).
//
// For small / L2-resident problems (C1: 44 MB per iteration = 7 us of HBM time) three launches per iteration
// cost more than the work.  Here 2 CTAs per SM stay resident for the entire solve and walk the iterations with
// grid barriers; the arithmetic, tile ownership and reduction tree are exactly those of the three-kernel path
// (kb_spmv_bulk + PcgUpdateOp + PcgXpayOp), so results are bit-identical to it and to the oracle.
//   per iteration:  SpMV (bulk-async staged) + p.Ap partials | barrier + last-CTA epilogue (alpha, pAp<=0 test)
//                   x,r,z update + r.z, ||r||^2 partials      | barrier + last-CTA epilogue (history, stop, beta)
//                   p = z + beta p                            | barrier
// The last CTA to arrive at a barrier runs the level-2 sums and the scalar epilogue before releasing it.
#pragma once
#include "kb_spmv_bulk.cuh"

struct KbPcgMegaArgs {
    KbCtl* ctl;
    double* x; double* r; double* z; double* p; double* ap;
    const double* inv;
    double* partials; size_t pstride;
    int n, ntiles;
    unsigned* bar;        // [0] arrivals, [1] generation, [2] error
};

#define KB_MEGA_SPIN (1u << 28)

// grid barrier among the consumer threads of all CTAs; `last_fn` runs in the CTA that arrives last
template <class F>
__device__ __forceinline__ void kb_grid_barrier(unsigned* bar, KbBulkSmem& S, F&& last_fn) {
    __shared__ unsigned s_gen;
    const int tid = threadIdx.x;
    kb_bar_consumers();
    if (tid == 0) {
        __threadfence();
        s_gen = *reinterpret_cast<volatile unsigned*>(bar + 1);
        const unsigned t = atomicAdd(bar, 1u);
        S.sflag = (t == gridDim.x - 1u);
    }
    kb_bar_consumers();
    if (S.sflag) {
        __threadfence();
        last_fn();
        kb_bar_consumers();
        if (tid == 0) { bar[0] = 0u; __threadfence(); atomicExch(bar + 1, s_gen + 1u); }
    } else if (tid == 0) {
        unsigned spins = 0;
        while (*reinterpret_cast<volatile unsigned*>(bar + 1) == s_gen) {
            if (++spins > KB_MEGA_SPIN) { atomicExch(bar + 2, 1u); break; }
        }
        __threadfence();
    }
    kb_bar_consumers();
}

template <class Fin>   // Fin: PcgApFin / PcgUpdateFin (device functors taking the summed values)
__device__ __forceinline__ void kb_mega_epilogue(const KbPcgMegaArgs& m, KbBulkSmem& S, int nred, Fin fin) {
    double* ssum = &S.d[0][0];
    for (int d = 0; d < nred; ++d) {
        double v = kb_level2_c(m.partials + (size_t)d * m.pstride, m.ntiles, S.red);
        if (threadIdx.x == 0) ssum[d] = v;
    }
    kb_bar_consumers();
    if (threadIdx.x == 0) fin(ssum);
}

template <class ApFin, class UpdFin, class UpdOp, class XpayOp>
__global__ void __launch_bounds__(KB_BULK_THREADS, 2) kb_pcg_persistent(KbSpmvArgs a, KbChunkTable tb, KbPcgMegaArgs m) {
    extern __shared__ __align__(128) unsigned char kb_smem_raw[];
    KbBulkSmem& S = *reinterpret_cast<KbBulkSmem*>(kb_smem_raw);
    __shared__ int s_stop;
    const int tid = threadIdx.x;
    kb_bulk_init_barriers(S);
    if (tid == 0) s_stop = (m.ctl->done != 0);
    __syncthreads();
    int it = 0;
    const unsigned long long pol = kb_policy_evict_first();
    while (true) {
        __syncthreads();                       // iteration boundary for producer and consumers; s_stop is valid
        if (s_stop) break;
        if (tid >= KB_THREADS) {
            if (tid == KB_THREADS) kb_bulk_produce(a, tb, S, it, pol);
            continue;                          // wait for the consumers at the next boundary
        }
        // ---- phase 1: ap = A p, partials of p.Ap
        kb_bulk_consume<true, false, false, false, false, false>(a, S, it, nullptr);   // coherent gathers: p changes every iteration
        kb_grid_barrier(m.bar, S, [&]() { kb_mega_epilogue(m, S, 1, ApFin{m.ctl}); });
        int stop = *reinterpret_cast<volatile int*>(&m.ctl->done) | (int)*reinterpret_cast<volatile unsigned*>(m.bar + 2);
        if (!stop) {
            // ---- phase 2: x += alpha p ; r -= alpha ap ; z = D^-1 r ; partials of r.z and the norm
            UpdOp op; op.n = m.n; op.x = m.x; op.p = m.p; op.r = m.r; op.ap = m.ap; op.inv = m.inv; op.z = m.z; op.ctl = m.ctl;
            for (int tile = blockIdx.x; tile < m.ntiles; tile += gridDim.x) {
                const long long i = (long long)tile * KB_TILE + 2 * tid;
                double red[2] = {0.0, 0.0}, out[2];
                if (i + 1 < m.n) op.pair(i, true, red);
                else if (i < m.n) op.pair(i, false, red);
                else { red[0] = 0.0 + 0.0; red[1] = 0.0 + 0.0; }
                kb_block_reduce_c<2>(red, S.red, out);
                if (tid == 0) { m.partials[tile] = out[0]; m.partials[m.pstride + tile] = out[1]; }
            }
            kb_grid_barrier(m.bar, S, [&]() { kb_mega_epilogue(m, S, 2, UpdFin{m.ctl}); });
            stop = *reinterpret_cast<volatile int*>(&m.ctl->done) | (int)*reinterpret_cast<volatile unsigned*>(m.bar + 2);
            if (!stop) {
                // ---- phase 3: p = z + beta p
                XpayOp xp; xp.n = m.n; xp.z = m.z; xp.p = m.p; xp.ctl = m.ctl;
                for (int tile = blockIdx.x; tile < m.ntiles; tile += gridDim.x) {
                    const long long i = (long long)tile * KB_TILE + 2 * tid;
                    if (i + 1 < m.n) xp.pair(i, true, nullptr);
                    else if (i < m.n) xp.pair(i, false, nullptr);
                }
                kb_grid_barrier(m.bar, S, [&]() {});
            }
        }
        if (tid == 0) s_stop = stop;
    }
}
