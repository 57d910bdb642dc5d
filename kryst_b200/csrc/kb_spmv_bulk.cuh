// kb_spmv_bulk.cuh — the sm_100a SpMV: warp-specialised, bulk-async (TMA engine) staged CSR-stream.
//
// Persistent CTAs (2 per SM).  Warp 8 is the producer: one elected lane walks the CTA's chunk
// sequence and, per chunk, issues three `cp.async.bulk` copies (values f64, column ids i32, the
// tile's row_ptr slice) global -> shared, completion counted on an mbarrier (`complete_tx`), with
// an L2 evict-first policy so the once-read matrix stream does not displace x.  Warps 0-7 are
// consumers: wait on the stage's `full` mbarrier, thread-per-row sums out of shared memory with
// all gathers of a thread's rows issued before the first add (ascending stored order, mul then
// add: bit-identical to the oracle), write y coalesced, arrive on the `empty` mbarrier.
// A chunk is a run of rows inside one canonical 512-row tile with <= KB_BULK_CAP nonzeros (table
// built on the device at upload); the tile's dot terms are re-paired into the canonical lanes
// exactly as in kb_spmv_stream, so results do not depend on the grid or the kernel variant.
// HBM-bound by construction: bytes in flight per SM = stages x CTAs x ~43 KB, independent of
// warp scheduling.  SASS evidence: UBLKCP (bulk copy), SYNCS (mbarrier).
#pragma once
#include "kb_spmv.cuh"

#define KB_BULK_CAP 4096                 // nonzeros per stage
#define KB_BULK_PAD 8
#define KB_BULK_STAGES 2
#define KB_BULK_THREADS (KB_THREADS + 32)

struct KbBulkStage {
    double vals[KB_BULK_CAP + KB_BULK_PAD];
    int cols[KB_BULK_CAP + KB_BULK_PAD];
    int rp[KB_TILE + 8];
    int hdr[8];          // written by the producer: {ra, rb, b0, r_al, tile, last_chunk_of_tile, -, -}
};
struct KbBulkSmem {
    KbBulkStage st[KB_BULK_STAGES];
    double d[2][KB_TILE];
    double red[2 * 8];
    unsigned long long full[KB_BULK_STAGES];
    unsigned long long empty[KB_BULK_STAGES];
    int sflag;
};

__device__ __forceinline__ unsigned kb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kb_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kb_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(kb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool kb_mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(kb_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void kb_mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!kb_mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ unsigned long long kb_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk async copy global -> shared (TMA engine), bytes multiple of 16, both addresses 16-B aligned
__device__ __forceinline__ void kb_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     kb_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(kb_smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void kb_bar_consumers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// canonical block reduce among the 256 consumer threads (named barrier 1)
template <int NRED>
__device__ __forceinline__ void kb_block_reduce_c(double (&v)[NRED], double* sm, double (&out)[NRED]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < NRED; ++r) {
        double t = kb_warp_butterfly(v[r]);
        if (lane == 0) sm[r * 8 + w] = t;
    }
    kb_bar_consumers();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < NRED; ++r) {
            double s = sm[r * 8];
#pragma unroll
            for (int k = 1; k < 8; ++k) s = s + sm[r * 8 + k];
            out[r] = s;
        }
    }
    kb_bar_consumers();
}
__device__ __forceinline__ double kb_level2_c(const double* __restrict__ part, int P, double* sm8) {
    double acc = 0.0;
#pragma unroll 8
    for (int k = threadIdx.x; k < P; k += KB_THREADS) acc = acc + __ldcg(part + k);
    double v[1] = {acc}, out[1] = {0.0};
    kb_block_reduce_c<1>(v, sm8, out);
    return out[0];
}

struct KbChunkTable {
    const int* __restrict__ tile_chunk;   // [ntiles+1] first chunk of each tile
    const int* __restrict__ chunk_row;    // [nchunks+1] first row of each chunk
    const int* __restrict__ chunk_nz;     // [nchunks+1] row_ptr[chunk_row[k]]
};

// ---- producer / consumer halves of the bulk SpMV (shared by kb_spmv_bulk and the persistent PCG kernel) ------
// `it` counts stages used so far by this CTA; it persists across calls so the mbarrier phases keep alternating.
__device__ __forceinline__ void kb_bulk_produce(const KbSpmvArgs& a, const KbChunkTable& tb, KbBulkSmem& S, int& it, unsigned long long pol) {
    const int ntl = a.ntiles_launch;
    for (int ti = blockIdx.x; ti < ntl; ti += gridDim.x) {
        const int tile = a.tile_list ? a.tile_list[ti] : (a.tile0 + ti);
        const int c0 = tb.tile_chunk[tile], c1 = tb.tile_chunk[tile + 1];
        for (int c = c0; c < c1; ++c, ++it) {
            const int s = it % KB_BULK_STAGES;
            const unsigned ph = (unsigned)(it / KB_BULK_STAGES) & 1u;
            const int ra = tb.chunk_row[c], rb = tb.chunk_row[c + 1];
            const int nz0 = tb.chunk_nz[c], nz1 = tb.chunk_nz[c + 1];
            const int b0 = nz0 & ~3, b1 = (nz1 + 3) & ~3;        // 16-B aligned window for both arrays
            const int r_al = ra & ~3;                            // row_ptr slice [r_al, rb] padded to 16 B
            const int nrp = ((rb + 1 - r_al) + 3) & ~3;
            const unsigned bytes = (unsigned)(b1 - b0) * 12u + (unsigned)nrp * 4u;
            kb_mbar_wait(&S.empty[s], ph ^ 1u);
            KbBulkStage& st = S.st[s];
            st.hdr[0] = ra; st.hdr[1] = rb; st.hdr[2] = b0; st.hdr[3] = r_al; st.hdr[4] = tile; st.hdr[5] = (c + 1 == c1); st.hdr[6] = b1 - b0;
            kb_mbar_expect_tx(&S.full[s], bytes);
            if (b1 > b0) {
                kb_bulk_g2s(st.vals, a.vals + b0, (unsigned)(b1 - b0) * 8u, &S.full[s], pol);
                kb_bulk_g2s(st.cols, a.col + b0, (unsigned)(b1 - b0) * 4u, &S.full[s], pol);
            }
            kb_bulk_g2s(st.rp, a.row_ptr + r_al, (unsigned)nrp * 4u, &S.full[s], pol);
        }
    }
}

// PROD (long rows, e.g. 27-point): a chunk holds fewer rows than consumer threads, so the gathers are not done
// per row but per nonzero — all 256 consumers turn the staged values into products in place, then a thread per
// row adds its products in stored order.  Same operation sequence (product rounded, then ascending adds) =>
// same bits as the row-wise path and the oracle.
template <bool WD, bool YD, bool RESID, bool GH, bool PROD, bool NC = true>
__device__ __forceinline__ void kb_bulk_consume(const KbSpmvArgs& a, KbBulkSmem& S, int& it, const double* xg) {
    constexpr int NDOT = (WD ? 1 : 0) + (YD ? 1 : 0);
    constexpr int ND = NDOT > 0 ? NDOT : 1;
    constexpr int YS = WD ? 1 : 0;                      // slot of <y,y>
    const int tid = threadIdx.x;
    const int ntl = a.ntiles_launch;
    bool halo_ready = !GH || a.lazy_from < 0;
    for (int ti = blockIdx.x; ti < ntl; ti += gridDim.x) {
        int tile = 0, last = 0;
        if (GH) {
            if (!halo_ready && ti >= a.lazy_from) {      // CTA-uniform: the first tile with ghost columns of this CTA
                (void)kb_halo_wait<true>(a);
                kb_bar_consumers();
                halo_ready = true;
            }
        }
        do {
            const int s = it % KB_BULK_STAGES;
            const unsigned ph = (unsigned)(it / KB_BULK_STAGES) & 1u;
            ++it;
            kb_mbar_wait(&S.full[s], ph);
            KbBulkStage& st = S.st[s];
            const int ra = st.hdr[0], rb = st.hdr[1], b0 = st.hdr[2], r_al = st.hdr[3];
            tile = st.hdr[4]; last = st.hdr[5];
            const int r0 = tile * KB_TILE;
            // rows ra + tid and ra + tid + 256 of this chunk
            const int rA = ra + tid, rB = ra + tid + KB_THREADS;
            const bool hA = rA < rb, hB = rB < rb;
            int qa0 = 0, qa1 = 0, qb0 = 0, qb1 = 0;
            if (hA) { qa0 = st.rp[rA - r_al] - b0; qa1 = st.rp[rA + 1 - r_al] - b0; }
            if (hB) { qb0 = st.rp[rB - r_al] - b0; qb1 = st.rp[rB + 1 - r_al] - b0; }
            double sA = 0.0, sB = 0.0;
            if (PROD) {
                const int nwin = st.hdr[6];
                for (int q0 = tid; q0 < nwin; q0 += 8 * KB_THREADS) {
                    double xv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int q = q0 + u * KB_THREADS; xv[u] = q < nwin ? kb_xload<GH, NC>(a, xg, st.cols[q]) : 0.0; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int q = q0 + u * KB_THREADS; if (q < nwin) st.vals[q] = st.vals[q] * xv[u]; }
                }
                kb_bar_consumers();
#pragma unroll 4
                for (int q = qa0; q < qa1; ++q) sA = sA + st.vals[q];
#pragma unroll 4
                for (int q = qb0; q < qb1; ++q) sB = sB + st.vals[q];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes before the next bulk refill
                qa0 = qa1; qb0 = qb1;
            }
            // groups of 8 entries per row: gather everything first, then add in stored order
            while (qa0 < qa1 || qb0 < qb1) {
                double pa[8], pb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    pa[u] = 0.0; pb[u] = 0.0;
                    if (qa0 + u < qa1) pa[u] = kb_xload<GH, NC>(a, xg, st.cols[qa0 + u]);
                    if (qb0 + u < qb1) pb[u] = kb_xload<GH, NC>(a, xg, st.cols[qb0 + u]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (qa0 + u < qa1) sA = sA + st.vals[qa0 + u] * pa[u];
                    if (qb0 + u < qb1) sB = sB + st.vals[qb0 + u] * pb[u];
                }
                qa0 += 8; qb0 += 8;
            }
            kb_mbar_arrive(&S.empty[s]);     // stage consumed (all smem reads of this thread are done)
            if (hA) {
                double yv = RESID ? (a.b[rA] - sA) : sA;
                a.y[rA] = yv;
                if constexpr (WD) S.d[0][rA - r0] = a.w[rA] * yv;
                if constexpr (YD) S.d[YS][rA - r0] = yv * yv;
            }
            if (hB) {
                double yv = RESID ? (a.b[rB] - sB) : sB;
                a.y[rB] = yv;
                if constexpr (WD) S.d[0][rB - r0] = a.w[rB] * yv;
                if constexpr (YD) S.d[YS][rB - r0] = yv * yv;
            }
        } while (!last);
        if constexpr (NDOT > 0) {
            // rows of the tile beyond n contribute +0.0 (nobody else writes them)
            const int nr = min(KB_TILE, a.n - tile * KB_TILE);
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                if (tid >= nr) S.d[d][tid] = 0.0;
                if (tid + KB_THREADS >= nr) S.d[d][tid + KB_THREADS] = 0.0;
            }
            kb_bar_consumers();
            double red[ND], out[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) red[d] = S.d[d][2 * tid] + S.d[d][2 * tid + 1];
            kb_block_reduce_c<ND>(red, S.red, out);
            if (tid == 0) {
#pragma unroll
                for (int d = 0; d < ND; ++d) a.partials[(size_t)d * a.pstride + tile] = out[d];
            }
        }
    }
}
__device__ __forceinline__ void kb_bulk_init_barriers(KbBulkSmem& S) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < KB_BULK_STAGES; ++s) { kb_mbar_init(&S.full[s], 1); kb_mbar_init(&S.empty[s], KB_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}

template <class Epi, bool RESID, bool GH = false, bool PROD = false>
__global__ void __launch_bounds__(KB_BULK_THREADS, 2) kb_spmv_bulk(KbSpmvArgs a, KbChunkTable tb, Epi epi) {
    kb_pdl_wait();
    kb_pdl_launch_dependents();
    if (epi.skip()) return;
    const double* xg = nullptr;
    if (GH) xg = a.lazy_from >= 0 ? kb_halo_wait<false>(a) : kb_halo_wait<true>(a);      // ordered before the gathers by the __syncthreads() below; lazy: see kb_bulk_consume
    constexpr bool WD = Epi::WDOT, YD = Epi::YDOT;      // fused <w,y> and/or <y,y>
    constexpr int NDOT = (WD ? 1 : 0) + (YD ? 1 : 0);
    constexpr int ND = NDOT > 0 ? NDOT : 1;
    extern __shared__ __align__(128) unsigned char kb_smem_raw[];
    KbBulkSmem& S = *reinterpret_cast<KbBulkSmem*>(kb_smem_raw);
    const int tid = threadIdx.x;
    kb_bulk_init_barriers(S);
    __syncthreads();
    int it = 0;
    if (tid >= KB_THREADS) {
        // producer warp: one elected lane feeds the shared-memory ring
        if (tid == KB_THREADS) kb_bulk_produce(a, tb, S, it, kb_policy_evict_first());
        return;
    }
    kb_bulk_consume<WD, YD, RESID, GH, PROD>(a, S, it, xg);
    if constexpr (NDOT > 0) {
        if (a.finalize) {
            if (tid == 0) {
                __threadfence();
                unsigned t = atomicAdd(a.ticket, 1u);
                S.sflag = (t == gridDim.x - 1u);
                if (S.sflag) { *a.ticket = 0u; __threadfence(); }
            }
            kb_bar_consumers();
            if (S.sflag) {
                double* ssum = &S.d[0][0];     // tile buffer is free now
#pragma unroll
                for (int d = 0; d < ND; ++d) { double v = kb_level2_c(a.partials + (size_t)d * a.pstride, a.ntiles_total, S.red); if (tid == 0) ssum[d] = v; }
                kb_bar_consumers();
                epi.template finish_block<1>(ssum);
            }
        }
    }
}

// ---- chunk table construction (device, at upload) -------------------------------------------------
// pass 0: count chunks per tile (and flag rows longer than the stage); pass 1: fill rows / nz offsets.
static __global__ void kb_chunk_build(const int* __restrict__ rp, int n, int ntiles, int* __restrict__ tile_chunk,
                               int* __restrict__ chunk_row, int* __restrict__ chunk_nz, int fill, int* too_long) {
    int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntiles) return;
    const int r0 = tile * KB_TILE, r1 = min(n, r0 + KB_TILE);
    int c = fill ? tile_chunk[tile] : 0;
    int start = r0, base = rp[r0];
    if (fill) { chunk_row[c] = r0; chunk_nz[c] = base; }
    int cnt = 1;
    for (int r = r0; r < r1; ++r) {
        const int e = rp[r + 1];
        if (e - rp[r] > KB_BULK_CAP) atomicExch(too_long, 1);
        if (e - base > KB_BULK_CAP && r > start) {          // row r does not fit any more: it opens a new chunk
            start = r; base = rp[r];
            if (fill) { chunk_row[c + cnt] = r; chunk_nz[c + cnt] = base; }
            ++cnt;
        }
    }
    if (!fill) tile_chunk[tile] = cnt;
    else if (tile == ntiles - 1) { chunk_row[c + cnt] = r1; chunk_nz[c + cnt] = rp[r1]; }
}
