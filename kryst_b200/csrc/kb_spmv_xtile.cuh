// kb_spmv_xtile.cuh — bulk-async SpMV with the x operand staged in shared memory (north-star kernel (1):
// "shared-memory/TMA staging of x tiles" for MatVec::matvec, src/core/traits.rs:4-7, src/matrix/sparse.rs:56-67).
//
// kb_spmv_bulk gathers x per non-zero through L1/L2; with 27 entries per row that gather is the limiter (ncu:
// long-scoreboard 12 per issue, 43 % of the DRAM peak).  Here the gather never leaves the SM: at upload every chunk
// of rows gets (a) the <= 16 aligned column intervals that cover all columns it references (device-built: block
// radix sort of the chunk's column ids, gaps <= 8 columns merged) and (b) a second column array of 16-bit ids
// local to the concatenation of those intervals (10 instead of 12 bytes per stored entry).  The producer warp
// stages, per chunk and on one mbarrier, the values, the 16-bit ids, the row_ptr slice (lane 0) and one
// `cp.async.bulk` per x interval (lane k copies interval k); the consumers read x from shared memory only.
// Per-row operation sequence (product rounded, ascending adds) is the one of kb_spmv_bulk and the oracle: same bits.
// Operators whose chunks do not fit (more than 16 intervals, or more x than a stage holds), shards with ghost
// columns and operands that are not 16-byte aligned keep kb_spmv_bulk.  Default for long-row operators (more than 12
// entries per row on average: C3 2 333 -> 3 695 it/s, DRAM traffic per SpMV 711 -> 603 MB); 7-point operators keep
// kb_spmv_bulk, whose L1-served gather already runs at 0.92 of the HBM peak (staged x measured 0.38 vs 0.29 ms on C4).
#pragma once
#include <cub/block/block_radix_sort.cuh>
#include "kb_spmv_bulk.cuh"

#define KB_XT_KMAX 16          // x intervals per chunk
#ifndef KB_XT_DEFAULT_MODE
#define KB_XT_DEFAULT_MODE 2   // KB_SPMV_XTILE when the environment does not say (0 off, 1 whenever it fits, 2 long rows only)
#endif
#ifndef KB_XT_DEFAULT_CFG
#define KB_XT_DEFAULT_CFG 1
#endif
#define KB_XT_GAP 4            // merge intervals whose gap is <= 4 aligned pairs (8 columns)

template <int CFG> struct KbXtCfg;
// Stage geometries (two CTAs per SM, two stages each).  Measured on B200 with the 27-point 128^3 operator (C3), SpMV
// with two fused dots: kb_spmv_bulk 0.175 ms; geometry 1 thread-per-row 0.103 ms, product phase 0.157 ms; geometry 0
// 0.128 / 0.132 ms; 3 stages of 2048 entries 0.150 ms; one CTA per SM with 5 stages 0.19-0.20 ms (the consumers, not the
// bytes in flight, are the limiter).  Geometry 1 is tried first, geometry 0 (more room for x) when a chunk does not fit.
template <> struct KbXtCfg<0> { static constexpr int CAP = 3072, XCAP = 2048, STAGES = 2, MAXROWS = 512, CTAS = 2; };
template <> struct KbXtCfg<1> { static constexpr int CAP = 3584, XCAP = 1280, STAGES = 2, MAXROWS = 256, CTAS = 2; };
// short rows (7-point), opt-in (KB_SPMV_XTILE=1 KB_XT_CFG=2): half a tile per chunk, three CTAs per SM.  Measured on C4 (256^3):
// 0.311 ms per SpMV against 0.301 ms for kb_spmv_bulk (a whole tile per chunk with a 2056-entry x window, two CTAs per SM:
// 0.344 ms; 256-row chunks in geometry 0: 0.384 ms) - staging x costs more shared-memory fill than the L1-served gather
// of seven entries per row saves, so short rows keep kb_spmv_bulk.
template <> struct KbXtCfg<2> { static constexpr int CAP = 1792, XCAP = 1296, STAGES = 2, MAXROWS = 256, CTAS = 3; };
#define KB_XT_NCFG 3
#define KB_XT_SMEM_LIMIT(ctas) ((233472 / (ctas)) - 1024)     // 228 KB per SM, 1 KB reserved per CTA

template <class C>
struct KbXtStage {
    double vals[C::CAP + 16];
    double xs[C::XCAP + 8];
    unsigned short lcol[C::CAP + 16];
    int rp[C::MAXROWS + 8];
    int hdr[8];          // written by the producer: {ra, rb, b0, r_al, tile, last_chunk_of_tile, window, -}
};
template <class C, int ND>
struct KbXtSmem {
    KbXtStage<C> st[C::STAGES];
    double d[ND][KB_TILE];       // dot terms of the current tile (ND = fused dots, at least 1: the epilogue's scratch)
    double red[2 * 8];
    unsigned long long full[C::STAGES];
    unsigned long long empty[C::STAGES];
    int sflag;
};

struct KbXtTable {
    const int* __restrict__ tile_chunk;   // [ntiles+1]
    const int* __restrict__ chunk_row;    // [nchunks+1]
    const int* __restrict__ chunk_nz;     // [nchunks+1]
    const int* __restrict__ lo;           // [nchunks*KMAX] first column of interval k (even)
    const int* __restrict__ len;          // [nchunks*KMAX] its length in columns (even; 0: unused)
    const int* __restrict__ tail;         // [nchunks] shared-memory slot that receives x[xlast] (odd ncols), or -1
    const unsigned short* __restrict__ lcol;   // [nnz+16] chunk-local column ids
    int xlast;                            // ncols - 1
};

// bounded mbarrier wait: a byte-count mismatch must fail the launch, not hang the GPU
__device__ __forceinline__ void kb_xt_wait(unsigned long long* bar, unsigned parity) {
    unsigned spins = 0;
    while (!kb_mbar_try_wait(bar, parity)) { if (++spins > (1u << 24)) __trap(); }
}
__device__ __forceinline__ void kb_bulk_g2s_plain(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(kb_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(kb_smem_u32(bar))
                 : "memory");
}

// producer: the whole warp 8.  Lane 0 owns the ring (waits, header, expect_tx, matrix copies); lane k copies x interval k.
template <class C, class SM>
__device__ __forceinline__ void kb_xt_produce(const KbSpmvArgs& a, const KbXtTable& tb, SM& S, unsigned long long pol) {
    const int lane = threadIdx.x & 31;
    const int ntl = a.ntiles_launch;
    int it = 0;
    for (int ti = blockIdx.x; ti < ntl; ti += gridDim.x) {
        const int tile = a.tile_list ? a.tile_list[ti] : (a.tile0 + ti);
        const int c0 = tb.tile_chunk[tile], c1 = tb.tile_chunk[tile + 1];
        for (int c = c0; c < c1; ++c, ++it) {
            const int s = it % C::STAGES;
            const unsigned ph = (unsigned)(it / C::STAGES) & 1u;
            int mylo = 0, mylen = 0;
            if (lane < KB_XT_KMAX) { mylo = tb.lo[(size_t)c * KB_XT_KMAX + lane]; mylen = tb.len[(size_t)c * KB_XT_KMAX + lane]; }
            int incl = mylen;
#pragma unroll
            for (int off = 1; off < KB_XT_KMAX; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            const int xoff = incl - mylen;
            const int xtot = __shfl_sync(0xffffffffu, incl, KB_XT_KMAX - 1);
            KbXtStage<C>& st = S.st[s];
            const int ra = tb.chunk_row[c], rb = tb.chunk_row[c + 1];
            const int nz0 = tb.chunk_nz[c], nz1 = tb.chunk_nz[c + 1];
            const int b0 = nz0 & ~7, b1 = (nz1 + 7) & ~7;        // 16-B aligned window for the f64 and the u16 array
            const int r_al = ra & ~3;                            // row_ptr slice [r_al, rb] padded to 16 B
            const int nrp = ((rb + 1 - r_al) + 3) & ~3;
            if (lane == 0) {
                const int tl = tb.tail[c];
                const unsigned bytes = (unsigned)(b1 - b0) * 10u + (unsigned)nrp * 4u + (unsigned)xtot * 8u;
                kb_xt_wait(&S.empty[s], ph ^ 1u);
                st.hdr[0] = ra; st.hdr[1] = rb; st.hdr[2] = b0; st.hdr[3] = r_al; st.hdr[4] = tile; st.hdr[5] = (c + 1 == c1); st.hdr[6] = b1 - b0;
                if (tl >= 0) st.xs[tl] = __ldg(a.x + tb.xlast);   // odd ncols: the last column cannot ride a 16-byte copy
                kb_mbar_expect_tx(&S.full[s], bytes);             // release: header and tail are visible to the waiters
            }
            __syncwarp();
            if (lane == 0) {
                if (b1 > b0) {
                    kb_bulk_g2s(st.vals, a.vals + b0, (unsigned)(b1 - b0) * 8u, &S.full[s], pol);
                    kb_bulk_g2s(st.lcol, tb.lcol + b0, (unsigned)(b1 - b0) * 2u, &S.full[s], pol);
                }
                kb_bulk_g2s(st.rp, a.row_ptr + r_al, (unsigned)nrp * 4u, &S.full[s], pol);
            }
            if (mylen > 0) kb_bulk_g2s_plain(st.xs + xoff, a.x + mylo, (unsigned)mylen * 8u, &S.full[s]);
        }
    }
}

template <class C, bool WD, bool YD, bool RESID, bool PROD, class SM>
__device__ __forceinline__ void kb_xt_consume(const KbSpmvArgs& a, SM& S) {
    constexpr int NDOT = (WD ? 1 : 0) + (YD ? 1 : 0);
    constexpr int ND = NDOT > 0 ? NDOT : 1;
    constexpr int YS = WD ? 1 : 0;                      // slot of <y,y>
    const int tid = threadIdx.x;
    const int ntl = a.ntiles_launch;
    int it = 0;
    for (int ti = blockIdx.x; ti < ntl; ti += gridDim.x) {
        int tile = 0, last = 0;
        do {
            const int s = it % C::STAGES;
            const unsigned ph = (unsigned)(it / C::STAGES) & 1u;
            ++it;
            kb_xt_wait(&S.full[s], ph);
            KbXtStage<C>& st = S.st[s];
            const int ra = st.hdr[0], rb = st.hdr[1], b0 = st.hdr[2], r_al = st.hdr[3];
            tile = st.hdr[4]; last = st.hdr[5];
            const int r0 = tile * KB_TILE;
            // rows ra + tid and ra + tid + 256 of this chunk
            const int rA = ra + tid, rB = ra + tid + KB_THREADS;
            const bool hA = rA < rb, hB = (C::MAXROWS > KB_THREADS) && rB < rb;
            int qa0 = 0, qa1 = 0, qb0 = 0, qb1 = 0;
            if (hA) { qa0 = st.rp[rA - r_al] - b0; qa1 = st.rp[rA + 1 - r_al] - b0; }
            if (hB) { qb0 = st.rp[rB - r_al] - b0; qb1 = st.rp[rB + 1 - r_al] - b0; }
            double sA = 0.0, sB = 0.0;
            if (PROD) {
                // all consumers turn the staged values into products in place, then a thread per row adds them in stored order
                const int nwin = st.hdr[6];
                for (int q0 = tid; q0 < nwin; q0 += 4 * KB_THREADS) {
                    double xv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int q = q0 + u * KB_THREADS; xv[u] = q < nwin ? st.xs[st.lcol[q]] : 0.0; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int q = q0 + u * KB_THREADS; if (q < nwin) st.vals[q] = st.vals[q] * xv[u]; }
                }
                kb_bar_consumers();
#pragma unroll 4
                for (int q = qa0; q < qa1; ++q) sA = sA + st.vals[q];
#pragma unroll 4
                for (int q = qb0; q < qb1; ++q) sB = sB + st.vals[q];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes before the next bulk refill
                qa0 = qa1; qb0 = qb1;
            }
            while (qa0 < qa1 || qb0 < qb1) {
                double pa[8], pb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    pa[u] = 0.0; pb[u] = 0.0;
                    if (qa0 + u < qa1) pa[u] = st.xs[st.lcol[qa0 + u]];
                    if (qb0 + u < qb1) pb[u] = st.xs[st.lcol[qb0 + u]];
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

template <class Epi, bool RESID, bool PROD, int CFG>
__global__ void __launch_bounds__(KB_BULK_THREADS, KbXtCfg<CFG>::CTAS) kb_spmv_xtile(KbSpmvArgs a, KbXtTable tb, Epi epi) {
    using C = KbXtCfg<CFG>;
    kb_pdl_wait();
    kb_pdl_launch_dependents();
    if (epi.skip()) return;
    constexpr bool WD = Epi::WDOT, YD = Epi::YDOT;      // fused <w,y> and/or <y,y>
    constexpr int NDOT = (WD ? 1 : 0) + (YD ? 1 : 0);
    constexpr int ND = NDOT > 0 ? NDOT : 1;
    extern __shared__ __align__(128) unsigned char kb_smem_raw[];
    using SM = KbXtSmem<C, ND>;
    SM& S = *reinterpret_cast<SM*>(kb_smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < C::STAGES; ++s) { kb_mbar_init(&S.full[s], 1); kb_mbar_init(&S.empty[s], KB_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= KB_THREADS) {
        kb_xt_produce<C, SM>(a, tb, S, kb_policy_evict_first());
        return;
    }
    kb_xt_consume<C, WD, YD, RESID, PROD, SM>(a, S);
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

// ---- upload-time construction (device) -----------------------------------------------------------
// Chunks: runs of rows inside one canonical tile with <= cap entries and <= maxrows rows, rows spread evenly over
// the chunks a tile needs.  pass 0 counts, pass 1 fills (same walk).
static __global__ void kb_xt_chunk_build(const int* __restrict__ rp, int n, int ntiles, int cap, int maxrows, int* __restrict__ tile_chunk,
                                         int* __restrict__ chunk_row, int* __restrict__ chunk_nz, int fill) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntiles) return;
    const int r0 = tile * KB_TILE, r1 = min(n, r0 + KB_TILE);
    const int tnz = rp[r1] - rp[r0];
    const int k = max(1, (tnz + cap - 1) / cap);
    const int rt = max(1, min(maxrows, (r1 - r0 + k - 1) / k));
    const int c = fill ? tile_chunk[tile] : 0;
    int start = r0, base = rp[r0];
    if (fill) { chunk_row[c] = r0; chunk_nz[c] = base; }
    int cnt = 1;
    for (int r = r0; r < r1; ++r) {
        const int e = rp[r + 1];
        if (r > start && (e - base > cap || r - start >= rt)) {      // row r opens a new chunk
            start = r; base = rp[r];
            if (fill) { chunk_row[c + cnt] = r; chunk_nz[c + cnt] = base; }
            ++cnt;
        }
    }
    if (!fill) tile_chunk[tile] = cnt;
    else if (tile == ntiles - 1) { chunk_row[c + cnt] = r1; chunk_nz[c + cnt] = rp[r1]; }
}

// One CTA per chunk: sort the chunk's column ids, cut them into intervals of aligned pairs, number the columns.
template <int CAP>
static __global__ void __launch_bounds__(KB_THREADS) kb_xt_build(const int* __restrict__ col, const int* __restrict__ chunk_nz, int ncols, int xcap,
                                                                 int* __restrict__ lo, int* __restrict__ len, int* __restrict__ tail,
                                                                 unsigned short* __restrict__ lcol, int* fail) {
    constexpr int IPT = CAP / KB_THREADS;
    using Sort = cub::BlockRadixSort<int, KB_THREADS, IPT>;
    __shared__ typename Sort::TempStorage tmp;
    __shared__ int s_key[CAP];
    __shared__ int s_start[KB_XT_KMAX], s_lo[KB_XT_KMAX], s_len[KB_XT_KMAX], s_off[KB_XT_KMAX];
    __shared__ int s_cnt, s_tail, s_ok;
    const int tid = threadIdx.x;
    const int c = blockIdx.x;
    const int nz0 = chunk_nz[c], nz1 = chunk_nz[c + 1], nv = nz1 - nz0;
    int keys[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) { const int q = tid * IPT + i; keys[i] = q < nv ? col[nz0 + q] : 0x7fffffff; }
    Sort(tmp).Sort(keys);
#pragma unroll
    for (int i = 0; i < IPT; ++i) s_key[tid * IPT + i] = keys[i];
    if (tid == 0) s_cnt = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int q = tid * IPT + i;
        if (q < nv) {
            const bool st = q == 0 || ((s_key[q] >> 1) - (s_key[q - 1] >> 1) > KB_XT_GAP);
            if (st) { const int k = atomicAdd(&s_cnt, 1); if (k < KB_XT_KMAX) s_start[k] = q; }
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int cnt = s_cnt;
        bool ok = cnt <= KB_XT_KMAX;
        int tl = -1, off = 0;
        if (ok) {
            for (int i = 1; i < cnt; ++i) {                  // the starts arrive unordered
                const int v = s_start[i];
                int j = i - 1;
                while (j >= 0 && s_start[j] > v) { s_start[j + 1] = s_start[j]; --j; }
                s_start[j + 1] = v;
            }
            const int ne = ncols & ~1;
            for (int k = 0; k < cnt; ++k) {
                const int q0 = s_start[k], q1 = k + 1 < cnt ? s_start[k + 1] : nv;
                int l = (s_key[q0] >> 1) << 1, h = ((s_key[q1 - 1] >> 1) + 1) << 1;
                if (h > ne) { h = ne; tl = 0; }             // odd ncols and the chunk references the last column
                if (l > h) l = h;
                s_lo[k] = l; s_len[k] = h - l; s_off[k] = off;
                off += h - l;
            }
            if (tl == 0) { tl = off; off += 1; }
            ok = off <= xcap;
        }
        s_ok = ok ? 1 : 0; s_tail = ok ? tl : -1;
        if (!ok) atomicExch(fail, 1);
        for (int k = 0; k < KB_XT_KMAX; ++k) {
            const bool used = ok && k < cnt;
            lo[(size_t)c * KB_XT_KMAX + k] = used ? s_lo[k] : 0;
            len[(size_t)c * KB_XT_KMAX + k] = used ? s_len[k] : 0;
        }
        tail[c] = ok ? tl : -1;
    }
    __syncthreads();
    if (!s_ok) return;
    const int cnt = s_cnt, tl = s_tail;
    for (int q = nz0 + tid; q < nz1; q += KB_THREADS) {
        const int cc = col[q];
        int lid = 0;
        if (tl >= 0 && cc == ncols - 1) lid = tl;
        else {
            for (int k = 0; k < cnt; ++k) {
                if (cc >= s_lo[k] && cc < s_lo[k] + s_len[k]) { lid = s_off[k] + cc - s_lo[k]; break; }
            }
        }
        lcol[q] = (unsigned short)lid;
    }
}
