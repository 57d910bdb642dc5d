// kb_spmv.cuh — CSR SpMV kernels (K1/K2/K14 of SURVEY §7.2), templated on a fused epilogue.
//
// Replaces: MatVec::matvec dense loop (src/core/wrappers.rs:27-38) and the densifying
// CsrMatrix::spmv (src/matrix/sparse.rs:56-67).  y = A x (or y = b - A x), optionally fused with
// up to two canonical dots  <w, y>  and  <y, y>  (p.Ap for PCG: pcg.rs:150-160; r^.v and t.s/t.t
// for BiCGStab: bicgstab.rs:144-160,208-234; ||b-Ax|| for GMRES: gmres.rs:388-392).
//
// CSR-stream design (HBM-bound, no tensor cores): one block owns one canonical tile of 512 rows.
//   load phase : the tile's contiguous nnz range is streamed with coalesced 128-bit (values) and
//                64-bit (column index) evict-first loads; products vals[k]*x[col[k]] go to shared
//                memory.  x is gathered through L1/L2 (stencil reuse distance << L2).
//   row phase  : one thread per row sums its products from shared memory in ascending stored
//                order (mul, then add: bit-identical to the oracle's row sum), writes y coalesced.
//   dot phase  : per-row dot terms are re-paired through shared memory into the canonical lanes.
// Tiles whose nnz exceed the staging capacity are processed in several row rounds; a single row
// longer than the capacity is accumulated chunk by chunk (same order).
#pragma once
#include "kb_internal.cuh"

#define KB_SPMV_CAP 4096   // products staged per round (32 KB)

struct KbSpmvArgs {
    const int* __restrict__ row_ptr;
    const int* __restrict__ col;
    const double* __restrict__ vals;
    const double* __restrict__ x;     // length ncols_local (owned + ghosts)
    double* __restrict__ y;           // length n
    const double* __restrict__ b;     // RESID: y = b - A x
    const double* __restrict__ w;     // NDOT>=1: <w, y>
    int n;
    int tile0;                        // first tile handled by this launch
    int ntiles_launch;                // number of tiles handled by this launch
    int ntiles_total;                 // total tiles of the vector (level-2 length)
    const int* __restrict__ tile_list;// optional indirection (boundary / interior tile sets)
    int finalize;                     // 1: last block of this launch reduces level 2 and runs the epilogue
    double* partials;
    size_t pstride;
    unsigned* ticket;
    // shard + peer path (GH kernels): ghost values are read straight from the IPC mailbox the neighbours
    // pushed into, after waiting for their sequence-numbered flags (no receive kernel, no ghost copy)
    const double* xg_base;              // ghost_in[2][xg_stride] (nullptr: ghosts live in the tail of x)
    long long xg_stride;
    int n_loc;
    const unsigned long long* hseq;     // pushes done by this rank == number of the exchange being consumed
    const unsigned long long* hflags;   // flags[2][hsize]
    int hsize; unsigned hsrc_mask;
    unsigned* herr;
    int lazy_from;                      // bulk kernel: >= 0 -> positions [lazy_from, ntiles_launch) of tile_list are the tiles with ghost columns;
                                        // a CTA waits for the halo only when it reaches the first of them (interior rows hide the NVLink flight)
};

// Wait until every source rank has published exchange #seq; returns the ghost pointer biased by -n_loc, so
// that xg[c] is the value of local column c >= n_loc.  All threads call it; the caller synchronises after.
template <bool WAIT = true>
__device__ __forceinline__ const double* kb_halo_wait(const KbSpmvArgs& a) {
    const unsigned long long seq = *reinterpret_cast<const volatile unsigned long long*>(a.hseq);
    const size_t par = (size_t)(seq & 1ull);
    const int tid = threadIdx.x;
    if (WAIT && tid < a.hsize && ((a.hsrc_mask >> tid) & 1u)) {
        const volatile unsigned long long* f = a.hflags + par * a.hsize + tid;
        unsigned spins = 0;
        while (*f < seq) { if (++spins > (1u << 25)) { atomicExch(a.herr, 1u); break; } }
        __threadfence_system();
    }
    return a.xg_base + par * a.xg_stride - a.n_loc;
}
// NC: the operand is read-only for the kernel's lifetime -> non-coherent (texture path) loads.  The persistent
// PCG kernel rewrites p between phases, so there the gather must be an ordinary coherent load.
template <bool GH, bool NC = true>
__device__ __forceinline__ double kb_xload(const KbSpmvArgs& a, const double* xg, int c) {
    if (GH) { if (c >= a.n_loc) return __ldcg(xg + c); }
    return NC ? __ldg(a.x + c) : __ldca(a.x + c);
}

// Epi: struct with static constexpr bool WDOT, YDOT (slot order: <w,y> then <y,y>); __device__ bool skip() const; template <int BAR> __device__ void finish_block(double* ssum) const  (all threads of the last CTA)
template <class Epi, bool RESID, bool GH = false>
__global__ void __launch_bounds__(KB_THREADS) kb_spmv_stream(KbSpmvArgs a, Epi epi) {
    kb_pdl_wait();
    kb_pdl_launch_dependents();
    if (epi.skip()) return;
    const double* xg = nullptr;
    if (GH) xg = kb_halo_wait(a);
    constexpr bool WD = Epi::WDOT, YD = Epi::YDOT;      // fused <w,y> and/or <y,y>
    constexpr int NDOT = (WD ? 1 : 0) + (YD ? 1 : 0);
    constexpr int ND = NDOT > 0 ? NDOT : 1;
    constexpr int YS = WD ? 1 : 0;                      // slot of <y,y>
    __shared__ int s_rp[KB_TILE + 1];
    __shared__ double s_prod[KB_SPMV_CAP];
    __shared__ double s_d[ND][KB_TILE];
    __shared__ double s_red[ND * 8];
    __shared__ int sflag;

    const int tid = threadIdx.x;
    const int tile = a.tile_list ? a.tile_list[blockIdx.x] : (a.tile0 + (int)blockIdx.x);
    const int r0 = tile * KB_TILE;
    const int nr = min(KB_TILE, a.n - r0);

    for (int t = tid; t <= nr; t += KB_THREADS) s_rp[t] = a.row_ptr[r0 + t];
    if constexpr (NDOT > 0) {
        for (int t = tid; t < KB_TILE; t += KB_THREADS) {
#pragma unroll
            for (int d = 0; d < ND; ++d) s_d[d][t] = 0.0;
        }
    }
    __syncthreads();

    int rs = 0;
    while (rs < nr) {
        const int base = s_rp[rs];
        int re;
        if (s_rp[nr] - base <= KB_SPMV_CAP) re = nr;
        else {
            // largest re in (rs, nr] with s_rp[re]-base <= CAP  (uniform across the block)
            int lo = rs, hi = nr;
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_rp[mid] - base <= KB_SPMV_CAP) lo = mid; else hi = mid; }
            re = lo;
        }
        if (re == rs) {
            // single row longer than the staging buffer: stream it in chunks, thread 0 keeps the
            // running sum so the addition order stays ascending.
            const int end = s_rp[rs + 1];
            double s = 0.0;
            for (int cb = base; cb < end; cb += KB_SPMV_CAP) {
                const int ce = min(end, cb + KB_SPMV_CAP);
                for (int k = cb + tid; k < ce; k += KB_THREADS) s_prod[k - cb] = __ldcs(a.vals + k) * kb_xload<GH>(a, xg, __ldcs(a.col + k));
                __syncthreads();
                if (tid == 0) for (int q = 0; q < ce - cb; ++q) s = s + s_prod[q];
                __syncthreads();
            }
            if (tid == 0) {
                const int r = r0 + rs;
                double yv = RESID ? (a.b[r] - s) : s;
                a.y[r] = yv;
                if constexpr (WD) s_d[0][rs] = a.w[r] * yv;
                if constexpr (YD) s_d[YS][rs] = yv * yv;
            }
            rs = rs + 1;
            continue;
        }
        const int end = s_rp[re];
        const int k0 = base & ~1;
        for (int k = k0 + 2 * tid; k < end; k += 2 * KB_THREADS) {
            const double2 v = __ldcs(reinterpret_cast<const double2*>(a.vals + k));
            const int2 c = __ldcs(reinterpret_cast<const int2*>(a.col + k));
            const double p0 = v.x * kb_xload<GH>(a, xg, c.x);
            const double p1 = v.y * kb_xload<GH>(a, xg, c.y);
            if (k >= base) s_prod[k - base] = p0;
            if (k + 1 < end) s_prod[k + 1 - base] = p1;
        }
        __syncthreads();
        for (int t = rs + tid; t < re; t += KB_THREADS) {
            double s = 0.0;
            const int qe = s_rp[t + 1] - base;
            for (int q = s_rp[t] - base; q < qe; ++q) s = s + s_prod[q];
            const int r = r0 + t;
            double yv = RESID ? (a.b[r] - s) : s;
            a.y[r] = yv;
            if constexpr (WD) s_d[0][t] = a.w[r] * yv;
            if constexpr (YD) s_d[YS][t] = yv * yv;
        }
        __syncthreads();
        rs = re;
    }

    if constexpr (NDOT > 0) {
        __syncthreads();
        double red[ND], out[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) red[d] = s_d[d][2 * tid] + s_d[d][2 * tid + 1];
        kb_block_reduce<ND>(red, s_red, out);
        if (tid == 0) {
#pragma unroll
            for (int d = 0; d < ND; ++d) a.partials[(size_t)d * a.pstride + tile] = out[d];
        }
        if (a.finalize && kb_arrive_last(a.ticket, gridDim.x, &sflag)) {
            __shared__ double ssum[ND + 4];      // + room for sums appended by Fin::pre
#pragma unroll
            for (int d = 0; d < ND; ++d) { double v = kb_level2(a.partials + (size_t)d * a.pstride, a.ntiles_total, s_red); if (tid == 0) ssum[d] = v; }
            __syncthreads();
            epi.template finish_block<0>(ssum);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Vector-per-row variant for long rows (selected by the row-length histogram at upload):
// a sub-warp of VEC lanes owns one row, strided partial sums + shuffle tree.  The summation
// order differs from the oracle's sequential row sum, so parity for this variant is <= 1e-12
// relative instead of bit-exact.  The dot epilogue is shared with the stream kernel.
// ---------------------------------------------------------------------------------------------
template <class Epi, bool RESID, int VEC, bool GH = false>
__global__ void __launch_bounds__(KB_THREADS) kb_spmv_vector(KbSpmvArgs a, Epi epi) {
    kb_pdl_wait();
    kb_pdl_launch_dependents();
    if (epi.skip()) return;
    const double* xg = nullptr;
    if (GH) { xg = kb_halo_wait(a); __syncthreads(); }
    constexpr bool WD = Epi::WDOT, YD = Epi::YDOT;      // fused <w,y> and/or <y,y>
    constexpr int NDOT = (WD ? 1 : 0) + (YD ? 1 : 0);
    constexpr int ND = NDOT > 0 ? NDOT : 1;
    constexpr int YS = WD ? 1 : 0;                      // slot of <y,y>
    __shared__ double s_d[ND][KB_TILE];
    __shared__ double s_red[ND * 8];
    __shared__ int sflag;
    const int tid = threadIdx.x;
    const int tile = a.tile_list ? a.tile_list[blockIdx.x] : (a.tile0 + (int)blockIdx.x);
    const int r0 = tile * KB_TILE;
    const int nr = min(KB_TILE, a.n - r0);
    if constexpr (NDOT > 0) {
        for (int t = tid; t < KB_TILE; t += KB_THREADS) {
#pragma unroll
            for (int d = 0; d < ND; ++d) s_d[d][t] = 0.0;
        }
        __syncthreads();
    }
    constexpr int ROWS_PER_PASS = KB_THREADS / VEC;
    const int sub = tid / VEC, sl = tid % VEC;
    for (int t = sub; t < nr; t += ROWS_PER_PASS) {
        const int r = r0 + t;
        const int pb = a.row_ptr[r], pe = a.row_ptr[r + 1];
        double s = 0.0;
        for (int k = pb + sl; k < pe; k += VEC) s = s + __ldcs(a.vals + k) * kb_xload<GH>(a, xg, __ldcs(a.col + k));
#pragma unroll
        for (int off = VEC / 2; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off, VEC);
        if (sl == 0) {
            double yv = RESID ? (a.b[r] - s) : s;
            a.y[r] = yv;
            if constexpr (WD) s_d[0][t] = a.w[r] * yv;
            if constexpr (YD) s_d[YS][t] = yv * yv;
        }
    }
    if constexpr (NDOT > 0) {
        __syncthreads();
        double red[ND], out[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) red[d] = s_d[d][2 * tid] + s_d[d][2 * tid + 1];
        kb_block_reduce<ND>(red, s_red, out);
        if (tid == 0) {
#pragma unroll
            for (int d = 0; d < ND; ++d) a.partials[(size_t)d * a.pstride + tile] = out[d];
        }
        if (a.finalize && kb_arrive_last(a.ticket, gridDim.x, &sflag)) {
            __shared__ double ssum[ND + 4];      // + room for sums appended by Fin::pre
#pragma unroll
            for (int d = 0; d < ND; ++d) { double v = kb_level2(a.partials + (size_t)d * a.pstride, a.ntiles_total, s_red); if (tid == 0) ssum[d] = v; }
            __syncthreads();
            epi.template finish_block<0>(ssum);
        }
    }
}

// Plain y = A x epilogue (MatVec::matvec)
struct KbEpiNone {
    static constexpr bool WDOT = false, YDOT = false;
    __device__ bool skip() const { return false; }
    template <int BAR>
    __device__ void finish_block(double*) const {}
};
