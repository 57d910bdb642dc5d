// kb_epilogue.cuh — glue between fused reductions and the scalar recurrences of the solvers.
//
// Single GPU: the last block of the reducing kernel runs the scalar epilogue (alpha, beta,
// Convergence::check ...) inline, so no extra launch and no host round trip.
// Row-block shards: the last block stores its local sums to `slots`; they are all-gathered over
// NVLink, summed in rank order (deterministic) and the same epilogue runs in a 1-thread kernel.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "kb_objects.h"
#include "kb_spmv.cuh"
#include "kb_spmv_bulk.cuh"
#include "kb_spmv_xtile.cuh"
#include "kb_p2p.cuh"

struct KbSpmvArgs;
bool kb_halo_fill_args(kb_csr_s* A, KbSpmvArgs* a);   // peer path: point the SpMV at the IPC mailbox (false: NCCL path)
const KbP2PDev* kb_p2p_dev(kb_ctx_s* c);        // host copy
const KbP2PDev* kb_p2p_dev_ptr(kb_ctx_s* c);    // device-resident copy (kernel argument)
const KbHaloDev* kb_halo_fused_dev(kb_csr_s* A); // non-null: the operand's producer may push the halo itself (kb_halo_push_tile)

// What the last CTA of a reducing kernel does with the local canonical sums (called by ALL its threads):
//   single GPU            : thread 0 runs the scalar epilogue `fin`
//   shard, peer path (p2p): NVLink all-reduce INSIDE this kernel, then `fin` — compute + collective in one launch
//   shard, NCCL path      : store the local sums to `slots`; all-gather + epilogue follow as separate launches
// Fin may define `__device__ void pre(double* ssum) const`: thread 0 of the last CTA calls it before the
// all-reduce to append sums produced by an earlier kernel (single-reduction PCG), so they ride the same collective.
template <class F, class = void> struct kb_has_pre : std::false_type {};
template <class F> struct kb_has_pre<F, std::void_t<decltype(&F::pre)>> : std::true_type {};

template <class Fin>
struct KbFinish {
    Fin fin;
    double* slots = nullptr;
    int nred = 0;
    const KbP2PDev* p2p = nullptr;
    int recv_only = 0;       // pipelined PCG: ssum[1..nred) were SENT by the previous kernel (kb_p2p_allreduce_send); only receive here
    template <int BAR>
    __device__ void coop(double* ssum) const {
        if constexpr (kb_has_pre<Fin>::value) {
            if (threadIdx.x == 0) fin.pre(ssum);
            kb_sync<BAR>();
        }
        if (p2p) {
            if (recv_only) kb_p2p_allreduce_recv<BAR>(*p2p, ssum + 1, nred - 1);
            else kb_p2p_allreduce_block<BAR>(*p2p, ssum, nred);
            if (threadIdx.x == 0) fin(ssum);
        } else if (threadIdx.x == 0) {
            if (slots) { for (int r = 0; r < nred; ++r) slots[r] = ssum[r]; }
            else fin(ssum);
        }
    }
};
template <class Fin>
static KbFinish<Fin> kb_make_fin(kb_ctx_s* c, Fin fin, bool dist, double* slots, int nred) {
    KbFinish<Fin> f;
    f.fin = fin; f.nred = nred;
    if (dist && c->size > 1) {
        f.p2p = kb_p2p_dev_ptr(c);
        f.slots = f.p2p ? nullptr : slots;
    }
    return f;
}

template <class Fin, bool WD, bool YD>
struct KbSpmvEpi {
    static constexpr bool WDOT = WD, YDOT = YD;
    KbCtl* ctl;          // may be null (never skip)
    int skip_mask = 0;   // bit 0: also skip when ctl->early (BiCGStab's 2nd SpMV); bit 1: when ctl->cycle_break (GMRES)
    KbFinish<Fin> fin;
    __device__ bool skip() const {
        return ctl != nullptr && (ctl->done != 0 || ((skip_mask & 1) && ctl->early != 0) || ((skip_mask & 2) && ctl->cycle_break != 0));
    }
    template <int BAR>
    __device__ void finish_block(double* ssum) const { fin.template coop<BAR>(ssum); }
};

template <class Fin>
__global__ void kb_fin_kernel(Fin fin, KbCtl* ctl, const double* sums, int skip_early) {
    if (ctl->done || (skip_early && ctl->early)) return;
    fin(sums);
}

template <class Fin>
static int kb_finish_dist(kb_ctx_s* c, Fin fin, KbCtl* ctl, double* slots, int nred, bool skip_early = false) {
    if (kb_p2p_dev_ptr(c)) return KB_OK;   // peer path: all-reduce + epilogue already ran inside the reducing kernel
    KB_TRY(kb_allreduce_slots(c, slots, nred));
    KbLaunch L(c, KB_K_SMALL);
    kb_fin_kernel<Fin><<<1, 1, 0, c->stream>>>(fin, ctl, slots, skip_early ? 1 : 0);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

template <class Epi, bool RESID, int CFG>
static int kb_launch_spmv_xtile(kb_csr_s* A, const KbSpmvArgs& a, Epi epi, int count, bool pdl) {
    kb_ctx_s* c = A->ctx;
    constexpr int ND = (Epi::WDOT ? 1 : 0) + (Epi::YDOT ? 1 : 0) > 0 ? (Epi::WDOT ? 1 : 0) + (Epi::YDOT ? 1 : 0) : 1;
    using S = KbXtSmem<KbXtCfg<CFG>, ND>;
    static_assert(sizeof(S) <= KB_XT_SMEM_LIMIT(KbXtCfg<CFG>::CTAS), "CTAs per SM");
    auto kfn = A->xt_prod ? kb_spmv_xtile<Epi, RESID, true, CFG> : kb_spmv_xtile<Epi, RESID, false, CFG>;
    if (!c->configured.count((const void*)kfn)) {
        KB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S)));
        c->configured.insert((const void*)kfn);
    }
    KbXtTable tb{A->xt_tile_chunk, A->xt_chunk_row, A->xt_chunk_nz, A->xt_lo, A->xt_len, A->xt_tail, A->xt_lcol, (int)A->ncols_local - 1};
    const int grid = std::min(KbXtCfg<CFG>::CTAS * c->sm_count, count);
    KbLaunch L(c, KB_K_SPMV);
    KB_CUDA(kb_launch_ex(pdl && kb_pdl_enabled(), kfn, dim3(grid), dim3(KB_BULK_THREADS), sizeof(S), c->stream, a, tb, epi));
    return KB_OK;
}

// one launch over a set of canonical tiles (all tiles when list == nullptr); dispatches on the kernel kind
template <class Epi, bool RESID, bool GH>
static int kb_launch_spmv_tiles(kb_csr_s* A, KbSpmvArgs a, Epi epi, const int* list, int count, int finalize, bool pdl = false) {
    if (count <= 0) return KB_OK;
    kb_ctx_s* c = A->ctx;
    a.tile_list = list; a.tile0 = 0; a.ntiles_launch = count; a.finalize = finalize;
    if constexpr (!GH) {
        // x staged in shared memory (operators whose chunks fit, 16-byte aligned operand; no ghost columns)
        if (A->kind == 2 && A->xt && (reinterpret_cast<uintptr_t>(a.x) & 15u) == 0) {
            if (A->xt == 1) return kb_launch_spmv_xtile<Epi, RESID, 0>(A, a, epi, count, pdl);
            if (A->xt == 2) return kb_launch_spmv_xtile<Epi, RESID, 1>(A, a, epi, count, pdl);
            return kb_launch_spmv_xtile<Epi, RESID, 2>(A, a, epi, count, pdl);
        }
    }
    if (A->kind == 2) {
        auto kfn = A->prod ? kb_spmv_bulk<Epi, RESID, GH, true> : kb_spmv_bulk<Epi, RESID, GH, false>;
        if (!c->configured.count((const void*)kfn)) {
            KB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbBulkSmem)));
            c->configured.insert((const void*)kfn);
        }
        KbChunkTable tb{A->tile_chunk, A->chunk_row, A->chunk_nz};
        static const int per_sm = getenv("KB_BULK_CTAS_PER_SM") ? atoi(getenv("KB_BULK_CTAS_PER_SM")) : 2;
        const int grid = std::min(per_sm * c->sm_count, count);
        KbLaunch L(c, KB_K_SPMV);
        KB_CUDA(kb_launch_ex(pdl && kb_pdl_enabled(), kfn, dim3(grid), dim3(KB_BULK_THREADS), sizeof(KbBulkSmem), c->stream, a, tb, epi));
        return KB_OK;
    }
    KbLaunch L(c, KB_K_SPMV);
    if (A->kind == 0) KB_CUDA(kb_launch_ex(pdl && kb_pdl_enabled(), kb_spmv_stream<Epi, RESID, GH>, dim3(count), dim3(KB_THREADS), 0, c->stream, a, epi));
    else if (A->vec == 8) kb_spmv_vector<Epi, RESID, 8, GH><<<count, KB_THREADS, 0, c->stream>>>(a, epi);
    else if (A->vec == 16) kb_spmv_vector<Epi, RESID, 16, GH><<<count, KB_THREADS, 0, c->stream>>>(a, epi);
    else kb_spmv_vector<Epi, RESID, 32, GH><<<count, KB_THREADS, 0, c->stream>>>(a, epi);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// y = A x | y = b - A x with fused dots.  On a shard (halo_x != nullptr: the operand with its ghost tail) the
// halo transfer is started first, the tiles without ghost columns are multiplied while it is in flight, and
// the boundary tiles follow once the ghosts have landed; the dot partials of both launches meet in the last
// CTA of the second one, so the reduction tree is unchanged.
template <class Epi, bool RESID>
static int kb_launch_spmv(kb_csr_s* A, const double* x, double* y, const double* b, const double* w, double* partials,
                          size_t pstride, Epi epi, double* halo_x = nullptr, bool halo_prepushed = false, bool pdl = false) {
    kb_ctx_s* c = A->ctx;
    const bool dist = halo_x != nullptr && A->dist && c->size > 1;
    if (A->n == 0 && !dist) return KB_OK;
    KbSpmvArgs a{};
    a.lazy_from = -1;
    a.row_ptr = A->row_ptr; a.col = A->col; a.vals = A->vals; a.x = x; a.y = y; a.b = b; a.w = w;
    a.n = (int)A->n; a.ntiles_total = A->ntiles;
    a.partials = partials; a.pstride = pstride; a.ticket = c->ticket;
    if (dist && !halo_prepushed) KB_TRY(kb_halo_begin(A, halo_x));     // prepushed: the kernel that produced x pushed its boundary entries itself
    static const int split_env = getenv("KB_HALO_SPLIT") ? atoi(getenv("KB_HALO_SPLIT")) : -1;
    // Interior/boundary split hides the NVLink latency behind the interior rows but costs one more launch.
    // Measured on B200 (256^3, 2 and 8 GPUs) the single launch whose CTAs wait on the flags themselves is
    // faster (8 GPUs: 8089 vs 7627 it/s), so the split is opt-in (KB_HALO_SPLIT=1).
    const bool split = dist && A->n_interior > 0 && A->n_boundary > 0 && split_env > 0;
    // peer path: the kernel that touches ghost columns waits for the neighbours' flags itself and reads the
    // ghosts from the mailbox; NCCL path: ghosts were received into the tail of x by kb_halo_begin
    const bool gh = dist && kb_halo_fill_args(A, &a);
    if (split) {
        KB_TRY((kb_launch_spmv_tiles<Epi, RESID, false>(A, a, epi, A->tiles_interior, A->n_interior, 0)));
        if (gh) return kb_launch_spmv_tiles<Epi, RESID, true>(A, a, epi, A->tiles_boundary, A->n_boundary, 1);
        return kb_launch_spmv_tiles<Epi, RESID, false>(A, a, epi, A->tiles_boundary, A->n_boundary, 1);
    }
    // Measured on B200 (256^3, 2 GPUs): 3272 it/s against 3295 it/s with every CTA waiting at its start - the fused push of the
    // kernel that produced x (sending tiles first) already lands before this launch starts, and the tile indirection costs more
    // than the hidden wait.  Opt-in (KB_HALO_LAZY=1).
    static const bool lazy_env = getenv("KB_HALO_LAZY") && atoi(getenv("KB_HALO_LAZY")) == 1;
    if (gh && lazy_env && A->kind == 2 && A->tiles_order && A->n_boundary > 0 && A->n_interior > 0) {
        // one launch, interior tiles first: only the last tiles of a persistent CTA wait for the neighbours' flags
        a.lazy_from = A->n_interior;
        return kb_launch_spmv_tiles<Epi, RESID, true>(A, a, epi, A->tiles_order, A->ntiles, 1, pdl);
    }
    if (gh) return kb_launch_spmv_tiles<Epi, RESID, true>(A, a, epi, nullptr, A->ntiles, 1, pdl);
    return kb_launch_spmv_tiles<Epi, RESID, false>(A, a, epi, nullptr, A->ntiles, 1, pdl);
}
