// kb_epilogue.cuh — glue between fused reductions and the scalar recurrences of the solvers.
//
// Single GPU: the last block of the reducing kernel runs the scalar epilogue (alpha, beta,
// Convergence::check ...) inline, so no extra launch and no host round trip.
// Row-block shards: the last block stores its local sums to `slots`; they are all-gathered over
// NVLink, summed in rank order (deterministic) and the same epilogue runs in a 1-thread kernel.
#pragma once
#include <algorithm>
#include "kb_objects.h"
#include "kb_spmv.cuh"
#include "kb_spmv_bulk.cuh"

template <class Fin>
struct KbFinish {
    Fin fin;
    double* slots;   // nullptr => run the epilogue inline
    int nred;
    __device__ void operator()(const double* s) const {
        if (slots) { for (int r = 0; r < nred; ++r) slots[r] = s[r]; }
        else fin(s);
    }
};

template <class Fin, bool WD, bool YD>
struct KbSpmvEpi {
    static constexpr bool WDOT = WD, YDOT = YD;
    KbCtl* ctl;          // may be null (never skip)
    int skip_mask = 0;   // bit 0: also skip when ctl->early (BiCGStab's 2nd SpMV); bit 1: when ctl->cycle_break (GMRES)
    KbFinish<Fin> fin;
    __device__ bool skip() const {
        return ctl != nullptr && (ctl->done != 0 || ((skip_mask & 1) && ctl->early != 0) || ((skip_mask & 2) && ctl->cycle_break != 0));
    }
    __device__ void finish(const double* s) const { fin(s); }
};

template <class Fin>
__global__ void kb_fin_kernel(Fin fin, KbCtl* ctl, const double* sums, int skip_early) {
    if (ctl->done || (skip_early && ctl->early)) return;
    fin(sums);
}

template <class Fin>
static int kb_finish_dist(kb_ctx_s* c, Fin fin, KbCtl* ctl, double* slots, int nred, bool skip_early = false) {
    KB_TRY(kb_allreduce_slots(c, slots, nred));
    KbLaunch L(c, KB_K_SMALL);
    kb_fin_kernel<Fin><<<1, 1, 0, c->stream>>>(fin, ctl, slots, skip_early ? 1 : 0);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// y = A x | y = b - A x with fused dots; dispatches on the kernel kind chosen at upload.
template <class Epi, bool RESID>
static int kb_launch_spmv(kb_csr_s* A, const double* x, double* y, const double* b, const double* w, double* partials,
                          size_t pstride, Epi epi) {
    if (A->n == 0) return KB_OK;
    kb_ctx_s* c = A->ctx;
    KbSpmvArgs a{};
    a.row_ptr = A->row_ptr; a.col = A->col; a.vals = A->vals; a.x = x; a.y = y; a.b = b; a.w = w;
    a.n = (int)A->n; a.tile0 = 0; a.ntiles_launch = A->ntiles; a.ntiles_total = A->ntiles; a.tile_list = nullptr; a.finalize = 1;
    a.partials = partials; a.pstride = pstride; a.ticket = c->ticket;
    if (A->kind == 2) {
        auto kfn = kb_spmv_bulk<Epi, RESID>;
        if (!c->configured.count((const void*)kfn)) {
            KB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbBulkSmem)));
            c->configured.insert((const void*)kfn);
        }
        KbChunkTable tb{A->tile_chunk, A->chunk_row, A->chunk_nz};
        const int grid = std::min(2 * c->sm_count, A->ntiles);
        KbLaunch L(c, KB_K_SPMV);
        kfn<<<grid, KB_BULK_THREADS, sizeof(KbBulkSmem), c->stream>>>(a, tb, epi);
        KB_CUDA(cudaGetLastError());
        return KB_OK;
    }
    KbLaunch L(c, KB_K_SPMV);
    if (A->kind == 0) kb_spmv_stream<Epi, RESID><<<A->ntiles, KB_THREADS, 0, c->stream>>>(a, epi);
    else if (A->vec == 8) kb_spmv_vector<Epi, RESID, 8><<<A->ntiles, KB_THREADS, 0, c->stream>>>(a, epi);
    else if (A->vec == 16) kb_spmv_vector<Epi, RESID, 16><<<A->ntiles, KB_THREADS, 0, c->stream>>>(a, epi);
    else kb_spmv_vector<Epi, RESID, 32><<<A->ntiles, KB_THREADS, 0, c->stream>>>(a, epi);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}
