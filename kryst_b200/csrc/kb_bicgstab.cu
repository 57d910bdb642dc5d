// kb_bicgstab.cu — BiCGStab, device-resident (replaces src/solver/bicgstab.rs:69-293).
//
// Two variants behind one driver:
//   literal  (default): the reference's behaviour — preconditioner ignored (bicgstab.rs:70), ABSOLUTE
//            tolerance (:98,:189,:281), |.| < f64::EPSILON breakdown `break`s that return Ok with the
//            previous stats (:117,:161,:235,:285).
//   textbook (KB_FLAG_TEXTBOOK): right-preconditioned (p^ = M^-1 p, s^ = M^-1 s), relative tolerance,
//            relative breakdown tests — what "BiCGStab + Jacobi" (config 3) needs.
// One iteration = 5 kernels instead of the reference's 2 SpMV + 6 reductions + 4 updates:
//   Kp   p = r + beta (p - omega v)                 [+ p^ = D^-1 p]            (bicgstab.rs:126-136)
//   Kv   v = A p^   fused with <r^,v> (, <v,v>)     -> alpha                    (:144-164)
//   Ks   s = r - alpha v  fused with ||s||^2        [+ s^ = D^-1 s]            (:166-206)
//   Kt   t = A s^   fused with <t,s>, <t,t>         -> omega                    (:208-238)
//   Kxr  x += alpha p^ + omega s^ ; r = s - omega t  fused with ||r||^2, <r^,r> -> rho, beta  (:240-289)
#include <cstring>
#include <cfloat>
#include "kb_objects.h"
#include "kb_epilogue.cuh"
#include "kb_driver.cuh"

#define KB_EPS DBL_EPSILON

// bicgstab.rs:105-124 — test rho, compute beta for iteration i (the iteration about to start)
__device__ __forceinline__ void bicg_pre_iteration(KbCtl* c, unsigned long long i) {
    const bool lit = !c->textbook;
    if (lit ? (fabs(c->rho) < KB_EPS) : (fabs(c->rho) < KB_EPS * c->res0 * c->rnorm)) { c->breakdown = 1; c->done = 1; return; }
    c->beta = (i == 1) ? 0.0 : (c->rho / c->rho_prev) * (c->alpha / c->omega_prev);
}
struct BicgInitFin {   // bicgstab.rs:85-102
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double res0 = sqrt(s[0]);
        c->res0 = res0; c->res = res0; c->rnorm = res0;
        c->thr = c->textbook ? c->tol * res0 : c->tol;
        if (res0 <= c->thr) { c->converged = 1; c->done = 1; return; }
        if (c->max_iters == 0) { c->done = 1; return; }
        c->rho = s[0];            // <r^, r> with r^ == r
        bicg_pre_iteration(c, 1);
    }
};
struct BicgVFin {      // bicgstab.rs:147-164
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double ad = s[0];
        c->alpha_den = ad;
        const bool lit = !c->textbook;
        if (lit ? (fabs(ad) < KB_EPS) : (fabs(ad) < KB_EPS * c->res0 * sqrt(s[1]))) { c->breakdown = 2; c->done = 1; return; }
        c->alpha = c->rho / ad;
    }
};
struct BicgSFin {      // bicgstab.rs:176-206
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double sn = sqrt(s[0]);
        if (sn <= c->thr) {
            c->early = 1; c->iter = c->iter + 1; c->res = sn; c->converged = 1;
            if (c->hist_len < c->hist_cap) c->hist[c->hist_len] = sn;
            c->hist_len += 1;
        }
    }
};
struct BicgTFin {      // bicgstab.rs:211-239
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double on = s[0], od = s[1];
        if (!c->textbook ? (fabs(od) < KB_EPS) : (od == 0.0)) { c->breakdown = 3; c->done = 1; return; }
        c->omega = on / od;
    }
};
struct BicgXrFin {     // bicgstab.rs:266-289 and the head of the next iteration (:105-124)
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double rn = sqrt(s[0]);
        c->rnorm = rn; c->res = rn;
        c->iter = c->iter + 1;
        if (c->hist_len < c->hist_cap) c->hist[c->hist_len] = rn;          // one entry per iteration (the reference keeps none)
        c->hist_len += 1;
        c->converged = (rn <= c->thr) ? 1 : 0;
        if (c->converged) { c->done = 1; return; }
        if (!c->textbook ? (fabs(c->omega) < KB_EPS) : (c->omega == 0.0)) { c->breakdown = 4; c->done = 1; return; }
        c->rho_prev = c->rho; c->omega_prev = c->omega;
        if (c->iter >= c->max_iters) { c->done = 1; return; }
        c->rho = s[1];
        bicg_pre_iteration(c, c->iter + 1);
    }
};

// ---- fused vector passes ----------------------------------------------------------------------------
struct BicgInitOp : KbRedBase {      // r^ = r ; p = r ; v = 0
    static constexpr int NRED = 0;
    const double* r; double* rhat; double* p; double* v;
    __device__ bool skip() const { return false; }
    __device__ void pair(long long i, bool has1, double*) const {
        if (has1) { double2 rr = kb_ld2(r + i); kb_st2(rhat + i, rr); kb_st2(p + i, rr); kb_st2(v + i, make_double2(0.0, 0.0)); }
        else { rhat[i] = r[i]; p[i] = r[i]; v[i] = 0.0; }
    }
    __device__ void finish_block(double*) const {}
};
struct BicgPOp : KbRedBase {         // Kp
    static constexpr int NRED = 0;
    const double* r; double* p; const double* v; const double* inv; double* ph; KbCtl* ctl;
    __device__ bool skip() const { return ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double*) const {
        const double beta = ctl->beta, om = ctl->omega_prev;
        if (has1) {
            double2 rr = kb_ld2(r + i), pp = kb_ld2(p + i), vv = kb_ld2(v + i);
            pp.x = rr.x + beta * (pp.x - om * vv.x); pp.y = rr.y + beta * (pp.y - om * vv.y);
            kb_st2(p + i, pp);
            if (inv) { double2 d = kb_ld2(inv + i); kb_st2(ph + i, make_double2(d.x * pp.x, d.y * pp.y)); }
        } else {
            double pp = r[i] + beta * (p[i] - om * v[i]);
            p[i] = pp;
            if (inv) ph[i] = inv[i] * pp;
        }
    }
    __device__ void finish_block(double*) const {}
};
template <class Fin>
struct BicgSOp : KbRedBase {         // Ks
    static constexpr int NRED = 1;
    const double* r; const double* v; double* s; const double* inv; double* sh; KbCtl* ctl; KbFinish<Fin> fin;
    __device__ bool skip() const { return ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double* red) const {
        const double alpha = ctl->alpha;
        if (has1) {
            double2 rr = kb_ld2(r + i), vv = kb_ld2(v + i), ss;
            ss.x = rr.x - alpha * vv.x; ss.y = rr.y - alpha * vv.y;
            kb_st2(s + i, ss);
            if (inv) { double2 d = kb_ld2(inv + i); kb_st2(sh + i, make_double2(d.x * ss.x, d.y * ss.y)); }
            red[0] = ss.x * ss.x + ss.y * ss.y;
        } else {
            double ss = r[i] - alpha * v[i];
            s[i] = ss;
            if (inv) sh[i] = inv[i] * ss;
            red[0] = ss * ss + 0.0;
        }
    }
    __device__ void finish_block(double* sums) const { fin.template coop<0>(sums); }
};
template <class Fin>
struct BicgXrOp : KbRedBase {        // Kxr (also the early-exit x += alpha p^, bicgstab.rs:189-206)
    static constexpr int NRED = 2;
    double* x; const double* ph; const double* sh; const double* s; const double* t; double* r; const double* rhat; KbCtl* ctl;
    KbFinish<Fin> fin;
    __device__ bool skip() const { return ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double* red) const {
        const double alpha = ctl->alpha;
        if (ctl->early) {
            if (has1) { double2 xx = kb_ld2(x + i), pp = kb_ld2(ph + i); kb_st2(x + i, make_double2(xx.x + alpha * pp.x, xx.y + alpha * pp.y)); }
            else x[i] = x[i] + alpha * ph[i];
            red[0] = 0.0; red[1] = 0.0;
            return;
        }
        const double om = ctl->omega;
        if (has1) {
            double2 xx = kb_ld2(x + i), pp = kb_ld2(ph + i), hh = kb_ld2(sh + i), ss = kb_ld2(s + i), tt = kb_ld2(t + i), rh = kb_ld2(rhat + i), rr;
            xx.x = xx.x + alpha * pp.x + om * hh.x; xx.y = xx.y + alpha * pp.y + om * hh.y;
            rr.x = ss.x - om * tt.x; rr.y = ss.y - om * tt.y;
            kb_st2(x + i, xx); kb_st2(r + i, rr);
            red[0] = rr.x * rr.x + rr.y * rr.y;
            red[1] = rh.x * rr.x + rh.y * rr.y;
        } else {
            x[i] = x[i] + alpha * ph[i] + om * sh[i];
            double rr = s[i] - om * t[i];
            r[i] = rr;
            red[0] = rr * rr + 0.0;
            red[1] = rhat[i] * rr + 0.0;
        }
    }
    __device__ void finish_block(double* sums) const {
        if (ctl->early) { if (threadIdx.x == 0) ctl->done = 1; return; }   // identical on every rank: nobody enters the collective
        fin.template coop<0>(sums);
    }
};

// ---- workspace -----------------------------------------------------------------------------------------
struct KbBicgWs {
    uint64_t n = 0, nx = 0;
    double *x = nullptr, *b = nullptr, *r = nullptr, *rhat = nullptr, *p = nullptr, *v = nullptr, *s = nullptr, *t = nullptr;
    double *ph = nullptr, *sh = nullptr;     // operands of the SpMVs (with ghost space); alias p/s when unpreconditioned
    double* partials = nullptr; size_t pstride = 0;
    double* slots = nullptr;
    KbCtl* ctl = nullptr; KbCtl* h_ctl = nullptr;
    KbGraphCache gc;
};
void kb_bicg_ws_free(KbBicgWs* w) {
    if (!w) return;
    w->gc.reset();
    KB_FREE(w->x); KB_FREE(w->b); KB_FREE(w->r); KB_FREE(w->rhat); KB_FREE(w->p); KB_FREE(w->v); KB_FREE(w->s); KB_FREE(w->t);
    KB_FREE(w->ph); KB_FREE(w->sh); KB_FREE(w->partials); KB_FREE(w->slots); KB_FREE(w->ctl);
    if (w->h_ctl) cudaFreeHost(w->h_ctl);
    delete w;
}
static int bicg_ws_get(kb_csr_s* A, KbBicgWs** out) {
    KbBicgWs* w = A->bicg_ws;
    if (!w) {
        w = new KbBicgWs; A->bicg_ws = w;
        w->n = A->n; w->nx = A->ncols_local;
        KB_TRY(kb_alloc(&w->x, w->nx + 2)); KB_TRY(kb_alloc(&w->b, w->n + 2)); KB_TRY(kb_alloc(&w->r, w->n + 2));
        KB_TRY(kb_alloc(&w->rhat, w->n + 2)); KB_TRY(kb_alloc(&w->p, w->nx + 2)); KB_TRY(kb_alloc(&w->v, w->n + 2));
        KB_TRY(kb_alloc(&w->s, w->nx + 2)); KB_TRY(kb_alloc(&w->t, w->n + 2));
        KB_TRY(kb_alloc(&w->ph, w->nx + 2)); KB_TRY(kb_alloc(&w->sh, w->nx + 2));
        w->pstride = (size_t)A->ntiles + 1;
        KB_TRY(kb_alloc(&w->partials, 2 * w->pstride));
        KB_TRY(kb_alloc(&w->slots, 64));
        KB_TRY(kb_alloc(&w->ctl, 1));
        KB_CUDA(cudaMallocHost((void**)&w->h_ctl, sizeof(KbCtl)));
    }
    *out = w;
    return KB_OK;
}

template <class Op>
static int launch_tile(kb_csr_s* A, Op& op, int cls) {
    kb_ctx_s* c = A->ctx;
    op.n = (long long)A->n; op.ticket = c->ticket;
    { KbLaunch L(c, cls); kb_tile_kernel<<<A->ntiles, KB_THREADS, 0, c->stream>>>(op); }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// mode: 0 literal/unpreconditioned, 1 textbook + Jacobi (fused), 2 textbook + generic pc apply
static int bicg_iteration(kb_csr_s* A, kb_pc_s* pc, KbBicgWs* w, int mode, bool dist) {
    kb_ctx_s* c = A->ctx;
    const double* inv = (mode == 1) ? pc->inv_diag : nullptr;
    double* ph = (mode == 0) ? w->p : w->ph;
    double* sh = (mode == 0) ? w->s : w->sh;
    {   // Kp
        BicgPOp op; op.partials = nullptr; op.pstride = 0; op.r = w->r; op.p = w->p; op.v = w->v; op.inv = inv; op.ph = w->ph; op.ctl = w->ctl;
        KB_TRY(launch_tile(A, op, KB_K_BICG));
        if (mode == 2) KB_TRY(kb_pc_apply_dev(pc, w->p, w->ph));
    }
    {   // Kv
        KbSpmvEpi<BicgVFin, true, true> epi; epi.ctl = w->ctl; epi.fin = kb_make_fin(c, BicgVFin{w->ctl}, dist, w->slots, 2);
        KB_TRY((kb_launch_spmv<KbSpmvEpi<BicgVFin, true, true>, false>(A, ph, w->v, nullptr, w->rhat, w->partials, w->pstride, epi, dist ? ph : nullptr)));
        if (dist) KB_TRY((kb_finish_dist<BicgVFin>(c, BicgVFin{w->ctl}, w->ctl, w->slots, 2)));
    }
    {   // Ks
        BicgSOp<BicgSFin> op; op.partials = w->partials; op.pstride = w->pstride; op.r = w->r; op.v = w->v; op.s = w->s; op.inv = inv; op.sh = w->sh;
        op.ctl = w->ctl; op.fin = kb_make_fin(c, BicgSFin{w->ctl}, dist, w->slots, 1);
        KB_TRY(launch_tile(A, op, KB_K_BICG));
        if (dist) KB_TRY((kb_finish_dist<BicgSFin>(c, BicgSFin{w->ctl}, w->ctl, w->slots, 1)));
        if (mode == 2) KB_TRY(kb_pc_apply_dev(pc, w->s, w->sh));
    }
    {   // Kt (skipped on early exit)
        KbSpmvEpi<BicgTFin, true, true> epi; epi.ctl = w->ctl; epi.skip_mask = 1; epi.fin = kb_make_fin(c, BicgTFin{w->ctl}, dist, w->slots, 2);
        KB_TRY((kb_launch_spmv<KbSpmvEpi<BicgTFin, true, true>, false>(A, sh, w->t, nullptr, w->s, w->partials, w->pstride, epi, dist ? sh : nullptr)));
        if (dist) KB_TRY((kb_finish_dist<BicgTFin>(c, BicgTFin{w->ctl}, w->ctl, w->slots, 2, true)));
    }
    {   // Kxr
        BicgXrOp<BicgXrFin> op; op.partials = w->partials; op.pstride = w->pstride; op.x = w->x; op.ph = ph; op.sh = sh; op.s = w->s; op.t = w->t;
        op.r = w->r; op.rhat = w->rhat; op.ctl = w->ctl; op.fin = kb_make_fin(c, BicgXrFin{w->ctl}, dist, w->slots, 2);
        KB_TRY(launch_tile(A, op, KB_K_BICG));
        if (dist) KB_TRY((kb_finish_dist<BicgXrFin>(c, BicgXrFin{w->ctl}, w->ctl, w->slots, 2)));
    }
    return KB_OK;
}

extern "C" int kb_bicgstab_solve(kb_csr A, kb_pc pc, const double* b, double* x, double tol, uint64_t max_iters, uint32_t flags,
                                 kb_stats* stats) {
    if (!A || !b || !x || !stats) { kb_set_error("kb_bicgstab_solve: null argument"); return KB_SOLVE_ERROR; }
    if (pc && pc->a != A) { kb_set_error("preconditioner was set up for a different operator"); return KB_SOLVE_ERROR; }
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    const bool dev = (flags & KB_FLAG_DEVICE_PTRS) != 0;
    const bool dist = A->dist && c->size > 1;
    const bool textbook = (flags & KB_FLAG_TEXTBOOK) != 0;
    // literal: `let _ = pc;` (bicgstab.rs:70)
    const int mode = (!textbook || !pc) ? 0 : (pc->kind == KB_PC_JACOBI ? 1 : 2);
    KbBicgWs* w = nullptr;
    KB_TRY(bicg_ws_get(A, &w));
    memset(stats, 0, sizeof(*stats));
    if (A->n == 0 && !dist) return KB_OK;
    KB_TRY(kb_upload_or_alias(c, b, w->b, w->n, dev));
    KB_TRY(kb_upload_or_alias(c, x, w->x, w->n, dev));
    KbCtl* h = w->h_ctl;
    memset(h, 0, offsetof(KbCtl, h));
    h->max_iters = max_iters; h->tol = tol; h->textbook = textbook ? 1 : 0;
    h->rho_prev = 1.0; h->alpha = 1.0; h->omega_prev = 1.0;
    KB_TRY(kb_hist_prepare(A, flags, max_iters, h));
    KbMonitor mon;
    if ((flags & KB_FLAG_MONITOR) && A->monitor) { mon.fn = A->monitor; mon.user = A->monitor_user; mon.d_hist = h->hist; mon.cap = h->hist_cap; mon.index_offset = 1; }
    KB_CUDA(cudaMemcpyAsync(w->ctl, h, offsetof(KbCtl, h), cudaMemcpyHostToDevice, c->stream));
    const bool profile = (flags & KB_FLAG_PROFILE) != 0;
    const bool use_graph = !(flags & (KB_FLAG_NO_GRAPH | KB_FLAG_PROFILE));
    const bool was_prof = c->profiling;
    c->profiling = profile;
    int st = KB_OK;
    do {
        // r = b - A x ; res0 = ||r|| (bicgstab.rs:73-102)
        {
            KbSpmvEpi<BicgInitFin, false, true> epi; epi.ctl = nullptr; epi.fin = kb_make_fin(c, BicgInitFin{w->ctl}, dist, w->slots, 1);
            if ((st = kb_launch_spmv<KbSpmvEpi<BicgInitFin, false, true>, true>(A, w->x, w->r, w->b, nullptr, w->partials, w->pstride, epi, dist ? w->x : nullptr)) != KB_OK) break;
            if (dist && (st = kb_finish_dist<BicgInitFin>(c, BicgInitFin{w->ctl}, w->ctl, w->slots, 1)) != KB_OK) break;
        }
        {
            BicgInitOp op; op.partials = nullptr; op.pstride = 0; op.r = w->r; op.rhat = w->rhat; op.p = w->p; op.v = w->v;
            if ((st = launch_tile(A, op, KB_K_INIT)) != KB_OK) break;
        }
        const double bytes_iter = 24.0 * (double)A->nnz + 170.0 * (double)A->n;
        const int B = kb_batch_size(bytes_iter, 5);
        st = kb_run_iterations(c, &w->gc, (kb_pc_serial(pc) + 1) * 4 + (uint64_t)mode, B, max_iters, use_graph, w->ctl, h,
                               [&]() { return bicg_iteration(A, pc, w, mode, dist); }, &mon);
        if (st != KB_OK) break;
        if (cudaMemcpyAsync(h, w->ctl, offsetof(KbCtl, h), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("bicgstab: readback failed"); st = KB_SOLVE_ERROR; break; }
        stats->iterations = h->iter; stats->final_residual = h->res; stats->converged = h->converged; stats->breakdown = h->breakdown;
        A->hist_len = std::min<uint64_t>(h->hist_len, h->hist_cap);
        st = h->status;
        if (dist && kb_p2p_error(c)) { kb_set_error("%s: peer-memory collective timed out", "bicgstab"); st = KB_SOLVE_ERROR; break; }
        if (pc && kb_ilu0_error(const_cast<kb_pc_s*>(pc))) { kb_set_error("%s: a triangular-solve dependency wait timed out", "bicgstab"); st = KB_SOLVE_ERROR; break; }
        if (st == KB_OK) {
            if (cudaMemcpyAsync(x, w->x, w->n * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("bicgstab: copy-out of x failed"); st = KB_SOLVE_ERROR; }
        }
    } while (0);
    c->profiling = was_prof;
    if (profile) kb_prof_collect(c);
    return st;
}
