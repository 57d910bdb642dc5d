// kb_pcg.cu — Preconditioned CG, device-resident (replaces src/solver/pcg.rs:114-222).
//
// Per iteration the reference runs 1 SpMV, 3 dots, 3 vector updates and 1 pc
// apply as ~18 separate vector passes.  Here one iteration is three kernels:
//   K2  ap = A p          fused with  p.Ap  -> alpha = rz / pAp          (IndefiniteMatrix test)
//   K3  x += a p ; r -= a ap ; z = D^-1 r   fused with  r.z  and  ||r||^2 (or ||z||^2)
//       -> history push, Convergence::check, beta = rz_new/rz            (IndefinitePreconditioner test)
//   K4  p = z + beta p
// All scalars stay in the device control block; the host replays a CUDA graph of several
// iterations and polls `done` once per replay.  Kernels of iterations issued after convergence
// exit immediately, so the reported iteration count is exactly the reference's.
#include <cstring>
#include <cstddef>
#include <cstdlib>
#include <algorithm>
#include "kb_objects.h"
#include "kb_spmv.cuh"
#include "kb_epilogue.cuh"
#include "kb_driver.cuh"
#include "kb_pcg_mega.cuh"
#include "kb_pcg_resident.cuh"

// ---- scalar epilogues (device, single thread) ------------------------------------------------
struct PcgInitFin {   // pcg.rs:132-146
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        c->rz = s[0];
        c->res0 = sqrt(fabs(s[0]));
        // first history entry: dp.sqrt() of the raw dot (pcg.rs:137-146: no abs, NaN when r.z < 0); the loop takes abs (:191)
        double nrm = (c->norm_type == KB_NORM_PRECONDITIONED || c->norm_type == KB_NORM_UNPRECONDITIONED) ? sqrt(s[1])
                     : (c->norm_type == KB_NORM_NATURAL ? sqrt(s[0]) : 0.0);
        c->res = nrm;
        if (c->hist_len < c->hist_cap) c->hist[c->hist_len] = nrm;
        c->hist_len += 1;
        if (c->max_iters == 0) { c->res = c->res0; c->done = 1; }   // loop body never runs: stats = {0,res0,false}
    }
};
struct PcgApFin {     // pcg.rs:151-173
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        c->pAp = s[0];
        if (s[0] <= 0.0) { c->status = KB_INDEFINITE_MATRIX; c->iter = c->iter + 1; c->converged = 0; c->done = 1; return; }
        c->alpha = c->rz / s[0];
    }
};
struct PcgUpdateFin { // pcg.rs:188-218
    KbCtl* ctl;
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double rz_new = s[0];
        double res = (c->norm_type == KB_NORM_PRECONDITIONED || c->norm_type == KB_NORM_UNPRECONDITIONED) ? sqrt(s[1])
                     : (c->norm_type == KB_NORM_NATURAL ? sqrt(fabs(rz_new)) : 0.0);
        const unsigned long long it = c->iter + 1;
        c->iter = it;
        c->res = res;
        if (c->hist_len < c->hist_cap) c->hist[c->hist_len] = res;
        c->hist_len += 1;
        const double rel = res / c->res0;                       // Convergence::check (convergence.rs:18-34)
        if (rel <= c->tol || it >= c->max_iters) { c->converged = 1; c->done = 1; return; }
        const double beta = rz_new / c->rz;
        if (beta < 0.0) { c->status = KB_INDEFINITE_PC; c->converged = 0; c->done = 1; return; }
        c->beta = beta;
        c->rz = rz_new;
    }
};

// ---- fused vector passes ------------------------------------------------------------------------
template <class Fin>
struct PcgInitOp : KbRedBase {        // z = D^-1 r (or r) ; p = z ; r.z ; norm
    static constexpr int NRED = 2;
    const double* r; const double* inv; double* z; double* p; KbCtl* ctl; KbFinish<Fin> fin;
    __device__ bool skip() const { return false; }
    __device__ void pair(long long i, bool has1, double* red) const {
        const int nt = ctl->norm_type;
        if (has1) {
            double2 rr = kb_ld2(r + i), zz;
            if (inv) { double2 d = kb_ld2(inv + i); zz = make_double2(d.x * rr.x, d.y * rr.y); } else zz = rr;
            kb_st2(z + i, zz); kb_st2(p + i, zz);
            red[0] = rr.x * zz.x + rr.y * zz.y;
            red[1] = nt == KB_NORM_PRECONDITIONED ? (zz.x * zz.x + zz.y * zz.y) : nt == KB_NORM_UNPRECONDITIONED ? (rr.x * rr.x + rr.y * rr.y) : 0.0;
        } else {
            double rr = r[i], zz = inv ? inv[i] * rr : rr;
            z[i] = zz; p[i] = zz;
            red[0] = rr * zz + 0.0;
            red[1] = nt == KB_NORM_PRECONDITIONED ? (zz * zz + 0.0) : nt == KB_NORM_UNPRECONDITIONED ? (rr * rr + 0.0) : 0.0;
        }
    }
    __device__ void finish_block(double* s) const { fin.template coop<0>(s); }
};

template <class Fin>
struct PcgUpdateOp : KbRedBase {      // K3
    static constexpr int NRED = 2;
    double* x; const double* p; double* r; const double* ap; const double* inv; double* z; KbCtl* ctl; KbFinish<Fin> fin;
    __device__ bool skip() const { return ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double* red) const {
        const double alpha = ctl->alpha;
        const int nt = ctl->norm_type;
        if (has1) {
            double2 xx = kb_ld2(x + i), pp = kb_ld2(p + i), rr = kb_ld2(r + i), aa = kb_ld2(ap + i), zz;
            xx.x = xx.x + alpha * pp.x; xx.y = xx.y + alpha * pp.y;
            rr.x = rr.x - alpha * aa.x; rr.y = rr.y - alpha * aa.y;
            if (inv) { double2 d = kb_ld2(inv + i); zz = make_double2(d.x * rr.x, d.y * rr.y); } else zz = rr;
            kb_st2(x + i, xx); kb_st2(r + i, rr); kb_st2(z + i, zz);
            red[0] = rr.x * zz.x + rr.y * zz.y;
            red[1] = nt == KB_NORM_PRECONDITIONED ? (zz.x * zz.x + zz.y * zz.y) : nt == KB_NORM_UNPRECONDITIONED ? (rr.x * rr.x + rr.y * rr.y) : 0.0;
        } else {
            double xx = x[i] + alpha * p[i];
            double rr = r[i] - alpha * ap[i];
            double zz = inv ? inv[i] * rr : rr;
            x[i] = xx; r[i] = rr; z[i] = zz;
            red[0] = rr * zz + 0.0;
            red[1] = nt == KB_NORM_PRECONDITIONED ? (zz * zz + 0.0) : nt == KB_NORM_UNPRECONDITIONED ? (rr * rr + 0.0) : 0.0;
        }
    }
    __device__ void finish_block(double* s) const { fin.template coop<0>(s); }
};

// --- generic preconditioner (e.g. ILU(0)): the pc apply cannot be fused, so K3 splits into update / apply / dots
struct PcgXrOp : KbRedBase {          // x += alpha p ; r -= alpha ap   (pcg.rs:175-181)
    static constexpr int NRED = 0;
    double* x; const double* p; double* r; const double* ap; KbCtl* ctl;
    __device__ bool skip() const { return ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double*) const {
        const double alpha = ctl->alpha;
        if (has1) {
            double2 xx = kb_ld2(x + i), pp = kb_ld2(p + i), rr = kb_ld2(r + i), aa = kb_ld2(ap + i);
            kb_st2(x + i, make_double2(xx.x + alpha * pp.x, xx.y + alpha * pp.y));
            kb_st2(r + i, make_double2(rr.x - alpha * aa.x, rr.y - alpha * aa.y));
        } else { x[i] = x[i] + alpha * p[i]; r[i] = r[i] - alpha * ap[i]; }
    }
    __device__ void finish_block(double*) const {}
};
template <class Fin, bool COPY_P>
struct PcgRzOp : KbRedBase {          // r.z and the norm (pcg.rs:188-195); COPY_P: p = z (pcg.rs:132)
    static constexpr int NRED = 2;
    const double* r; const double* z; double* p; KbCtl* ctl; KbFinish<Fin> fin;
    __device__ bool skip() const { return COPY_P ? false : ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double* red) const {
        const int nt = ctl->norm_type;
        if (has1) {
            double2 rr = kb_ld2(r + i), zz = kb_ld2(z + i);
            if (COPY_P) kb_st2(p + i, zz);
            red[0] = rr.x * zz.x + rr.y * zz.y;
            red[1] = nt == KB_NORM_PRECONDITIONED ? (zz.x * zz.x + zz.y * zz.y) : nt == KB_NORM_UNPRECONDITIONED ? (rr.x * rr.x + rr.y * rr.y) : 0.0;
        } else {
            double rr = r[i], zz = z[i];
            if (COPY_P) p[i] = zz;
            red[0] = rr * zz + 0.0;
            red[1] = nt == KB_NORM_PRECONDITIONED ? (zz * zz + 0.0) : nt == KB_NORM_UNPRECONDITIONED ? (rr * rr + 0.0) : 0.0;
        }
    }
    __device__ void finish_block(double* s) const { fin.template coop<0>(s); }
};

struct PcgXpayOp : KbRedBase {        // K4: p = z + beta p  (pcg.rs:215-217)
    static constexpr int NRED = 0;
    const double* z; double* p; KbCtl* ctl;
    __device__ bool skip() const { return ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double*) const {
        const double beta = ctl->beta;
        if (has1) { double2 zz = kb_ld2(z + i), pp = kb_ld2(p + i); kb_st2(p + i, make_double2(zz.x + beta * pp.x, zz.y + beta * pp.y)); }
        else p[i] = z[i] + beta * p[i];
    }
    __device__ void finish_block(double*) const {}
};

// K4 on a shard: p = z + beta p, and the tile's boundary entries of the new p go straight into the neighbours' ghost
// mailboxes (NVLink peer stores) - the separate 29 us kb_halo_push launch of round 1 is gone, and the transfer
// overlaps the rest of this kernel and the launch gap before the SpMV that consumes it.
__global__ void __launch_bounds__(KB_THREADS) kb_pcg_xpay_push(PcgXpayOp op, const KbHaloDev* __restrict__ hp) {
    kb_pdl_wait();
    kb_pdl_launch_dependents();
    if (op.skip()) return;
    const int tile = hp->tile_perm[blockIdx.x];            // sending tiles first
    const long long i = (long long)tile * KB_TILE + 2 * threadIdx.x;
    if (i + 1 < op.n) op.pair(i, true, nullptr);
    else if (i < op.n) op.pair(i, false, nullptr);
    kb_halo_push_tile(*hp, op.p, tile);
}

// ---- single-reduction variant (Chronopoulos-Gear) ---------------------------------------
// One iteration = two kernels and ONE reduction (one all-reduce on shards) instead of three kernels and two:
//   S1  p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = D^-1 r ; local sums of r.u and the norm
//   S2  w = A u  fused with  u.w ; its last CTA appends S1's local sums, all-reduces the three together and runs
//       history push, Convergence::check, beta = g'/g, p.Ap = delta - beta g'/alpha (IndefiniteMatrix test), alpha
// Bytes per iteration: B_spmv + 96 n (literal path: B_spmv + 88 n).
struct PcgSrLocalFin {   // S1: keep this rank's sums for S2's collective
    KbCtl* ctl;
    __device__ void operator()(const double* s) const { ctl->sr_loc[0] = s[0]; ctl->sr_loc[1] = s[1]; }
};
struct PcgSrFin {        // S2 epilogue; s = {u.w, r.u, norm sum} (global)
    KbCtl* ctl; int init;
    __device__ void pre(double* s) const { s[1] = ctl->sr_loc[0]; s[2] = ctl->sr_loc[1]; }
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double delta = s[0], g_new = s[1];
        const double res = (c->norm_type == KB_NORM_PRECONDITIONED || c->norm_type == KB_NORM_UNPRECONDITIONED) ? sqrt(s[2])
                           : (c->norm_type == KB_NORM_NATURAL ? sqrt(fabs(g_new)) : 0.0);
        if (c->hist_len < c->hist_cap) c->hist[c->hist_len] = res;
        c->hist_len += 1;
        c->res = res;
        if (init) {
            c->rz = g_new;
            c->res0 = sqrt(fabs(g_new));
            if (c->max_iters == 0) { c->res = c->res0; c->done = 1; return; }
            if (delta <= 0.0) { c->status = KB_INDEFINITE_MATRIX; c->iter = 1; c->converged = 0; c->done = 1; return; }
            c->pAp = delta; c->alpha = g_new / delta; c->beta = 0.0;
            return;
        }
        const unsigned long long it = c->iter + 1;
        c->iter = it;
        const double rel = res / c->res0;                       // Convergence::check (convergence.rs:18-34)
        if (rel <= c->tol || it >= c->max_iters) { c->converged = 1; c->done = 1; return; }
        const double beta = g_new / c->rz;
        if (beta < 0.0) { c->status = KB_INDEFINITE_PC; c->converged = 0; c->done = 1; return; }
        const double pAp = delta - beta * g_new / c->alpha;
        if (pAp <= 0.0) { c->status = KB_INDEFINITE_MATRIX; c->iter = it + 1; c->converged = 0; c->done = 1; return; }
        c->pAp = pAp; c->beta = beta; c->alpha = g_new / pAp; c->rz = g_new;
    }
};
template <class Fin, bool INIT>
struct PcgSrOp : KbRedBase {          // S1 (INIT: u = D^-1 r ; p = s = 0 and the sums only)
    static constexpr int NRED = 2;
    double* x; double* p; double* r; double* s; const double* w; const double* inv; double* u; KbCtl* ctl; KbFinish<Fin> fin;
    __device__ bool skip() const { return INIT ? false : ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double* red) const {
        const int nt = ctl->norm_type;
        if (has1) {
            double2 rr = kb_ld2(r + i), uu;
            if (!INIT) {
                const double alpha = ctl->alpha, beta = ctl->beta;
                double2 pp = kb_ld2(p + i), ss = kb_ld2(s + i), ww = kb_ld2(w + i), xx = kb_ld2(x + i), u0 = kb_ld2(u + i);
                pp.x = u0.x + beta * pp.x; pp.y = u0.y + beta * pp.y;
                ss.x = ww.x + beta * ss.x; ss.y = ww.y + beta * ss.y;
                xx.x = xx.x + alpha * pp.x; xx.y = xx.y + alpha * pp.y;
                rr.x = rr.x - alpha * ss.x; rr.y = rr.y - alpha * ss.y;
                kb_st2(p + i, pp); kb_st2(s + i, ss); kb_st2(x + i, xx); kb_st2(r + i, rr);
            } else { kb_st2(p + i, make_double2(0.0, 0.0)); kb_st2(s + i, make_double2(0.0, 0.0)); }
            if (inv) { double2 d = kb_ld2(inv + i); uu = make_double2(d.x * rr.x, d.y * rr.y); } else uu = rr;
            kb_st2(u + i, uu);
            red[0] = rr.x * uu.x + rr.y * uu.y;
            red[1] = nt == KB_NORM_PRECONDITIONED ? (uu.x * uu.x + uu.y * uu.y) : nt == KB_NORM_UNPRECONDITIONED ? (rr.x * rr.x + rr.y * rr.y) : 0.0;
        } else {
            double rr = r[i];
            if (!INIT) {
                const double alpha = ctl->alpha, beta = ctl->beta;
                const double pp = u[i] + beta * p[i];
                const double ss = w[i] + beta * s[i];
                x[i] = x[i] + alpha * pp;
                rr = rr - alpha * ss;
                p[i] = pp; s[i] = ss; r[i] = rr;
            } else { p[i] = 0.0; s[i] = 0.0; }
            const double uu = inv ? inv[i] * rr : rr;
            u[i] = uu;
            red[0] = rr * uu + 0.0;
            red[1] = nt == KB_NORM_PRECONDITIONED ? (uu * uu + 0.0) : nt == KB_NORM_UNPRECONDITIONED ? (rr * rr + 0.0) : 0.0;
        }
    }
    __device__ void finish_block(double* sm) const { fin.template coop<0>(sm); }
};

// ---- pipelined variant (Ghysels-Vanroose) --------------------------------------------------
// One iteration = two kernels; its ONE reduction is SENT by the first and RECEIVED by the second:
//   P1  z = n + beta z ; q = m + beta q ; s = w + beta s ; p = u + beta p ; x += alpha p ; r -= alpha s ;
//       u -= alpha q ; w -= alpha z ; m = D^-1 w ; sums of r.u, w.u and the norm; on shards its last CTA stores the
//       three sums into every peer's mailbox (kb_p2p_allreduce_send) and returns
//   P2  n = A m ; its last CTA receives the sums (they crossed NVLink while the SpMV ran), then history push,
//       Convergence::check and the Chronopoulos-Gear scalars (beta = g'/g, p.Ap = delta - beta g'/alpha, alpha)
// Bytes per iteration: B_spmv + 168 n (literal path: B_spmv + 88 n): more vector traffic for a hidden all-reduce.
struct PcgPipeFin {      // P2 epilogue; s = {m.n (unused), r.u, w.u, norm sum} (global)
    KbCtl* ctl;
    __device__ void pre(double* s) const { s[1] = ctl->sr_loc[0]; s[2] = ctl->sr_loc[1]; s[3] = ctl->sr_loc[2]; }
    __device__ void operator()(const double* s) const {
        KbCtl* c = ctl;
        const double g_new = s[1], delta = s[2];
        const double res = (c->norm_type == KB_NORM_PRECONDITIONED || c->norm_type == KB_NORM_UNPRECONDITIONED) ? sqrt(s[3])
                           : (c->norm_type == KB_NORM_NATURAL ? sqrt(fabs(g_new)) : 0.0);
        if (c->hist_len < c->hist_cap) c->hist[c->hist_len] = res;
        c->hist_len += 1;
        c->res = res;
        const unsigned long long it = c->iter + 1;
        c->iter = it;
        const double rel = res / c->res0;                       // Convergence::check (convergence.rs:18-34)
        if (rel <= c->tol || it >= c->max_iters) { c->converged = 1; c->done = 1; return; }
        const double beta = g_new / c->rz;
        if (beta < 0.0) { c->status = KB_INDEFINITE_PC; c->converged = 0; c->done = 1; return; }
        const double pAp = delta - beta * g_new / c->alpha;
        if (pAp <= 0.0) { c->status = KB_INDEFINITE_MATRIX; c->iter = it + 1; c->converged = 0; c->done = 1; return; }
        c->pAp = pAp; c->beta = beta; c->alpha = g_new / pAp; c->rz = g_new;
    }
};
struct PcgPipeInitOp : KbRedBase {     // m = D^-1 w ; z = q = 0   (after the single-reduction set-up: u, w, p = s = 0 and the scalars exist)
    static constexpr int NRED = 0;
    const double* w; const double* inv; double* m; double* z; double* q;
    __device__ bool skip() const { return false; }
    __device__ void pair(long long i, bool has1, double*) const {
        if (has1) {
            double2 ww = kb_ld2(w + i);
            if (inv) { const double2 d = kb_ld2(inv + i); ww = make_double2(d.x * ww.x, d.y * ww.y); }
            kb_st2(m + i, ww); kb_st2(z + i, make_double2(0.0, 0.0)); kb_st2(q + i, make_double2(0.0, 0.0));
        } else { m[i] = inv ? inv[i] * w[i] : w[i]; z[i] = 0.0; q[i] = 0.0; }
    }
    __device__ void finish_block(double*) const {}
};
struct PcgPipeOp : KbRedBase {         // P1
    static constexpr int NRED = 3;
    double *x, *r, *u, *w, *m, *z, *q, *s, *p; const double* nn; const double* inv; KbCtl* ctl; const KbP2PDev* p2p;
    __device__ bool skip() const { return ctl->done != 0; }
    __device__ void pair(long long i, bool has1, double* red) const {
        const int nt = ctl->norm_type;
        const double alpha = ctl->alpha, beta = ctl->beta;
        if (has1) {
            double2 zz = kb_ld2(z + i), qq = kb_ld2(q + i), ss = kb_ld2(s + i), pp = kb_ld2(p + i);
            const double2 n2 = kb_ld2(nn + i), m2 = kb_ld2(m + i);
            double2 ww = kb_ld2(w + i), uu = kb_ld2(u + i), xx = kb_ld2(x + i), rr = kb_ld2(r + i);
            zz.x = n2.x + beta * zz.x; zz.y = n2.y + beta * zz.y;
            qq.x = m2.x + beta * qq.x; qq.y = m2.y + beta * qq.y;
            ss.x = ww.x + beta * ss.x; ss.y = ww.y + beta * ss.y;
            pp.x = uu.x + beta * pp.x; pp.y = uu.y + beta * pp.y;
            xx.x = xx.x + alpha * pp.x; xx.y = xx.y + alpha * pp.y;
            rr.x = rr.x - alpha * ss.x; rr.y = rr.y - alpha * ss.y;
            uu.x = uu.x - alpha * qq.x; uu.y = uu.y - alpha * qq.y;
            ww.x = ww.x - alpha * zz.x; ww.y = ww.y - alpha * zz.y;
            kb_st2(z + i, zz); kb_st2(q + i, qq); kb_st2(s + i, ss); kb_st2(p + i, pp);
            kb_st2(x + i, xx); kb_st2(r + i, rr); kb_st2(u + i, uu); kb_st2(w + i, ww);
            double2 mm = ww;
            if (inv) { const double2 d = kb_ld2(inv + i); mm = make_double2(d.x * ww.x, d.y * ww.y); }
            kb_st2(m + i, mm);
            red[0] = rr.x * uu.x + rr.y * uu.y;
            red[1] = ww.x * uu.x + ww.y * uu.y;
            red[2] = nt == KB_NORM_PRECONDITIONED ? (uu.x * uu.x + uu.y * uu.y) : nt == KB_NORM_UNPRECONDITIONED ? (rr.x * rr.x + rr.y * rr.y) : 0.0;
        } else {
            const double zz = nn[i] + beta * z[i];
            const double qq = m[i] + beta * q[i];
            const double ss = w[i] + beta * s[i];
            const double pp = u[i] + beta * p[i];
            const double xx = x[i] + alpha * pp;
            const double rr = r[i] - alpha * ss;
            const double uu = u[i] - alpha * qq;
            const double ww = w[i] - alpha * zz;
            z[i] = zz; q[i] = qq; s[i] = ss; p[i] = pp; x[i] = xx; r[i] = rr; u[i] = uu; w[i] = ww;
            m[i] = inv ? inv[i] * ww : ww;
            red[0] = rr * uu + 0.0;
            red[1] = ww * uu + 0.0;
            red[2] = nt == KB_NORM_PRECONDITIONED ? (uu * uu + 0.0) : nt == KB_NORM_UNPRECONDITIONED ? (rr * rr + 0.0) : 0.0;
        }
    }
    __device__ void finish_block(double* sm) const {      // called by all threads of the last CTA
        if (p2p) kb_p2p_allreduce_send<0>(*p2p, sm, 3);
        else if (threadIdx.x == 0) { ctl->sr_loc[0] = sm[0]; ctl->sr_loc[1] = sm[1]; ctl->sr_loc[2] = sm[2]; }
    }
};

// ---- workspace ----------------------------------------------------------------------------------
struct KbPcgWs {
    uint64_t n = 0, nx = 0;
    double *x = nullptr, *r = nullptr, *z = nullptr, *p = nullptr, *ap = nullptr, *b = nullptr;
    double* partials = nullptr; size_t pstride = 0;
    double* slots = nullptr;          // dist: local sums / gathered sums
    KbCtl* ctl = nullptr; KbCtl* h_ctl = nullptr;
    double* hist = nullptr; uint64_t hist_cap = 0;
    KbGraphCache gc;
    unsigned* mega_bar = nullptr;     // grid-barrier words of the persistent kernel
    ulonglong2* res_pk = nullptr;     // resident kernel: tagged packets [n + 3 ntiles], marks [n], barrier / verdict words
    unsigned char* res_needed = nullptr;
    unsigned* res_bar = nullptr;
    double *u_sr = nullptr, *s_sr = nullptr;   // single-reduction variant: u (SpMV operand, with ghost tail) and s = A p
    double *m_pp = nullptr, *n_pp = nullptr, *q_pp = nullptr;   // pipelined variant: m = M^-1 w (SpMV operand, with ghost tail), n = A m, q
};
void kb_pcg_ws_free(KbPcgWs* w) {
    if (!w) return;
    w->gc.reset();
    KB_FREE(w->x); KB_FREE(w->r); KB_FREE(w->z); KB_FREE(w->p); KB_FREE(w->ap); KB_FREE(w->b);
    KB_FREE(w->partials); KB_FREE(w->slots); KB_FREE(w->ctl); KB_FREE(w->hist); KB_FREE(w->mega_bar); KB_FREE(w->res_pk); KB_FREE(w->res_needed); KB_FREE(w->res_bar); KB_FREE(w->u_sr); KB_FREE(w->s_sr); KB_FREE(w->m_pp); KB_FREE(w->n_pp); KB_FREE(w->q_pp);
    if (w->h_ctl) cudaFreeHost(w->h_ctl);
    delete w;
}
static int pcg_ws_get(kb_csr_s* A, uint64_t hist_cap, KbPcgWs** out) {
    KbPcgWs* w = A->pcg_ws;
    if (!w) {
        w = new KbPcgWs;
        A->pcg_ws = w;
        w->n = A->n; w->nx = A->ncols_local;
        KB_TRY(kb_alloc(&w->x, w->nx + 2)); KB_TRY(kb_alloc(&w->p, w->nx + 2));
        KB_TRY(kb_alloc(&w->r, w->n + 2)); KB_TRY(kb_alloc(&w->z, w->n + 2));
        KB_TRY(kb_alloc(&w->ap, w->n + 2)); KB_TRY(kb_alloc(&w->b, w->n + 2));
        w->pstride = (size_t)A->ntiles + 1;
        KB_TRY(kb_alloc(&w->partials, 3 * w->pstride));      // up to three fused sums per kernel (pipelined variant)
        KB_TRY(kb_alloc(&w->slots, 64));
        KB_TRY(kb_alloc(&w->ctl, 1));
        KB_CUDA(cudaMallocHost((void**)&w->h_ctl, sizeof(KbCtl)));
    }
    if (hist_cap > w->hist_cap || !w->hist) {
        KB_FREE(w->hist);
        w->hist_cap = std::max<uint64_t>(hist_cap, 1);
        KB_TRY(kb_alloc(&w->hist, w->hist_cap));
    }
    *out = w;
    return KB_OK;
}

// ---- one iteration's launches ---------------------------------------------------------------------
template <bool DIST>
static int pcg_launch_iteration(kb_csr_s* A, const kb_pc_s* pc, KbPcgWs* w) {
    kb_ctx_s* c = A->ctx;
    const KbHaloDev* fused_push = DIST ? kb_halo_fused_dev(A) : nullptr;
    const double* inv = pc ? pc->inv_diag : nullptr;
    {   // K2
        KbSpmvEpi<PcgApFin, true, false> epi; epi.ctl = w->ctl; epi.fin = kb_make_fin(c, PcgApFin{w->ctl}, DIST, w->slots, 1);
        KB_TRY((kb_launch_spmv<KbSpmvEpi<PcgApFin, true, false>, false>(A, w->p, w->ap, nullptr, w->p, w->partials, w->pstride, epi, DIST ? w->p : nullptr, fused_push != nullptr, /*pdl=*/fused_push != nullptr || !DIST)));
        if (DIST) KB_TRY((kb_finish_dist<PcgApFin>(c, PcgApFin{w->ctl}, w->ctl, w->slots, 1)));
    }
    if (pc && pc->kind != KB_PC_JACOBI) {   // K3 with a generic preconditioner
        PcgXrOp u; u.n = (long long)w->n; u.partials = nullptr; u.pstride = 0; u.ticket = c->ticket;
        u.x = w->x; u.p = w->p; u.r = w->r; u.ap = w->ap; u.ctl = w->ctl;
        { KbLaunch L(c, KB_K_PCG_UPDATE); kb_tile_kernel<<<A->ntiles, KB_THREADS, 0, c->stream>>>(u); }
        KB_CUDA(cudaGetLastError());
        KB_TRY(kb_pc_apply_dev(const_cast<kb_pc_s*>(pc), w->r, w->z, w->ctl, 0));
        PcgRzOp<PcgUpdateFin, false> op; op.n = (long long)w->n; op.partials = w->partials; op.pstride = w->pstride; op.ticket = c->ticket;
        op.r = w->r; op.z = w->z; op.p = nullptr; op.ctl = w->ctl; op.fin = kb_make_fin(c, PcgUpdateFin{w->ctl}, DIST, w->slots, 2);
        { KbLaunch L(c, KB_K_PCG_UPDATE); kb_tile_kernel<<<A->ntiles, KB_THREADS, 0, c->stream>>>(op); }
        KB_CUDA(cudaGetLastError());
        if (DIST) KB_TRY((kb_finish_dist<PcgUpdateFin>(c, PcgUpdateFin{w->ctl}, w->ctl, w->slots, 2)));
    } else {   // K3
        PcgUpdateOp<PcgUpdateFin> op; op.n = (long long)w->n; op.partials = w->partials; op.pstride = w->pstride; op.ticket = c->ticket;
        op.x = w->x; op.p = w->p; op.r = w->r; op.ap = w->ap; op.inv = inv; op.z = w->z; op.ctl = w->ctl;
        op.fin = kb_make_fin(c, PcgUpdateFin{w->ctl}, DIST, w->slots, 2);
        { KbLaunch L(c, KB_K_PCG_UPDATE); KB_CUDA(kb_launch_ex(kb_pdl_enabled(), kb_tile_kernel<PcgUpdateOp<PcgUpdateFin>>, dim3(A->ntiles), dim3(KB_THREADS), 0, c->stream, op)); }
        if (DIST) KB_TRY((kb_finish_dist<PcgUpdateFin>(c, PcgUpdateFin{w->ctl}, w->ctl, w->slots, 2)));
    }
    {   // K4
        PcgXpayOp op; op.n = (long long)w->n; op.partials = nullptr; op.pstride = 0; op.ticket = c->ticket;
        op.z = w->z; op.p = w->p; op.ctl = w->ctl;
        if (fused_push) { KbLaunch L(c, KB_K_XPAY); KB_CUDA(kb_launch_ex(kb_pdl_enabled(), kb_pcg_xpay_push, dim3(A->ntiles), dim3(KB_THREADS), 0, c->stream, op, fused_push)); }
        else { KbLaunch L(c, KB_K_XPAY); KB_CUDA(kb_launch_ex(kb_pdl_enabled(), kb_tile_kernel<PcgXpayOp>, dim3(A->ntiles), dim3(KB_THREADS), 0, c->stream, op)); }
    }
    return KB_OK;
}

// single-reduction variant: S1 (INIT: setup form) then S2
template <bool DIST, bool INIT>
static int pcg_sr_launch(kb_csr_s* A, const kb_pc_s* pc, KbPcgWs* w) {
    kb_ctx_s* c = A->ctx;
    {
        PcgSrOp<PcgSrLocalFin, INIT> op; op.n = (long long)w->n; op.partials = w->partials; op.pstride = w->pstride; op.ticket = c->ticket;
        op.x = w->x; op.p = w->p; op.r = w->r; op.s = w->s_sr; op.w = w->ap; op.inv = pc ? pc->inv_diag : nullptr; op.u = w->u_sr; op.ctl = w->ctl;
        op.fin = kb_make_fin(c, PcgSrLocalFin{w->ctl}, false, nullptr, 2);       // local sums only: no collective here
        { KbLaunch L(c, INIT ? KB_K_INIT : KB_K_PCG_UPDATE); kb_tile_kernel<<<A->ntiles, KB_THREADS, 0, c->stream>>>(op); }
        KB_CUDA(cudaGetLastError());
    }
    {
        typedef KbSpmvEpi<PcgSrFin, true, false> Epi;
        Epi epi; epi.ctl = INIT ? nullptr : w->ctl; epi.fin = kb_make_fin(c, PcgSrFin{w->ctl, INIT ? 1 : 0}, DIST, w->slots, 3);
        KB_TRY((kb_launch_spmv<Epi, false>(A, w->u_sr, w->ap, nullptr, w->u_sr, w->partials, w->pstride, epi, DIST ? w->u_sr : nullptr)));
        if (DIST) KB_TRY((kb_finish_dist<PcgSrFin>(c, PcgSrFin{w->ctl, INIT ? 1 : 0}, w->ctl, w->slots, 3)));
    }
    return KB_OK;
}
static int pcg_sr_iteration(kb_csr_s* A, const kb_pc_s* pc, KbPcgWs* w) {
    return A->dist && A->ctx->size > 1 ? pcg_sr_launch<true, false>(A, pc, w) : pcg_sr_launch<false, false>(A, pc, w);
}

// pipelined variant: set-up (after the single-reduction set-up) and one iteration
template <bool DIST>
static int pcg_pipe_init(kb_csr_s* A, const kb_pc_s* pc, KbPcgWs* w) {
    kb_ctx_s* c = A->ctx;
    PcgPipeInitOp op; op.n = (long long)w->n; op.partials = nullptr; op.pstride = 0; op.ticket = c->ticket;
    op.w = w->ap; op.inv = pc ? pc->inv_diag : nullptr; op.m = w->m_pp; op.z = w->z; op.q = w->q_pp;
    { KbLaunch L(c, KB_K_INIT); kb_tile_kernel<<<std::max(A->ntiles, 1), KB_THREADS, 0, c->stream>>>(op); }
    KB_CUDA(cudaGetLastError());
    KbEpiNone epi;
    return kb_launch_spmv<KbEpiNone, false>(A, w->m_pp, w->n_pp, nullptr, nullptr, nullptr, 0, epi, DIST ? w->m_pp : nullptr);
}
template <bool DIST>
static int pcg_pipe_launch(kb_csr_s* A, const kb_pc_s* pc, KbPcgWs* w) {
    kb_ctx_s* c = A->ctx;
    const KbP2PDev* p2p = DIST ? kb_p2p_dev_ptr(c) : nullptr;
    {
        PcgPipeOp op; op.n = (long long)w->n; op.partials = w->partials; op.pstride = w->pstride; op.ticket = c->ticket;
        op.x = w->x; op.r = w->r; op.u = w->u_sr; op.w = w->ap; op.m = w->m_pp; op.z = w->z; op.q = w->q_pp; op.s = w->s_sr; op.p = w->p;
        op.nn = w->n_pp; op.inv = pc ? pc->inv_diag : nullptr; op.ctl = w->ctl; op.p2p = p2p;
        { KbLaunch L(c, KB_K_PCG_UPDATE); kb_tile_kernel<<<std::max(A->ntiles, 1), KB_THREADS, 0, c->stream>>>(op); }
        KB_CUDA(cudaGetLastError());
    }
    {
        typedef KbSpmvEpi<PcgPipeFin, true, false> Epi;
        Epi epi; epi.ctl = w->ctl; epi.fin = kb_make_fin(c, PcgPipeFin{w->ctl}, DIST, w->slots, 4);
        epi.fin.recv_only = p2p ? 1 : 0;
        KB_TRY((kb_launch_spmv<Epi, false>(A, w->m_pp, w->n_pp, nullptr, w->m_pp, w->partials, w->pstride, epi, DIST ? w->m_pp : nullptr)));
        if (DIST) KB_TRY((kb_finish_dist<PcgPipeFin>(c, PcgPipeFin{w->ctl}, w->ctl, w->slots, 4)));
    }
    return KB_OK;
}
static int pcg_pipe_iteration(kb_csr_s* A, const kb_pc_s* pc, KbPcgWs* w) {
    return A->dist && A->ctx->size > 1 ? pcg_pipe_launch<true>(A, pc, w) : pcg_pipe_launch<false>(A, pc, w);
}

// whole solve in one cooperative launch (single GPU, bulk SpMV, Jacobi or no preconditioner)
static int pcg_persistent(kb_csr_s* A, const kb_pc_s* pc, KbPcgWs* w) {
    kb_ctx_s* c = A->ctx;
    auto kfn = kb_pcg_persistent<PcgApFin, PcgUpdateFin, PcgUpdateOp<PcgUpdateFin>, PcgXpayOp>;
    if (!c->configured.count((const void*)kfn)) {
        KB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KbBulkSmem)));
        c->configured.insert((const void*)kfn);
    }
    int occ = 0;
    KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, KB_BULK_THREADS, sizeof(KbBulkSmem)));
    if (occ < 1) { kb_set_error("persistent PCG kernel does not fit on this device"); return KB_UNSUPPORTED; }
    const int grid = std::min(std::min(occ, 2) * c->sm_count, A->ntiles);
    if (!w->mega_bar) { KB_TRY(kb_alloc(&w->mega_bar, 4)); }
    KB_CUDA(cudaMemsetAsync(w->mega_bar, 0, 4 * sizeof(unsigned), c->stream));
    KbSpmvArgs a{};
    a.lazy_from = -1;
    a.row_ptr = A->row_ptr; a.col = A->col; a.vals = A->vals; a.x = w->p; a.y = w->ap; a.b = nullptr; a.w = w->p;
    a.n = (int)A->n; a.tile0 = 0; a.ntiles_launch = A->ntiles; a.ntiles_total = A->ntiles; a.tile_list = nullptr; a.finalize = 0;
    a.partials = w->partials; a.pstride = w->pstride; a.ticket = c->ticket;
    KbChunkTable tb{A->tile_chunk, A->chunk_row, A->chunk_nz};
    KbPcgMegaArgs m{};
    m.ctl = w->ctl; m.x = w->x; m.r = w->r; m.z = w->z; m.p = w->p; m.ap = w->ap; m.inv = pc ? pc->inv_diag : nullptr;
    m.partials = w->partials; m.pstride = w->pstride; m.n = (int)A->n; m.ntiles = A->ntiles; m.bar = w->mega_bar;
    void* args[] = {&a, &tb, &m};
    KbLaunch L(c, KB_K_SPMV);
    KB_CUDA(cudaLaunchCooperativeKernel((const void*)kfn, dim3(grid), dim3(KB_BULK_THREADS), args, sizeof(KbBulkSmem), c->stream));
    return KB_OK;
}

// whole solve on chip in one cooperative launch (kb_pcg_resident.cuh); *ran = false: not eligible, nothing was modified
#ifndef KB_PCG_RESIDENT_DEFAULT
#define KB_PCG_RESIDENT_DEFAULT 1
#endif
static int pcg_resident(kb_csr_s* A, const kb_pc_s* pc, KbPcgWs* w, bool* ran) {
    *ran = false;
    kb_ctx_s* c = A->ctx;
    const int L = (int)A->max_row_len;
    if (L < 1 || L > KB_RES_MAXLEN || A->n != A->ncols_local || A->n == 0) return KB_OK;
    int G = std::min(c->sm_count, A->ntiles);
    if (getenv("KB_RES_GRID")) G = std::max(1, std::min(G, atoi(getenv("KB_RES_GRID"))));      // tuning: fewer, evenly loaded CTAs
    if ((A->ntiles + G - 1) / G > KB_RES_TEAMS || A->ntiles > 3 * KB_THREADS) return KB_OK;
    const int T = (A->ntiles + G - 1) / G;
    const size_t smem_max = 226 * 1024;                            // 227 KB per CTA minus the kernel's static shared memory
    const size_t fixed = kb_res_smem_fixed(L, T);
    if (fixed + 12 * 256 > smem_max) return KB_OK;
    const int ghost_cap = (int)std::min<size_t>((smem_max - fixed) / 12, (size_t)1 << 15);
    const size_t smem = kb_res_smem_bytes(L, T, ghost_cap);
    auto kfn = kb_pcg_resident;
    if (!c->configured.count((const void*)kfn)) {
        KB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        c->configured.insert((const void*)kfn);
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, KB_RES_THREADS, smem) != cudaSuccess) { (void)cudaGetLastError(); return KB_OK; }
    if (occ < 1) return KB_OK;
    const size_t npk = (size_t)A->n + 3 * (size_t)A->ntiles;
    if (!w->res_pk) {
        KB_TRY(kb_alloc(&w->res_pk, npk));
        KB_TRY(kb_alloc(&w->res_needed, (size_t)A->n));
        KB_TRY(kb_alloc(&w->res_bar, 4));
    }
    KB_CUDA(cudaMemsetAsync(w->res_pk, 0, npk * sizeof(ulonglong2), c->stream));
    KB_CUDA(cudaMemsetAsync(w->res_needed, 0, (size_t)A->n, c->stream));
    KB_CUDA(cudaMemsetAsync(w->res_bar, 0, 4 * sizeof(unsigned), c->stream));
    KbPcgResArgs m{};
    m.row_ptr = A->row_ptr; m.col = A->col; m.vals = A->vals; m.n = (int)A->n; m.ntiles = A->ntiles; m.maxlen = L; m.tiles_per_cta = T; m.ghost_cap = ghost_cap; m.dbg = getenv("KB_RES_DEBUG") ? atoi(getenv("KB_RES_DEBUG")) : 0;
    m.x = w->x; m.r = w->r; m.p = w->p; m.inv = pc ? pc->inv_diag : nullptr; m.ctl = w->ctl;
    m.pk_p = w->res_pk; m.pk_a = w->res_pk + A->n; m.pk_b = w->res_pk + A->n + A->ntiles;
    m.needed = w->res_needed; m.bar = w->res_bar;
    void* args[] = {&m};
    {
        KbLaunch Lc(c, KB_K_SPMV);
        // a grid that cannot be co-resident right now (SMs taken by another context) is not an error: the CUDA-graph path runs
        if (cudaLaunchCooperativeKernel((const void*)kfn, dim3(G), dim3(KB_RES_THREADS), args, smem, c->stream) != cudaSuccess) {
            (void)cudaGetLastError();
            return KB_OK;
        }
    }
    unsigned verdict = 0;
    if (cudaMemcpyAsync(&verdict, w->res_bar + 2, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("pcg: resident kernel failed: %s", cudaGetErrorString(cudaGetLastError())); return KB_SOLVE_ERROR; }
    if (verdict == 1u) { kb_set_error("pcg: resident kernel timed out waiting for a packet"); return KB_SOLVE_ERROR; }
    *ran = verdict == 0u;
    return KB_OK;
}

static int pcg_iteration(kb_csr_s* A, const kb_pc_s* pc, KbPcgWs* w) {
    return A->dist && A->ctx->size > 1 ? pcg_launch_iteration<true>(A, pc, w) : pcg_launch_iteration<false>(A, pc, w);
}

extern "C" int kb_pcg_solve(kb_csr A, kb_pc pc, const double* b, double* x, double tol, uint64_t max_iters, int norm_type,
                            uint32_t flags, double* history, uint64_t hist_cap, uint64_t* hist_len, kb_stats* stats) {
    if (!A || !b || !x || !stats) { kb_set_error("kb_pcg_solve: null argument"); return KB_SOLVE_ERROR; }
    if (pc && pc->a != A) { kb_set_error("preconditioner was set up for a different operator"); return KB_SOLVE_ERROR; }
    if (pc && pc->kind != KB_PC_JACOBI && pc->kind != KB_PC_ILU0 && pc->kind != KB_PC_ASM) { kb_set_error("unsupported preconditioner"); return KB_UNSUPPORTED; }
    if (norm_type < 0 || norm_type > 3) { kb_set_error("bad norm_type"); return KB_SOLVE_ERROR; }
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    const bool dev = (flags & KB_FLAG_DEVICE_PTRS) != 0;
    const bool dist = A->dist && c->size > 1;
    if (!history) hist_cap = 0;
    KbPcgWs* w = nullptr;
    KB_TRY(pcg_ws_get(A, hist_cap, &w));
    memset(stats, 0, sizeof(*stats));
    if (hist_len) *hist_len = 0;
    if (A->n == 0 && !dist) { stats->converged = 0; return KB_OK; }
    const bool jacobi_like = !pc || pc->kind == KB_PC_JACOBI;
    const bool pipelined = (flags & KB_FLAG_PIPELINED) != 0;
    const bool single_red = (flags & KB_FLAG_SINGLE_REDUCTION) != 0 || pipelined;      // the pipelined variant shares the single-reduction set-up
    if (single_red && !jacobi_like) { kb_set_error("single-reduction / pipelined PCG supports Jacobi or no preconditioner"); return KB_UNSUPPORTED; }
    if (single_red && !w->u_sr) { KB_TRY(kb_alloc(&w->u_sr, w->nx + 2)); KB_TRY(kb_alloc(&w->s_sr, w->n + 2)); }
    if (pipelined && !w->m_pp) { KB_TRY(kb_alloc(&w->m_pp, w->nx + 2)); KB_TRY(kb_alloc(&w->n_pp, w->n + 2)); KB_TRY(kb_alloc(&w->q_pp, w->n + 2)); }

    KB_TRY(kb_upload_or_alias(c, b, w->b, w->n, dev));
    KB_TRY(kb_upload_or_alias(c, x, w->x, w->n, dev));
    // control block head
    KbCtl* h = w->h_ctl;
    memset(h, 0, offsetof(KbCtl, h));
    h->max_iters = max_iters; h->tol = tol; h->norm_type = norm_type; h->hist = w->hist; h->hist_cap = hist_cap;
    const bool own_hist = hist_cap == 0 && (flags & (KB_FLAG_HISTORY | KB_FLAG_MONITOR));
    if (own_hist) KB_TRY(kb_hist_prepare(A, flags, max_iters + 1, h));     // the operator's buffer (kb_get_history)
    else A->hist_len = 0;
    KbMonitor mon;
    if ((flags & KB_FLAG_MONITOR) && A->monitor) { mon.fn = A->monitor; mon.user = A->monitor_user; mon.d_hist = h->hist; mon.cap = h->hist_cap; }
    KB_CUDA(cudaMemcpyAsync(w->ctl, h, offsetof(KbCtl, h), cudaMemcpyHostToDevice, c->stream));

    const bool profile = (flags & KB_FLAG_PROFILE) != 0;
    const bool use_graph = !(flags & (KB_FLAG_NO_GRAPH | KB_FLAG_PROFILE));
    const bool was_prof = c->profiling;
    c->profiling = profile;
    int st = KB_OK;
    do {
        // r = b - A x  (pcg.rs:119-124)
        {
            KbSpmvEpi<PcgApFin, false, false> epi; epi.ctl = nullptr; epi.fin = kb_make_fin(c, PcgApFin{w->ctl}, false, nullptr, 0);
            if ((st = kb_launch_spmv<KbSpmvEpi<PcgApFin, false, false>, true>(A, w->x, w->r, w->b, nullptr, w->partials, w->pstride, epi, dist ? w->x : nullptr)) != KB_OK) break;
        }
        if (single_red) {     // u = D^-1 r ; w = A u ; gamma, delta, res0, first history entry
            st = dist ? pcg_sr_launch<true, true>(A, pc, w) : pcg_sr_launch<false, true>(A, pc, w);
            if (st != KB_OK) break;
            if (pipelined && (st = dist ? pcg_pipe_init<true>(A, pc, w) : pcg_pipe_init<false>(A, pc, w)) != KB_OK) break;   // m = M^-1 w ; n = A m ; z = q = 0
        } else if (!jacobi_like) {   // z = M^-1 r ; p = z ; rz ; norm
            if ((st = kb_pc_apply_dev(pc, w->r, w->z, nullptr, 0)) != KB_OK) break;
            PcgRzOp<PcgInitFin, true> op; op.n = (long long)w->n; op.partials = w->partials; op.pstride = w->pstride; op.ticket = c->ticket;
            op.r = w->r; op.z = w->z; op.p = w->p; op.ctl = w->ctl; op.fin = kb_make_fin(c, PcgInitFin{w->ctl}, dist, w->slots, 2);
            { KbLaunch L(c, KB_K_INIT); kb_tile_kernel<<<A->ntiles, KB_THREADS, 0, c->stream>>>(op); }
            if (cudaGetLastError() != cudaSuccess) { kb_set_error("pcg init launch failed"); st = KB_SOLVE_ERROR; break; }
            if (dist && (st = kb_finish_dist<PcgInitFin>(c, PcgInitFin{w->ctl}, w->ctl, w->slots, 2)) != KB_OK) break;
        } else {   // z, p, rz, res0, first history entry (pcg.rs:126-146)
            PcgInitOp<PcgInitFin> op; op.n = (long long)w->n; op.partials = w->partials; op.pstride = w->pstride; op.ticket = c->ticket;
            op.r = w->r; op.inv = pc ? pc->inv_diag : nullptr; op.z = w->z; op.p = w->p; op.ctl = w->ctl;
            op.fin = kb_make_fin(c, PcgInitFin{w->ctl}, dist, w->slots, 2);
            { KbLaunch L(c, KB_K_INIT); kb_tile_kernel<<<A->ntiles, KB_THREADS, 0, c->stream>>>(op); }
            if (cudaGetLastError() != cudaSuccess) { kb_set_error("pcg init launch failed"); st = KB_SOLVE_ERROR; break; }
            if (dist && (st = kb_finish_dist<PcgInitFin>(c, PcgInitFin{w->ctl}, w->ctl, w->slots, 2)) != KB_OK) break;
        }
        // Optional: the whole loop in one persistent cooperative kernel (KB_PCG_PERSISTENT=1).  Measured on B200 the
        // software grid barriers (3 per iteration over 296 CTAs, ~3 us each) cost more than the three kernel
        // boundaries of the graph path (C1: 22.3 vs 18.5 us per iteration), so it is opt-in.
        // the loop's SpMV expects p's halo already pushed by the kernel that produced p: push the initial p = z here
        if (dist && !single_red && kb_halo_fused_dev(A) && (st = kb_halo_begin(A, w->p)) != KB_OK) break;
        const int mega_env = getenv("KB_PCG_PERSISTENT") ? atoi(getenv("KB_PCG_PERSISTENT")) : 0;
        const bool mega_ok = !single_red && !dist && jacobi_like && A->kind == 2 && !A->prod && !profile && max_iters > 0;
        const bool mega = mega_ok && mega_env != 0 && !mon.fn;
        // Problems that fit on chip (<= 4 tiles per SM, rows <= 8 entries): the whole loop in one launch out of shared
        // memory and registers (KB_PCG_RESIDENT=0 keeps the CUDA-graph path).
        const int res_env = getenv("KB_PCG_RESIDENT") ? atoi(getenv("KB_PCG_RESIDENT")) : KB_PCG_RESIDENT_DEFAULT;
        bool resident = false;
        if (res_env != 0 && !mega && !single_red && !dist && jacobi_like && !mon.fn && max_iters > 0 && max_iters < (1ull << 30) /* packet tags are 3k+phase in 32 bits */ && !(flags & KB_FLAG_NO_GRAPH) && (st = pcg_resident(A, pc, w, &resident)) != KB_OK) break;
        if (resident) {
        } else if (mega) {
            if ((st = pcg_persistent(A, pc, w)) != KB_OK) break;
            unsigned berr = 0;
            if (cudaMemcpyAsync(&berr, w->mega_bar + 2, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("pcg: persistent kernel failed: %s", cudaGetErrorString(cudaGetLastError())); st = KB_SOLVE_ERROR; break; }
            if (berr) { kb_set_error("pcg: grid barrier timed out"); st = KB_SOLVE_ERROR; break; }
        } else {
        // iterations per graph replay: ~2 ms of work, so the per-replay host poll is amortised
        const int B = kb_batch_size(12.0 * (double)A->nnz + 108.0 * (double)A->n, 3);
        // graph key: the captured launches differ between the two variants
        st = kb_run_iterations(c, &w->gc, (kb_pc_serial(pc) + 1) ^ (pipelined ? 0x5049ull << 48 : single_red ? 0x5352ull << 48 : 0ull), B, max_iters, use_graph, w->ctl, h,
                               [&]() { return pipelined ? pcg_pipe_iteration(A, pc, w) : single_red ? pcg_sr_iteration(A, pc, w) : pcg_iteration(A, pc, w); }, &mon);
        }
        if (st != KB_OK) break;
        if (cudaMemcpyAsync(h, w->ctl, offsetof(KbCtl, h), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("pcg: readback failed"); st = KB_SOLVE_ERROR; break; }
        stats->iterations = h->iter; stats->final_residual = h->res; stats->converged = h->converged; stats->breakdown = 0;
        if (own_hist) A->hist_len = std::min<uint64_t>(h->hist_len, A->hist_cap);
        st = h->status;
        if (dist && kb_p2p_error(c)) { kb_set_error("%s: peer-memory collective timed out", "pcg"); st = KB_SOLVE_ERROR; break; }
        if (pc && kb_ilu0_error(const_cast<kb_pc_s*>(pc))) { kb_set_error("%s: a triangular-solve dependency wait timed out", "pcg"); st = KB_SOLVE_ERROR; break; }
        uint64_t hl = std::min<uint64_t>(h->hist_len, hist_cap);
        if (hist_len) *hist_len = h->hist_len;
        if (hl && cudaMemcpyAsync(history, w->hist, hl * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        if (st == KB_OK) {   // x is written only on Ok (pcg.rs:203,220 vs :171,:212)
            if (cudaMemcpyAsync(x, w->x, w->n * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) {
                kb_set_error("pcg: copy-out of x failed"); st = KB_SOLVE_ERROR; break;
            }
        } else if (st == KB_INDEFINITE_MATRIX) kb_set_error("indefinite matrix detected (p^T A p <= 0)");
        else if (st == KB_INDEFINITE_PC) kb_set_error("indefinite preconditioner detected (beta < 0)");
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("pcg: final sync failed"); st = KB_SOLVE_ERROR; }
    } while (0);
    c->profiling = was_prof;
    if (profile) kb_prof_collect(c);
    return st;
}
