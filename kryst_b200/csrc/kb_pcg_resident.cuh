// kb_pcg_resident.cuh — PCG (pcg.rs:148-218) for problems that fit on chip: ONE cooperative launch per solve, one CTA
// per SM, every CTA keeps its rows of the operator in shared memory and its entries of x, r, p, D^-1 in registers
// for the whole solve.  Nothing streams from HBM inside the loop; what is left per iteration is two L2 round trips:
// the tile sums of p.Ap, and the tile sums of r.z and the norm together with the entries of z that other CTAs need
// (a CTA updates its ghost copies of p itself, p_g = z_g + beta p_g: same operands, same bits as the owner's update).
// Everything travels as 16-byte tagged packets ({lo32|tag, hi32|tag}: the data is its own flag, tag = 3*iteration +
// phase), so there is no grid barrier, no flag and no fence in the loop: a consumer polls the packet it needs.
// Every CTA sums ALL tile sums itself in the canonical order (level 2 of the reduction tree, kb_internal.cuh) and
// runs the scalar recurrences redundantly - same bits everywhere, so all CTAs leave the loop in the same iteration.
// Arithmetic (row sums in stored order, the 512-row tile sums, level 2, the recurrences) is that of the three-kernel
// path (kb_spmv_bulk + PcgUpdateOp + PcgXpayOp): results are bit-identical to it and to the oracle.
//
// Layout: 1024 threads = 4 teams of 256; a team owns one canonical 512-row tile (a CTA owns <= 4 consecutive tiles),
// thread l of a team owns rows 2l and 2l+1 of the tile - the pairing of the canonical tile sum.  The tile's entries
// sit in shared memory slot-major (slot s of rows 2l, 2l+1 is one 16-byte word: conflict-free), columns as indices
// into the CTA's p window in shared memory: its own rows first, then one slot per reference to a column of another
// CTA (filled from the owners' packets at the start of every iteration, one poll per thread on C1).
// Eligibility (host): single GPU, Jacobi or no preconditioner, rows <= 8 entries, tiles <= 4 * SMs (C1: 512 tiles),
// operator rows + p window within 227 KB; (kernel prologue): the CTA's ghost references fit the room left in shared
// memory - otherwise the kernel backs out before touching anything and the solver takes the CUDA-graph path.
#pragma once
#include "kb_internal.cuh"

#define KB_RES_TEAMS 4
#define KB_RES_THREADS (KB_RES_TEAMS * KB_THREADS)
#define KB_RES_MAXLEN 8
#define KB_RES_SPIN (1u << 21)

struct KbPcgResArgs {
    const int* __restrict__ row_ptr; const int* __restrict__ col; const double* __restrict__ vals;
    int n, ntiles, maxlen;
    int tiles_per_cta;       // max tiles of one CTA (<= KB_RES_TEAMS): sizes the operator rows in shared memory
    int ghost_cap;           // ghost references one CTA can hold
    int dbg;                 // timing ablations only (KB_RES_DEBUG; results are then meaningless): 1/2/4 do not wait for p.Ap sums /
                             // r.z sums / ghosts, 8 back off between failed polls, 16 fixed iteration count
    double* x; double* r; const double* p; const double* inv;
    KbCtl* ctl;
    ulonglong2* pk_p;        // [n]        tagged entries of the initial p, then of z (only rows that are ghost columns of another CTA)
    ulonglong2* pk_a;        // [ntiles]   tagged tile sums of p.Ap
    ulonglong2* pk_b;        // [2*ntiles] tagged tile sums of r.z and of the norm
    unsigned char* needed;   // [n]        row is a ghost column of another CTA
    unsigned* bar;           // [0] arrivals, [1] generation, [2] 1 = a poll timed out, 2 = not eligible (nothing was modified)
};

// operator rows (16 + 8 bytes per slot pair), own p window, reduction scratch; + 12 bytes per ghost reference
static inline size_t kb_res_smem_fixed(int maxlen, int tiles_per_cta) {
    return (size_t)tiles_per_cta * maxlen * KB_THREADS * (16 + 8) + (size_t)tiles_per_cta * KB_TILE * 8 + KB_RES_TEAMS * 16 * 8 + 8 * 8 + 16;
}
static inline size_t kb_res_smem_bytes(int maxlen, int tiles_per_cta, int ghost_cap) {
    return kb_res_smem_fixed(maxlen, tiles_per_cta) + (size_t)ghost_cap * 12;
}

#ifdef __CUDACC__
__device__ __forceinline__ void kb_res_store(ulonglong2* p, double v, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ __forceinline__ bool kb_res_load(const ulonglong2* p, unsigned tag, double* v) {
    unsigned long long x, y;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
    *v = __longlong_as_double((long long)((x & 0xffffffffull) | (y << 32)));
    return (unsigned)(x >> 32) == tag && (unsigned)(y >> 32) == tag;
}
// bounded poll; a timeout raises the CTA's error flag (checked by everybody after the next __syncthreads)
__device__ __forceinline__ double kb_res_poll(const ulonglong2* p, unsigned tag, int* s_err, int nowait = 0, int backoff = 0) {
    double v;
    unsigned spins = 0;
    while (!kb_res_load(p, tag, &v)) {
        if (nowait) break;
        if (backoff) __nanosleep(100);
        if (++spins > KB_RES_SPIN || *reinterpret_cast<volatile int*>(s_err) != 0) { *s_err = 1; break; }
    }
    return v;
}
__device__ __forceinline__ void kb_team_sync(int team) { asm volatile("bar.sync %0, 256;" ::"r"(team + 1) : "memory"); }

// canonical tile sum among the 256 threads of a team (named barrier team+1); result valid in lane 0 of the team
template <int NRED>
__device__ __forceinline__ void kb_team_reduce(double (&v)[NRED], double* sm /* [NRED*8] of this team */, double (&out)[NRED], int team, int l) {
    const int lane = l & 31, w = l >> 5;
#pragma unroll
    for (int r = 0; r < NRED; ++r) {
        double t = kb_warp_butterfly(v[r]);
        if (lane == 0) sm[r * 8 + w] = t;
    }
    kb_team_sync(team);
    if (l == 0) {
#pragma unroll
        for (int r = 0; r < NRED; ++r) {
            double s = sm[r * 8];
#pragma unroll
            for (int k = 1; k < 8; ++k) s = s + sm[r * 8 + k];
            out[r] = s;
        }
    }
    kb_team_sync(team);
}
// level 2 over the tagged tile sums (one team; result valid in lane 0 of the team).  P <= 768: a lane owns the sums
// l, l+256, l+512; their loads are issued together and repeated until every tag matches.
__device__ __forceinline__ double kb_team_level2(const ulonglong2* pk, int P, unsigned tag, double* sm, int team, int l, int* s_err, int nowait = 0, int backoff = 0) {
    const bool h0 = l < P, h1 = l + KB_THREADS < P, h2 = l + 2 * KB_THREADS < P;
    bool d0 = !h0, d1 = !h1, d2 = !h2;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    unsigned spins = 0;
    while (!(d0 && d1 && d2)) {
        if (!d0) d0 = kb_res_load(pk + l, tag, &v0);
        if (!d1) d1 = kb_res_load(pk + l + KB_THREADS, tag, &v1);
        if (!d2) d2 = kb_res_load(pk + l + 2 * KB_THREADS, tag, &v2);
        if (nowait) break;
        if (d0 && d1 && d2) break;
        if (backoff) __nanosleep(100);
        if (++spins > KB_RES_SPIN || *reinterpret_cast<volatile int*>(s_err) != 0) { *s_err = 1; break; }
    }
    double acc = 0.0;
    if (h0) acc = acc + v0;
    if (h1) acc = acc + v1;
    if (h2) acc = acc + v2;
    double v[1] = {acc}, out[1] = {0.0};
    kb_team_reduce<1>(v, sm, out, team, l);
    return out[0];
}
__global__ void __launch_bounds__(KB_RES_THREADS, 1) kb_pcg_resident(KbPcgResArgs m) {
    extern __shared__ __align__(16) unsigned char kb_res_raw[];
    const int L = m.maxlen, T = m.tiles_per_cta, OWN = T * KB_TILE;
    const int dbg = m.dbg, boff = dbg & 8;
    double2* vals2 = reinterpret_cast<double2*>(kb_res_raw);                                  // [T][L][256]
    int2* cols2 = reinterpret_cast<int2*>(vals2 + (size_t)T * L * KB_THREADS);                // [T][L][256]
    double* ps = reinterpret_cast<double*>(cols2 + (size_t)T * L * KB_THREADS);               // [T*512 own | ghost_cap]
    double* red = ps + OWN + m.ghost_cap;                                                     // [TEAMS][16]
    double* scal = red + KB_RES_TEAMS * 16;                                                   // [8]
    int* s_err = reinterpret_cast<int*>(scal + 8);                                            // [0] error, [1] ghost references
    int* gcol_s = s_err + 4;                                                                  // [ghost_cap] global column of each ghost slot
    __shared__ unsigned s_gen;

    KbCtl* c = m.ctl;
    if (c->done != 0) return;                 // max_iters == 0 (PcgInitFin); uniform over the grid
    const int tid = threadIdx.x, team = tid >> 8, l = tid & 255;
    const int G = (int)gridDim.x;
    const int tb0 = (int)(((long long)blockIdx.x * m.ntiles) / G), tb1 = (int)(((long long)(blockIdx.x + 1) * m.ntiles) / G);
    const int R0 = tb0 * KB_TILE, R1 = min(m.n, tb1 * KB_TILE);
    const int tile = tb0 + team;
    const bool active = tile < tb1;
    const int rowA = tile * KB_TILE + 2 * l, rowB = rowA + 1;
    const bool hasA = active && rowA < m.n, hasB = active && rowB < m.n;
    double* tred = red + team * 16;
    double2* tv = vals2 + (size_t)team * L * KB_THREADS;
    int2* tc = cols2 + (size_t)team * L * KB_THREADS;
    if (tid == 0) { s_err[0] = 0; s_err[1] = 0; }
    __syncthreads();

    // ---- prologue: this thread's two rows into shared memory, its vector entries into registers
    int lenA = 0, lenB = 0;
    {
        const int a0 = hasA ? m.row_ptr[rowA] : 0, b0 = hasB ? m.row_ptr[rowB] : 0;
        lenA = hasA ? m.row_ptr[rowA + 1] - a0 : 0;
        lenB = hasB ? m.row_ptr[rowB + 1] - b0 : 0;
        for (int s = 0; s < L; ++s) {
            double2 v = make_double2(0.0, 0.0);
            int2 cc = make_int2(0, 0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool on = h == 0 ? s < lenA : s < lenB;
                if (on) {
                    const int q = (h == 0 ? a0 : b0) + s;
                    const int cg = m.col[q];
                    int code;
                    if (cg >= R0 && cg < R1) code = cg - R0;
                    else {
                        const int k = atomicAdd(&s_err[1], 1);            // one slot per reference (duplicates only cost a poll)
                        if (k < m.ghost_cap) gcol_s[k] = cg;
                        code = OWN + k;
                        m.needed[cg] = 1;
                    }
                    if (h == 0) { v.x = m.vals[q]; cc.x = code; } else { v.y = m.vals[q]; cc.y = code; }
                }
            }
            if (active) { tv[s * KB_THREADS + l] = v; tc[s * KB_THREADS + l] = cc; }
        }
    }
    double xA = 0.0, xB = 0.0, rA = 0.0, rB = 0.0, pA = 0.0, pB = 0.0, dA = 1.0, dB = 1.0;
    const bool jac = m.inv != nullptr;
    if (hasA) { xA = m.x[rowA]; rA = m.r[rowA]; pA = m.p[rowA]; if (jac) dA = m.inv[rowA]; }
    if (hasB) { xB = m.x[rowB]; rB = m.r[rowB]; pB = m.p[rowB]; if (jac) dB = m.inv[rowB]; }
    if (active) { ps[team * KB_TILE + 2 * l] = pA; ps[team * KB_TILE + 2 * l + 1] = pB; }
    double rz = c->rz;
    const double res0 = c->res0, tol = c->tol;
    const unsigned long long max_iters = c->max_iters;
    const int nt = c->norm_type;
    unsigned long long iter = c->iter, hist_len = c->hist_len;
    const unsigned long long hist_cap = c->hist_cap;
    double* hist = c->hist;

    // one grid barrier: the `needed` marks and the eligibility verdict are complete
    __syncthreads();
    const int ng = s_err[1];
    if (tid == 0) {
        if (ng > m.ghost_cap) atomicExch(m.bar + 2, 2u);
        __threadfence();
        s_gen = *reinterpret_cast<volatile unsigned*>(m.bar + 1);
        const unsigned t = atomicAdd(m.bar, 1u);
        if (t == gridDim.x - 1u) { m.bar[0] = 0u; __threadfence(); atomicExch(m.bar + 1, s_gen + 1u); }
        else {
            unsigned spins = 0;
            while (*reinterpret_cast<volatile unsigned*>(m.bar + 1) == s_gen) { if (++spins > (1u << 26)) { atomicExch(m.bar + 2, 1u); break; } }
        }
        __threadfence();
    }
    __syncthreads();
    if (*reinterpret_cast<volatile unsigned*>(m.bar + 2) != 0u) return;      // backed out (or barrier timeout): nothing modified
    const bool needA = hasA && reinterpret_cast<volatile unsigned char*>(m.needed)[rowA] != 0;
    const bool needB = hasB && reinterpret_cast<volatile unsigned char*>(m.needed)[rowB] != 0;
    if (needA) kb_res_store(m.pk_p + rowA, pA, 3u);
    if (needB) kb_res_store(m.pk_p + rowB, pB, 3u);

    int status = KB_OK, converged = 0;
    double res = c->res, alpha = 0.0, beta = 0.0, pAp = 0.0;
    bool err = false;
    for (int q = tid; q < ng; q += KB_RES_THREADS) ps[OWN + q] = kb_res_poll(m.pk_p + gcol_s[q], 3u, s_err, 0, boff);   // ghost copies of the initial p
    __syncthreads();
    for (unsigned k = 1;; ++k) {
        const unsigned tagA = 3u * k + 1u, tagB = 3u * k + 2u;
        // ---- ap = A p out of shared memory (own rows and ghost copies), tile sum of p.Ap
        double apA = 0.0, apB = 0.0;
        if (active) {
#pragma unroll
            for (int s = 0; s < KB_RES_MAXLEN; ++s) {
                if (s < L) {
                    const double2 v = tv[s * KB_THREADS + l];
                    const int2 cc = tc[s * KB_THREADS + l];
                    if (s < lenA) apA = apA + v.x * ps[cc.x];
                    if (s < lenB) apB = apB + v.y * ps[cc.y];
                }
            }
            double e[1], o[1];
            e[0] = (hasA ? pA * apA : 0.0) + (hasB ? pB * apB : 0.0);
            kb_team_reduce<1>(e, tred, o, team, l);
            if (l == 0) kb_res_store(m.pk_a + tile, o[0], tagA);
        }
        if (team == 0) {
            const double s = kb_team_level2(m.pk_a, m.ntiles, tagA, tred, 0, l, s_err, dbg & 1, boff);
            if (l == 0) scal[0] = s;
        }
        __syncthreads();
        if (*s_err) { err = true; break; }
        pAp = scal[0];
        if (pAp <= 0.0 && !(dbg & 16)) { status = KB_INDEFINITE_MATRIX; iter = iter + 1; converged = 0; break; }      // pcg.rs:161-173
        alpha = rz / pAp;
        // ---- x += alpha p ; r -= alpha ap ; z = D^-1 r ; tile sums of r.z and the norm
        double zA = 0.0, zB = 0.0;
        if (active) {
            double e[2] = {0.0, 0.0}, o[2];
            double e0A = 0.0, e0B = 0.0, e1A = 0.0, e1B = 0.0;
            if (hasA) {
                xA = xA + alpha * pA; rA = rA - alpha * apA; zA = jac ? dA * rA : rA;
                e0A = rA * zA;
                e1A = nt == KB_NORM_PRECONDITIONED ? zA * zA : nt == KB_NORM_UNPRECONDITIONED ? rA * rA : 0.0;
            }
            if (hasB) {
                xB = xB + alpha * pB; rB = rB - alpha * apB; zB = jac ? dB * rB : rB;
                e0B = rB * zB;
                e1B = nt == KB_NORM_PRECONDITIONED ? zB * zB : nt == KB_NORM_UNPRECONDITIONED ? rB * rB : 0.0;
            }
            if (needA) kb_res_store(m.pk_p + rowA, zA, tagB);        // other CTAs rebuild their copy of p from z and beta
            if (needB) kb_res_store(m.pk_p + rowB, zB, tagB);
            e[0] = e0A + e0B; e[1] = e1A + e1B;
            kb_team_reduce<2>(e, tred, o, team, l);
            if (l == 0) { kb_res_store(m.pk_b + tile, o[0], tagB); kb_res_store(m.pk_b + m.ntiles + tile, o[1], tagB); }
        }
        if (team < 2) {
            const double s = kb_team_level2(m.pk_b + (size_t)team * m.ntiles, m.ntiles, tagB, tred, team, l, s_err, dbg & 2, boff);
            if (l == 0) scal[1 + team] = s;
        }
        __syncthreads();
        if (*s_err) { err = true; break; }
        // ---- PcgUpdateFin (pcg.rs:188-218), redundantly in every thread
        const double rz_new = scal[1];
        res = (nt == KB_NORM_PRECONDITIONED || nt == KB_NORM_UNPRECONDITIONED) ? sqrt(scal[2]) : (nt == KB_NORM_NATURAL ? sqrt(fabs(rz_new)) : 0.0);
        iter = iter + 1;
        if (blockIdx.x == 0 && tid == 0 && hist_len < hist_cap) hist[hist_len] = res;
        hist_len += 1;
        const double rel = res / res0;
        if ((rel <= tol && !(dbg & 16)) || iter >= max_iters) { converged = 1; break; }
        beta = rz_new / rz;
        if (beta < 0.0 && !(dbg & 16)) { status = KB_INDEFINITE_PC; converged = 0; break; }
        rz = rz_new;
        // ---- p = z + beta p : own rows, and the ghost copies from the owners' z
        if (active) {
            if (hasA) pA = zA + beta * pA;
            if (hasB) pB = zB + beta * pB;
            ps[team * KB_TILE + 2 * l] = pA; ps[team * KB_TILE + 2 * l + 1] = pB;
        }
        for (int q = tid; q < ng; q += KB_RES_THREADS) {
            const double zg = kb_res_poll(m.pk_p + gcol_s[q], tagB, s_err, dbg & 4, boff);
            ps[OWN + q] = zg + beta * ps[OWN + q];
        }
        __syncthreads();
    }
    if (err) { if (tid == 0) atomicExch(m.bar + 2, 1u); return; }
    if (hasA) { m.x[rowA] = xA; m.r[rowA] = rA; }
    if (hasB) { m.x[rowB] = xB; m.r[rowB] = rB; }
    if (blockIdx.x == 0 && tid == 0) {
        c->iter = iter; c->res = res; c->converged = converged; c->status = status; c->hist_len = hist_len;
        c->rz = rz; c->alpha = alpha; c->beta = beta; c->pAp = pAp; c->done = 1;
    }
}
#endif
