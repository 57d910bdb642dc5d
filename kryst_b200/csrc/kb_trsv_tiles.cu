// kb_trsv_tiles.cu — block-wavefront triangular solves for grid-structured ILU(0) factors.
//
// The level-scheduled solves (kb_ilu0.cu) pay one inter-CTA hop (~1-2 us through L2 flags) per level: 766 hops
// per solve on the 256^3 7-point operator, 2047 on the 1024^2 5-point one.  When the factor's pattern is that of a
// lexicographically numbered box grid whose lower neighbours are (i-1,j,k), (i,j-1,k), (i,j,k-1) — detected from the
// CSR pattern alone, never assumed — the rows are grouped into BX x BY x BZ tiles:
//   * a tile depends only on the (up to 7) tiles at (I-a, J-b, K-c), a,b,c in {0,1}: the tile DAG is acyclic and its
//     depth is TX+TY+TZ-2 (94 for 256^3 with 8^3 tiles, 63 for 1024^2 with 32^2 tiles) instead of the row-level depth;
//   * inside a tile ONE CTA walks the BX+BY+BZ-2 internal wavefront steps with __syncthreads (tens of ns), the
//     tile's solution living in shared memory; values from finished neighbour tiles are read once from L2;
//   * tiles are processed by a persistent co-resident grid in block-level order; a per-tile flag (release/acquire)
//     is the only inter-CTA communication.  No sentinel fill, no value polling.
// Every row is still computed as s = rhs - sum_{stored entries in ascending column order} l_ij * y_j (then * 1/u_ii
// for U): identical operation order to the oracle, so the result is bit-identical whatever the schedule.
// Patterns that do not pass the check (27-point stencils, general matrices) keep the level-scheduled kernels.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kb_objects.h"

struct KbTileSolve {
    int nx = 0, ny = 0, nz = 0;          // detected grid
    int bx = 0, by = 0, bz = 0;          // tile shape
    int tx = 0, ty = 0, tz = 0;          // tiles per dimension
    int ntiles = 0, rows_per_thread = 0, grid[2] = {0, 0};
    int* order[2] = {nullptr, nullptr};  // tile ids in processing order (lower: ascending block level; upper: descending)
    int* flags = nullptr;                // [2][ntiles] completion flags, cleared before every apply
    unsigned* err = nullptr;             // borrowed: the preconditioner's error word (set when a spin times out)
};

// ---- pattern detection ---------------------------------------------------------------------------------------
// pass 1: smallest and largest |r - c| > 1 over all off-diagonal entries
__global__ void k_tiles_offsets(const int* __restrict__ rp, const int* __restrict__ col, int n, int* __restrict__ smin, int* __restrict__ smax) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int lo = 0x7fffffff, hi = 0;
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
        const int d = abs(r - col[p]);
        if (d > 1) { lo = min(lo, d); hi = max(hi, d); }
    }
    if (hi > 0) { atomicMin(smin, lo); atomicMax(smax, hi); }
}
// pass 2: every off-diagonal entry is a +-1 step along exactly one grid axis and does not wrap around a line / plane
__global__ void k_tiles_verify(const int* __restrict__ rp, const int* __restrict__ col, int n, int nx, int ny, int sx, int sy,
                               int* __restrict__ bad) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int i = r % nx, j = (r / nx) % ny;
    const bool three_d = sy > sx;
    bool ok = true;
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
        const int c = col[p];
        if (c == r) continue;
        const int d = abs(r - c);
        if (c < 0 || c >= n) ok = false;
        else if (d == 1) ok = ok && (c < r ? i >= 1 : i + 1 < nx);
        else if (d == sx) ok = ok && (!three_d || (c < r ? j >= 1 : j + 1 < ny));
        else if (three_d && d == sy) ok = ok && true;
        else ok = false;
    }
    if (!ok) atomicExch(bad, 1);
}

// ---- the solve -------------------------------------------------------------------------------------------------
struct KbTileArgs {
    const int* __restrict__ rp; const int* __restrict__ col; const double* __restrict__ lu; const int* __restrict__ dptr;
    const double* __restrict__ inv_diag;
    const double* __restrict__ rhs; double* out;
    int n, nx, ny, nz, bx, by, bz, tx, ty, tz, ntiles;
    const int* __restrict__ order; int* flags; unsigned* err;
    const KbCtl* skip_ctl; int skip_mask;
    int lag;                                                       // poll back-off in ns (KB_TILES_SLEEP)
};

#define KB_TILE_ROWS_MAX 1024

template <bool UPPER, int R>
__global__ void __launch_bounds__(KB_THREADS) kb_trsv_tiles(KbTileArgs a) {
    if (kb_skip(a.skip_ctl, a.skip_mask)) return;
    __shared__ double ytile[KB_TILE_ROWS_MAX];
    const int tid = threadIdx.x;
    const int bxy = a.bx * a.by, trows = bxy * a.bz;
    const int nlevels = a.bx + a.by + a.bz - 2;          // internal wavefront depth of a full tile
    const int sx = a.nx;                                 // column stride of a j-step; anything else that is not 1 is a k-step
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const int tile = a.order[t];
        const int TI = tile % a.tx, TJ = (tile / a.tx) % a.ty, TK = tile / (a.tx * a.ty);
        // ---- everything that does not depend on other tiles: this thread's rows and their stored entries
        int row[R], lvl[R], di[R][3];
        double rh[R], dg[R], cv[R][3], xv[R][3];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int q = m * KB_THREADS + tid;
            row[m] = -1; lvl[m] = -1; rh[m] = 0.0; dg[m] = 1.0;
#pragma unroll
            for (int e = 0; e < 3; ++e) { di[m][e] = -2; cv[m][e] = 0.0; xv[m][e] = 0.0; }     // -2: no entry
            if (q < trows) {
                const int li = q % a.bx, lj = (q / a.bx) % a.by, lk = q / bxy;
                const int gi = TI * a.bx + li, gj = TJ * a.by + lj, gk = TK * a.bz + lk;
                const int r = gi + a.nx * (gj + a.ny * gk);
                if (gi < a.nx && gj < a.ny && gk < a.nz && r < a.n) {
                    row[m] = r;
                    lvl[m] = li + lj + lk;
                    rh[m] = a.rhs[r];
                    const int pd = a.dptr[r];
                    const int p0 = UPPER ? pd + 1 : a.rp[r];
                    const int p1 = UPPER ? a.rp[r + 1] : pd;
                    if (UPPER) dg[m] = a.inv_diag[r];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        if (p0 + e < p1) {
                            const int c = a.col[p0 + e];
                            cv[m][e] = a.lu[p0 + e];
                            const int d = UPPER ? c - r : r - c;
                            // inside the tile -> shared-memory slot, otherwise -1 (value fetched after the flag wait)
                            int slot = -1;
                            if (d == 1) { if (UPPER ? li + 1 < a.bx : li >= 1) slot = UPPER ? q + 1 : q - 1; }
                            else if (d == sx) { if (UPPER ? lj + 1 < a.by : lj >= 1) slot = UPPER ? q + a.bx : q - a.bx; }
                            else { if (UPPER ? lk + 1 < a.bz : lk >= 1) slot = UPPER ? q + bxy : q - bxy; }
                            di[m][e] = slot;
                            if (slot < 0) xv[m][e] = __longlong_as_double((long long)c);      // park the column id until the wait is over
                        }
                    }
                }
            }
        }
        // ---- wait for the (up to 7) predecessor tiles
        if (tid >= 1 && tid < 8) {
            const int da = tid & 1, db = (tid >> 1) & 1, dc = (tid >> 2) & 1;
            const int PI = UPPER ? TI + da : TI - da, PJ = UPPER ? TJ + db : TJ - db, PK = UPPER ? TK + dc : TK - dc;
            if (PI >= 0 && PI < a.tx && PJ >= 0 && PJ < a.ty && PK >= 0 && PK < a.tz) {
                const int* f = a.flags + (PI + a.tx * (PJ + a.ty * PK));
                unsigned spins = 0;
                int v;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                    if (v == 0 && a.lag > 0) __nanosleep(a.lag);   // polite polling (KB_TILES_SLEEP ns): fewer strong loads competing with the working tiles
                    if (v == 0 && (++spins & 1023u) == 0u) {      // bounded: give up (and let everybody give up) instead of hanging
                        if (spins > (1u << 26)) atomicExch(a.err, 1u);
                        if (*reinterpret_cast<volatile unsigned*>(a.err)) break;
                    }
                } while (v == 0);
            }
        }
        __syncthreads();
        // ---- values owned by finished tiles (L2 reads; never cached in L1)
#pragma unroll
        for (int m = 0; m < R; ++m)
#pragma unroll
            for (int e = 0; e < 3; ++e)
                if (di[m][e] == -1) xv[m][e] = __ldcg(a.out + (int)__double_as_longlong(xv[m][e]));
        // ---- internal wavefront
        for (int s0 = 0; s0 < nlevels; ++s0) {
            const int step = UPPER ? nlevels - 1 - s0 : s0;
#pragma unroll
            for (int m = 0; m < R; ++m) {
                if (lvl[m] == step) {
                    // the three operand loads are unconditional (clamped slot) and independent, so they overlap; an
                    // absent entry contributes s - 0.0*0.0, which leaves every finite s (and -0.0) unchanged
                    double v[3];
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const double y = ytile[max(di[m][e], 0)];
                        v[e] = di[m][e] >= 0 ? y : xv[m][e];
                    }
                    double s = rh[m];
#pragma unroll
                    for (int e = 0; e < 3; ++e) s = s - cv[m][e] * v[e];
                    if (UPPER) s = s * dg[m];
                    ytile[m * KB_THREADS + tid] = s;
                    a.out[row[m]] = s;
                }
            }
            __syncthreads();
        }
        // ---- publish: the barrier above ordered every thread's stores before this release
        if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.flags + tile), "r"(1) : "memory");
    }
}

// ---- host ------------------------------------------------------------------------------------------------------
void kb_tiles_grid(const KbTileSolve* t, int* nx, int* ny, int* nz) { *nx = t->nx; *ny = t->ny; *nz = t->nz; }
void kb_tiles_free(KbTileSolve* t) {
    if (!t) return;
    KB_FREE(t->order[0]); KB_FREE(t->order[1]); KB_FREE(t->flags);
    delete t;
}

template <bool UPPER, int R>
static int tiles_occupancy(int* occ) {
    KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kb_trsv_tiles<UPPER, R>, KB_THREADS, 0));
    return KB_OK;
}

// *out stays nullptr (and KB_OK is returned) when the pattern is not a 5-/7-point box grid.
int kb_tiles_build(kb_pc_s* pc, unsigned* d_err, KbTileSolve** out) {
    *out = nullptr;
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    const int n = (int)A->n;
    if (n < 64) return KB_OK;
    const unsigned nb = (unsigned)((n + 255) / 256);
    int* d_w = nullptr;     // [0] smin, [1] smax, [2] bad
    KB_TRY(kb_alloc(&d_w, 4));
    int h_w[3] = {0x7fffffff, 0, 0};
    KB_CUDA(cudaMemcpyAsync(d_w, h_w, sizeof(h_w), cudaMemcpyHostToDevice, c->stream));
    { KbLaunch L(c, KB_K_OTHER); k_tiles_offsets<<<nb, 256, 0, c->stream>>>(pc->l_rp, pc->l_col, n, d_w, d_w + 1); }
    KB_CUDA(cudaMemcpyAsync(h_w, d_w, sizeof(h_w), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    const int s1 = h_w[0], s2 = h_w[1];
    if (s2 == 0 || s1 < 2 || (s2 != s1 && s2 % s1 != 0)) { cudaFree(d_w); return KB_OK; }
    const int nx = s1, sy = s2 == s1 ? s1 : s2;                      // sy == nx: two-dimensional
    const int ny = s2 == s1 ? (n + nx - 1) / nx : s2 / s1;
    const int nz = s2 == s1 ? 1 : (n + s2 - 1) / s2;
    { KbLaunch L(c, KB_K_OTHER); k_tiles_verify<<<nb, 256, 0, c->stream>>>(pc->l_rp, pc->l_col, n, nx, ny, nx, sy, d_w + 2); }
    KB_CUDA(cudaMemcpyAsync(h_w, d_w, sizeof(h_w), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_w);
    if (h_w[2]) return KB_OK;

    KbTileSolve* t = new KbTileSolve;
    t->nx = nx; t->ny = ny; t->nz = nz;
    if (nz == 1) { t->bx = 32; t->by = 32; t->bz = 1; }
    else { t->bx = 8; t->by = 8; t->bz = 8; }
    if (const char* e = getenv("KB_TILES_SHAPE")) {      // tuning knob: "bx,by,bz"
        int b[3] = {0, 0, 0};
        if (sscanf(e, "%d,%d,%d", &b[0], &b[1], &b[2]) == 3 && b[0] > 0 && b[1] > 0 && b[2] > 0 && (long long)b[0] * b[1] * b[2] <= KB_TILE_ROWS_MAX) {
            t->bx = b[0]; t->by = b[1]; t->bz = nz == 1 ? 1 : b[2];
        }
    }
    { const int rows = t->bx * t->by * t->bz; t->rows_per_thread = rows <= KB_THREADS ? 1 : rows <= 2 * KB_THREADS ? 2 : 4; }
    t->tx = (nx + t->bx - 1) / t->bx; t->ty = (ny + t->by - 1) / t->by; t->tz = (nz + t->bz - 1) / t->bz;
    const long long nt = (long long)t->tx * t->ty * t->tz;
    if (nt > (1 << 24)) { delete t; return KB_OK; }
    t->ntiles = (int)nt;
    // processing order: stable sort by block level I+J+K (ascending for L, descending for U)
    std::vector<int> ord((size_t)nt), lev((size_t)nt);
    for (int k = 0, id = 0; k < t->tz; ++k)
        for (int j = 0; j < t->ty; ++j)
            for (int i = 0; i < t->tx; ++i, ++id) { ord[id] = id; lev[id] = i + j + k; }
    std::stable_sort(ord.begin(), ord.end(), [&](int p, int q) { return lev[p] < lev[q]; });
    int st = KB_OK;
    do {
        if ((st = kb_alloc(&t->order[0], (size_t)nt)) != KB_OK || (st = kb_alloc(&t->order[1], (size_t)nt)) != KB_OK ||
            (st = kb_alloc(&t->flags, 2 * (size_t)nt)) != KB_OK) break;
        t->err = d_err;
        if (cudaMemcpyAsync(t->order[0], ord.data(), (size_t)nt * sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        cudaStreamSynchronize(c->stream);
        std::reverse(ord.begin(), ord.end());
        if (cudaMemcpyAsync(t->order[1], ord.data(), (size_t)nt * sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        // (Measured on B200, 256^3, per solve: release/acquire flag hand-off 0.71 ms; tagged-packet hand-offs - polled inside the
        //  compute step 1.92 ms, by a communication warp 1.21 ms, at tile granularity 1.09-1.29 ms: ~200 polling threads per waiting
        //  tile against 7 here.  The packet variants were removed; the pencil march of kb_trsv_lean.cu is what replaced this kernel.)
        // the whole grid must be co-resident: a tile may wait on a tile owned by any other CTA
        int occ[2] = {1, 1};
        if (t->rows_per_thread == 4) { if ((st = tiles_occupancy<false, 4>(&occ[0])) != KB_OK || (st = tiles_occupancy<true, 4>(&occ[1])) != KB_OK) break; }
        else if (t->rows_per_thread == 2) { if ((st = tiles_occupancy<false, 2>(&occ[0])) != KB_OK || (st = tiles_occupancy<true, 2>(&occ[1])) != KB_OK) break; }
        else { if ((st = tiles_occupancy<false, 1>(&occ[0])) != KB_OK || (st = tiles_occupancy<true, 1>(&occ[1])) != KB_OK) break; }
        for (int u = 0; u < 2; ++u) {
            int g = std::max(1, occ[u]) * c->sm_count;
            if (getenv("KB_TILES_GRID")) g = std::min(g, std::max(1, atoi(getenv("KB_TILES_GRID"))));
            t->grid[u] = std::min(g, t->ntiles);
        }
    } while (0);
    if (st != KB_OK) { kb_set_error("ilu0: tile schedule setup failed"); kb_tiles_free(t); return st; }
    *out = t;
    return KB_OK;
}

int kb_tiles_apply(kb_pc_s* pc, KbTileSolve* t, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask) {
    kb_ctx_s* c = pc->a->ctx;
    KbTileArgs a{};
    a.rp = pc->l_rp; a.col = pc->l_col; a.lu = pc->lu; a.dptr = pc->diag_ptr; a.inv_diag = pc->inv_diag;
    a.n = (int)pc->a->n; a.nx = t->nx; a.ny = t->ny; a.nz = t->nz; a.bx = t->bx; a.by = t->by; a.bz = t->bz;
    a.tx = t->tx; a.ty = t->ty; a.tz = t->tz; a.ntiles = t->ntiles; a.err = t->err; a.skip_ctl = skip_ctl; a.skip_mask = skip_mask;
    a.lag = getenv("KB_TILES_SLEEP") ? atoi(getenv("KB_TILES_SLEEP")) : 0;      // (the flag kernel reuses the field as its poll back-off)
    KB_CUDA(cudaMemsetAsync(t->flags, 0, 2 * (size_t)t->ntiles * sizeof(int), c->stream));
    {
        a.order = t->order[0]; a.flags = t->flags; a.rhs = d_r; a.out = pc->tmp;
        KbLaunch L(c, KB_K_TRSV);
        if (t->rows_per_thread == 4) kb_trsv_tiles<false, 4><<<t->grid[0], KB_THREADS, 0, c->stream>>>(a);
        else if (t->rows_per_thread == 2) kb_trsv_tiles<false, 2><<<t->grid[0], KB_THREADS, 0, c->stream>>>(a);
        else kb_trsv_tiles<false, 1><<<t->grid[0], KB_THREADS, 0, c->stream>>>(a);
    }
    {
        a.order = t->order[1]; a.flags = t->flags + t->ntiles; a.rhs = pc->tmp; a.out = d_z;
        KbLaunch L(c, KB_K_TRSV);
        if (t->rows_per_thread == 4) kb_trsv_tiles<true, 4><<<t->grid[1], KB_THREADS, 0, c->stream>>>(a);
        else if (t->rows_per_thread == 2) kb_trsv_tiles<true, 2><<<t->grid[1], KB_THREADS, 0, c->stream>>>(a);
        else kb_trsv_tiles<true, 1><<<t->grid[1], KB_THREADS, 0, c->stream>>>(a);
    }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

