// kb_capi.cu — context, operator upload, MatVec, InnerProduct, Jacobi, profiling: the part of the
// C ABI (include/kryst_b200.h) that is not a Krylov driver.
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include "kb_objects.h"
#include "kb_epilogue.cuh"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void kb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* kb_last_error(void) { return g_err; }
extern "C" int kb_abi_version(void) { return KB_ABI_VERSION; }
#include <atomic>
uint64_t kb_next_serial() { static std::atomic<uint64_t> g{0}; return ++g; }

// ---------------------------------------------------------------------------------------------
// launch bookkeeping / profiling
// ---------------------------------------------------------------------------------------------
KbLaunch::KbLaunch(kb_ctx_s* ctx, int k) : c(ctx), cls(k) {
    if (c->capturing) c->captured_launches++;
    else c->launches++;
    if (c->profiling && !c->capturing) {
        if (c->event_pool.size() < 2) {
            for (int i = 0; i < 64; ++i) { cudaEvent_t e; cudaEventCreate(&e); c->event_pool.push_back(e); }
        }
        a = c->event_pool.back(); c->event_pool.pop_back();
        b = c->event_pool.back(); c->event_pool.pop_back();
        cudaEventRecord(a, c->stream);
    }
}
KbLaunch::~KbLaunch() {
    if (a) {
        cudaEventRecord(b, c->stream);
        c->prof_events.push_back({cls, a, b});
    }
}
int kb_prof_collect(kb_ctx_s* c) {
    if (c->prof_events.empty()) return KB_OK;
    KB_CUDA(cudaStreamSynchronize(c->stream));
    for (auto& e : c->prof_events) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        c->prof_acc.launches[e.cls] += 1;
        c->prof_acc.ms[e.cls] += ms;
        c->event_pool.push_back(e.a);
        c->event_pool.push_back(e.b);
    }
    c->prof_events.clear();
    return KB_OK;
}
static const char* k_class_names[KB_PROF_CLASSES] = {"spmv", "pcg_update", "xpay", "init", "bicgstab", "gs_dot",
                                                     "gs_update", "trsv", "small", "halo", "allreduce", "other"};
extern "C" const char* kb_profile_class_name(int cls) { return (cls >= 0 && cls < KB_PROF_CLASSES) ? k_class_names[cls] : "?"; }
extern "C" int kb_profile_reset(kb_ctx ctx) {
    KB_TRY(kb_prof_collect(ctx));
    memset(&ctx->prof_acc, 0, sizeof(ctx->prof_acc));
    return KB_OK;
}
extern "C" int kb_profile_get(kb_ctx ctx, kb_profile* out) {
    KB_TRY(kb_prof_collect(ctx));
    *out = ctx->prof_acc;
    return KB_OK;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int kb_ctx_create(int device, kb_ctx* out) {
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        kb_set_error("no usable CUDA device (%s); kryst_b200 has no CPU fallback", cudaGetErrorString(e));
        return KB_SOLVE_ERROR;
    }
    if (device < 0 || device >= ndev) { kb_set_error("device %d out of range (have %d)", device, ndev); return KB_SOLVE_ERROR; }
    KB_CUDA(cudaSetDevice(device));
    kb_ctx_s* c = new kb_ctx_s;
    c->device = device;
    cudaDeviceProp prop;
    KB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    KB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    KB_TRY(kb_alloc(&c->ticket, 4));
    KB_CUDA(cudaMemset(c->ticket, 0, 4 * sizeof(unsigned)));
    KB_TRY(kb_alloc(&c->comm_buf, 4096));
    KB_CUDA(cudaMallocHost((void**)&c->host_scalar, 4096 * sizeof(double)));
    *out = c;
    return KB_OK;
}
int kb_comm_destroy_internal(kb_ctx_s* c);
// Handles keep a reference on what they depend on (pc -> operator -> context), so destroying a
// parent handle first is safe: the object is released when its last dependant goes away.
extern "C" int kb_ctx_destroy(kb_ctx c) {
    if (!c) return KB_OK;
    kb_ctx_unref(c);
    return KB_OK;
}
void kb_ctx_unref(kb_ctx_s* c) {
    if (--c->refs > 0) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    kb_comm_destroy_internal(c);
    for (auto& e : c->prof_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto& e : c->event_pool) cudaEventDestroy(e);
    KB_FREE(c->ticket);
    KB_FREE(c->comm_buf);
    if (c->host_scalar) cudaFreeHost(c->host_scalar);
    cudaStreamDestroy(c->stream);
    delete c;
}
extern "C" void* kb_ctx_stream(kb_ctx c) { return (void*)c->stream; }
extern "C" int kb_ctx_device(kb_ctx c) { return c->device; }
extern "C" int kb_ctx_synchronize(kb_ctx c) { KB_CUDA(cudaSetDevice(c->device)); KB_CUDA(cudaStreamSynchronize(c->stream)); return KB_OK; }
extern "C" uint64_t kb_ctx_launch_count(kb_ctx c) { return c->launches; }

extern "C" void kb_partition_range(uint64_t n, uint64_t p, uint64_t r, uint64_t* lo, uint64_t* hi) {
    // chunk = (n + p - 1) / p ; block r = [r*chunk, min((r+1)*chunk, n))   (asm.rs:46-57)
    uint64_t chunk = (n + p - 1) / p;
    uint64_t s = r * chunk, e = (r + 1) * chunk;
    *lo = s > n ? n : s;
    *hi = e > n ? n : e;
}

int kb_hist_prepare(kb_csr_s* A, uint32_t flags, uint64_t max_entries, KbCtl* h) {
    A->hist_len = 0;
    if (!(flags & (KB_FLAG_HISTORY | KB_FLAG_MONITOR))) return KB_OK;
    const uint64_t want = std::min<uint64_t>(std::max<uint64_t>(max_entries, 1), 1ull << 22);
    if (want > A->hist_cap) {
        KB_FREE(A->hist_buf);
        A->hist_cap = 0;
        KB_TRY(kb_alloc(&A->hist_buf, want));
        A->hist_cap = want;
    }
    h->hist = A->hist_buf; h->hist_cap = A->hist_cap; h->hist_len = 0;
    return KB_OK;
}
extern "C" int kb_set_monitor(kb_csr A, kb_monitor_fn fn, void* user) {
    if (!A) { kb_set_error("null operator"); return KB_SOLVE_ERROR; }
    A->monitor = fn; A->monitor_user = user;
    return KB_OK;
}
extern "C" int kb_get_history(kb_csr A, double* out, uint64_t cap, uint64_t* len) {
    if (!A) { kb_set_error("null operator"); return KB_SOLVE_ERROR; }
    if (len) *len = A->hist_len;
    const uint64_t k = std::min<uint64_t>(std::min(A->hist_len, cap), A->hist_cap);
    if (k && out) {
        KB_CUDA(cudaSetDevice(A->ctx->device));
        KB_CUDA(cudaMemcpy(out, A->hist_buf, k * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return KB_OK;
}
int kb_upload_or_alias(kb_ctx_s* c, const double* src, double* dst, uint64_t n, bool device_ptrs) {
    if (n == 0) return KB_OK;
    KB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    return KB_OK;
}

// ---------------------------------------------------------------------------------------------
// operator upload: validation (new_checked semantics), i32 narrowing, row-length histogram
// ---------------------------------------------------------------------------------------------
__global__ void k_narrow(const unsigned long long* __restrict__ src, int* __restrict__ dst, size_t count,
                         unsigned long long limit, int* err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long v = src[i];
    if (v >= limit) { atomicExch(err, 1); v = 0; }
    dst[i] = (int)v;
}
// one thread per row: monotone row_ptr, strictly ascending in-range columns; histogram of lengths
__global__ void k_validate_hist(const int* __restrict__ rp, const int* __restrict__ col, int n, long long nnz,
                                unsigned long long* first_bad, unsigned long long* hist, unsigned long long* maxlen) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = rp[i], b = rp[i + 1];
    bool bad = (b < a) || (i == 0 && a != 0) || (i == n - 1 && (long long)b != nnz) || b > nnz || a < 0;
    if (!bad)
        for (int p = a + 1; p < b; ++p)
            if (col[p] <= col[p - 1]) { bad = true; break; }
    if (bad) { atomicMin(first_bad, (unsigned long long)i); return; }
    int len = b - a;
    int bucket = len <= 8 ? 0 : len <= 16 ? 1 : len <= 32 ? 2 : len <= 64 ? 3 : len <= 128 ? 4 : 5;
    atomicAdd(&hist[bucket], 1ull);
    atomicMax(maxlen, (unsigned long long)len);
}

static int upload_narrow(kb_ctx_s* c, const uint64_t* h_src, int* d_dst, size_t count, uint64_t limit, int* d_err) {
    const size_t CH = (size_t)1 << 24;   // 16M elements = 128 MiB staging
    unsigned long long* stage = nullptr;
    KB_TRY(kb_alloc(&stage, std::min(CH, std::max<size_t>(count, 1))));
    for (size_t off = 0; off < count; off += CH) {
        size_t m = std::min(CH, count - off);
        KB_CUDA(cudaMemcpyAsync(stage, h_src + off, m * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        {
            KbLaunch L(c, KB_K_OTHER);
            k_narrow<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(stage, d_dst + off, m, limit, d_err);
        }
        KB_CUDA(cudaStreamSynchronize(c->stream));
    }
    cudaFree(stage);
    return KB_OK;
}

int kb_csr_build_dist(kb_csr_s* A);   // kb_dist.cu: ghost list, column remap, halo plan

// chunk table of the bulk-async SpMV (rows grouped so each group's nnz fit one shared-memory stage)
static int build_chunk_table(kb_csr_s* A) {
    kb_ctx_s* c = A->ctx;
    if (getenv("KB_SPMV_KIND") && atoi(getenv("KB_SPMV_KIND")) == 0) return KB_OK;   // benchmarking knob: plain-load kernel
    const int nt = A->ntiles;
    int* d_flag = nullptr;
    KB_TRY(kb_alloc(&A->tile_chunk, (size_t)nt + 1));
    KB_TRY(kb_alloc(&d_flag, 1));
    KB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));
    { KbLaunch L(c, KB_K_OTHER); kb_chunk_build<<<(nt + 127) / 128, 128, 0, c->stream>>>(A->row_ptr, (int)A->n, nt, A->tile_chunk, nullptr, nullptr, 0, d_flag); }
    std::vector<int> cnt((size_t)nt + 1, 0);
    int flag = 0;
    KB_CUDA(cudaMemcpyAsync(cnt.data(), A->tile_chunk, nt * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_flag);
    if (flag) { KB_FREE(A->tile_chunk); return KB_OK; }   // some row exceeds a stage: keep the plain-load kernel
    int acc = 0;
    for (int t = 0; t < nt; ++t) { int k = cnt[t]; cnt[t] = acc; acc += k; }
    cnt[nt] = acc;
    A->nchunks = acc;
    KB_CUDA(cudaMemcpyAsync(A->tile_chunk, cnt.data(), ((size_t)nt + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    KB_TRY(kb_alloc(&A->chunk_row, (size_t)acc + 1));
    KB_TRY(kb_alloc(&A->chunk_nz, (size_t)acc + 1));
    KB_TRY(kb_alloc(&d_flag, 1));
    { KbLaunch L(c, KB_K_OTHER); kb_chunk_build<<<(nt + 127) / 128, 128, 0, c->stream>>>(A->row_ptr, (int)A->n, nt, A->tile_chunk, A->chunk_row, A->chunk_nz, 1, d_flag); }
    KB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_flag);
    A->kind = 2;
    // rows longer than ~12 leave most consumer threads without a row: gather per nonzero instead (histogram-driven)
    A->prod = A->n > 0 && (double)A->nnz / (double)A->n > 12.0;
    if (getenv("KB_SPMV_PROD")) A->prod = atoi(getenv("KB_SPMV_PROD")) != 0;
    return KB_OK;
}

// x-staging tables of kb_spmv_xtile (kb_spmv_xtile.cuh): a second chunk table, the x intervals of every chunk and the
// chunk-local 16-bit column ids.  All or nothing: if one chunk does not fit, the operator keeps kb_spmv_bulk.
// KB_SPMV_XTILE = 0 never, 1 whenever the chunks fit, 2 (default) only for long rows (the product-phase operators,
// where the gather is the limiter); KB_XT_CFG = 0 | 1 | 2 forces one stage geometry (default: 1, then 0; 2 is the short-row geometry).
static void free_xt_table(kb_csr_s* A) {
    A->xt = 0; A->xt_nchunks = 0;
    KB_FREE(A->xt_tile_chunk); KB_FREE(A->xt_chunk_row); KB_FREE(A->xt_chunk_nz);
    KB_FREE(A->xt_lo); KB_FREE(A->xt_len); KB_FREE(A->xt_tail); KB_FREE(A->xt_lcol);
}
// one geometry: A->xt = cfg + 1 when every chunk fits, tables freed otherwise
static int try_xt_table(kb_csr_s* A, int cfg) {
    kb_ctx_s* c = A->ctx;
    static const int caps[KB_XT_NCFG] = {KbXtCfg<0>::CAP, KbXtCfg<1>::CAP, KbXtCfg<2>::CAP};
    static const int xcaps[KB_XT_NCFG] = {KbXtCfg<0>::XCAP, KbXtCfg<1>::XCAP, KbXtCfg<2>::XCAP};
    static const int mrows[KB_XT_NCFG] = {KbXtCfg<0>::MAXROWS, KbXtCfg<1>::MAXROWS, KbXtCfg<2>::MAXROWS};
    if (cfg < 0 || cfg >= KB_XT_NCFG) return KB_OK;
    const int cap = caps[cfg], xcap = xcaps[cfg], maxrows = mrows[cfg];
    if (A->max_row_len > (uint64_t)cap) return KB_OK;
    const int nt = A->ntiles;
    int st = KB_OK;
    int* d_fail = nullptr;
    do {
        if ((st = kb_alloc(&A->xt_tile_chunk, (size_t)nt + 1)) != KB_OK) break;
        { KbLaunch L(c, KB_K_OTHER); kb_xt_chunk_build<<<(nt + 127) / 128, 128, 0, c->stream>>>(A->row_ptr, (int)A->n, nt, cap, maxrows, A->xt_tile_chunk, nullptr, nullptr, 0); }
        std::vector<int> cnt((size_t)nt + 1, 0);
        if (cudaMemcpyAsync(cnt.data(), A->xt_tile_chunk, nt * sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("x-tile chunk count failed: %s", cudaGetErrorString(cudaGetLastError())); st = KB_SOLVE_ERROR; break; }
        int acc = 0;
        for (int t = 0; t < nt; ++t) { int k = cnt[t]; cnt[t] = acc; acc += k; }
        cnt[nt] = acc;
        A->xt_nchunks = acc;
        if (cudaMemcpyAsync(A->xt_tile_chunk, cnt.data(), ((size_t)nt + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        if ((st = kb_alloc(&A->xt_chunk_row, (size_t)acc + 1)) != KB_OK) break;
        if ((st = kb_alloc(&A->xt_chunk_nz, (size_t)acc + 1)) != KB_OK) break;
        if ((st = kb_alloc(&A->xt_lo, (size_t)acc * KB_XT_KMAX)) != KB_OK) break;
        if ((st = kb_alloc(&A->xt_len, (size_t)acc * KB_XT_KMAX)) != KB_OK) break;
        if ((st = kb_alloc(&A->xt_tail, (size_t)acc)) != KB_OK) break;
        if ((st = kb_alloc(&A->xt_lcol, (size_t)A->nnz + 16)) != KB_OK) break;
        if ((st = kb_alloc(&d_fail, 1)) != KB_OK) break;
        cudaMemsetAsync(d_fail, 0, sizeof(int), c->stream);
        cudaMemsetAsync(A->xt_lcol + A->nnz, 0, 16 * sizeof(unsigned short), c->stream);
        { KbLaunch L(c, KB_K_OTHER); kb_xt_chunk_build<<<(nt + 127) / 128, 128, 0, c->stream>>>(A->row_ptr, (int)A->n, nt, cap, maxrows, A->xt_tile_chunk, A->xt_chunk_row, A->xt_chunk_nz, 1); }
        {
            KbLaunch L(c, KB_K_OTHER);
            if (cfg == 1) kb_xt_build<KbXtCfg<1>::CAP><<<acc, KB_THREADS, 0, c->stream>>>(A->col, A->xt_chunk_nz, (int)A->ncols_local, xcap, A->xt_lo, A->xt_len, A->xt_tail, A->xt_lcol, d_fail);
            else if (cfg == 2) kb_xt_build<KbXtCfg<2>::CAP><<<acc, KB_THREADS, 0, c->stream>>>(A->col, A->xt_chunk_nz, (int)A->ncols_local, xcap, A->xt_lo, A->xt_len, A->xt_tail, A->xt_lcol, d_fail);
            else kb_xt_build<KbXtCfg<0>::CAP><<<acc, KB_THREADS, 0, c->stream>>>(A->col, A->xt_chunk_nz, (int)A->ncols_local, xcap, A->xt_lo, A->xt_len, A->xt_tail, A->xt_lcol, d_fail);
        }
        int fail = 0;
        if (cudaMemcpyAsync(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("x-tile table build failed: %s", cudaGetErrorString(cudaGetLastError())); st = KB_SOLVE_ERROR; break; }
        if (!fail) A->xt = cfg + 1;
    } while (0);
    if (d_fail) cudaFree(d_fail);
    if (st != KB_OK || !A->xt) free_xt_table(A);
    return st;
}
static int build_xt_table(kb_csr_s* A) {
    const int mode = getenv("KB_SPMV_XTILE") ? atoi(getenv("KB_SPMV_XTILE")) : KB_XT_DEFAULT_MODE;
    if (mode <= 0 || A->kind != 2 || A->dist || A->n == 0 || A->nnz == 0) return KB_OK;
    if (mode >= 2 && !A->prod) return KB_OK;
    // thread per row out of shared memory up to 32 entries per row on average, per-nonzero product phase beyond
    A->xt_prod = (double)A->nnz / (double)A->n > 32.0;
    if (getenv("KB_SPMV_PROD")) A->xt_prod = atoi(getenv("KB_SPMV_PROD")) != 0;
    if (getenv("KB_XT_CFG")) return try_xt_table(A, atoi(getenv("KB_XT_CFG")));     // tuning: this geometry or none
    KB_TRY(try_xt_table(A, KB_XT_DEFAULT_CFG));
    if (!A->xt) KB_TRY(try_xt_table(A, 1 - KB_XT_DEFAULT_CFG));
    return KB_OK;
}

// allocate the device arrays of an operator (padded tails zeroed); the caller fills row_ptr / col / vals
int kb_csr_alloc(kb_ctx c, uint64_t nrows, uint64_t ncols_global, uint64_t nnz, kb_csr_s** out) {
    *out = nullptr;
    if (nrows >= (1ull << 31) - 2 * KB_TILE || nnz >= (1ull << 31) - 16 || ncols_global >= (1ull << 31) - 2 * KB_TILE) {
        kb_set_error("matrix shard too large for i32 device indices (rows %llu, nnz %llu): partition it across more GPUs",
                     (unsigned long long)nrows, (unsigned long long)nnz);
        return KB_UNSUPPORTED;
    }
    kb_csr_s* A = new kb_csr_s;
    A->ctx = c; A->n = nrows; A->ncols_global = ncols_global; A->ncols_local = ncols_global; A->nnz = nnz;
    A->ntiles = std::max(1, kb_num_tiles(nrows));   // an empty shard still launches one (empty) tile: it must join the in-kernel all-reduces
    A->dist = false; A->n_global = nrows; A->row_lo = 0; A->row_hi = nrows;
    c->refs++;
    int st = KB_OK;
    do {
        if ((st = kb_alloc(&A->row_ptr, nrows + 1 + 8)) != KB_OK) break;
        cudaMemsetAsync(A->row_ptr, 0, (nrows + 9) * sizeof(int), c->stream);
        if ((st = kb_alloc(&A->col, nnz + 8)) != KB_OK) break;
        if ((st = kb_alloc(&A->vals, nnz + 8)) != KB_OK) break;
        cudaMemsetAsync(A->col + nnz, 0, 8 * sizeof(int), c->stream);
        cudaMemsetAsync(A->vals + nnz, 0, 8 * sizeof(double), c->stream);
    } while (0);
    if (st != KB_OK) { kb_csr_destroy(A); return st; }
    *out = A;
    return KB_OK;
}

// validate the device arrays (new_checked rules), row-length histogram -> kernel choice, shard maps, chunk table.
// d_err (optional): device flag raised by the index-narrowing upload.
int kb_csr_finalize(kb_csr_s* A, const int* d_err) {
    kb_ctx_s* c = A->ctx;
    const uint64_t nrows = A->n, nnz = A->nnz;
    unsigned long long* d_stats = nullptr;   // [0] first_bad, [1..6] hist, [7] maxlen
    int st = KB_OK;
    do {
        if ((st = kb_alloc(&d_stats, 8)) != KB_OK) break;
        cudaMemsetAsync(d_stats, 0, 8 * sizeof(unsigned long long), c->stream);
        cudaMemsetAsync(d_stats, 0xFF, sizeof(unsigned long long), c->stream);
        if (nrows) {
            KbLaunch L(c, KB_K_OTHER);
            k_validate_hist<<<(unsigned)((nrows + 255) / 256), 256, 0, c->stream>>>(A->row_ptr, A->col, (int)nrows, (long long)nnz,
                                                                                  d_stats, d_stats + 1, d_stats + 7);
        }
        int h_err = 0;
        unsigned long long h_stats[8];
        if ((d_err && cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) ||
            cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) {
            kb_set_error("CSR validation failed to run: %s", cudaGetErrorString(cudaGetLastError())); st = KB_SOLVE_ERROR; break;
        }
        if (h_err) { kb_set_error("invalid CSR: index out of range (column >= ncols or row_ptr > nnz)"); st = KB_SOLVE_ERROR; break; }
        if (h_stats[0] != ~0ull) {
            kb_set_error("invalid CSR at row %llu: row_ptr not monotone or columns not strictly ascending", h_stats[0]);
            st = KB_SOLVE_ERROR; break;
        }
        for (int k = 0; k < 6; ++k) A->hist[k] = h_stats[1 + k];
        A->max_row_len = h_stats[7];
        // Row-length histogram -> kernel choice: CSR-stream (thread per row out of shared memory)
        // for short/regular rows, vector-per-row when most rows are long.
        uint64_t longrows = A->hist[4] + A->hist[5];          // > 64 nnz
        if (nrows && longrows * 2 > nrows) {
            A->kind = 1;
            double mean = (double)nnz / (double)nrows;
            A->vec = mean > 256 ? 32 : mean > 128 ? 16 : 8;
        }
        if (A->dist && (st = kb_csr_build_dist(A)) != KB_OK) break;
        if (A->kind == 0 && nrows && (st = build_chunk_table(A)) != KB_OK) break;
        if (A->kind == 2 && (st = build_xt_table(A)) != KB_OK) break;
    } while (0);
    if (d_stats) cudaFree(d_stats);
    return st;
}

static int csr_create_common(kb_ctx c, uint64_t nrows, uint64_t ncols_global, bool dist, uint64_t n_global, uint64_t lo,
                             uint64_t hi, const uint64_t* row_ptr, const uint64_t* col_idx, const double* vals, kb_csr* out) {
    *out = nullptr;
    if (!c) { kb_set_error("null context"); return KB_SOLVE_ERROR; }
    KB_CUDA(cudaSetDevice(c->device));
    if (!row_ptr || (nrows > 0 && row_ptr[nrows] > 0 && (!col_idx || !vals))) { kb_set_error("null CSR array"); return KB_SOLVE_ERROR; }
    const uint64_t nnz = nrows ? row_ptr[nrows] : 0;
    kb_csr_s* A = nullptr;
    KB_TRY(kb_csr_alloc(c, nrows, ncols_global, nnz, &A));
    A->dist = dist; A->n_global = n_global; A->row_lo = lo; A->row_hi = hi;
    int st = KB_OK;
    int* d_err = nullptr;
    do {
        if ((st = kb_alloc(&d_err, 1)) != KB_OK) break;
        cudaMemsetAsync(d_err, 0, sizeof(int), c->stream);
        if ((st = upload_narrow(c, row_ptr, A->row_ptr, nrows + 1, nnz + 1, d_err)) != KB_OK) break;
        if ((st = upload_narrow(c, col_idx, A->col, nnz, ncols_global, d_err)) != KB_OK) break;
        if (nnz && cudaMemcpyAsync(A->vals, vals, nnz * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
            kb_set_error("H2D copy of values failed"); st = KB_SOLVE_ERROR; break;
        }
        st = kb_csr_finalize(A, d_err);
    } while (0);
    if (d_err) cudaFree(d_err);
    if (st != KB_OK) { kb_csr_destroy(A); return st; }
    *out = A;
    return KB_OK;
}

extern "C" int kb_csr_create(kb_ctx c, uint64_t nrows, uint64_t ncols, const uint64_t* row_ptr, const uint64_t* col_idx,
                             const double* vals, kb_csr* out) {
    return csr_create_common(c, nrows, ncols, false, nrows, 0, nrows, row_ptr, col_idx, vals, out);
}
extern "C" int kb_csr_create_dist(kb_ctx c, uint64_t n_global, uint64_t row_lo, uint64_t row_hi, const uint64_t* row_ptr,
                                  const uint64_t* col_idx, const double* vals, kb_csr* out) {
    *out = nullptr;
    if (!c) { kb_set_error("null context"); return KB_SOLVE_ERROR; }
    uint64_t lo, hi;
    kb_partition_range(n_global, (uint64_t)c->size, (uint64_t)c->rank, &lo, &hi);
    if (lo != row_lo || hi != row_hi) {
        kb_set_error("rank %d must own rows [%llu,%llu) (chunk partition, asm.rs:46-57), got [%llu,%llu)", c->rank,
                     (unsigned long long)lo, (unsigned long long)hi, (unsigned long long)row_lo, (unsigned long long)row_hi);
        return KB_SOLVE_ERROR;
    }
    return csr_create_common(c, row_hi - row_lo, n_global, true, n_global, row_lo, row_hi, row_ptr, col_idx, vals, out);
}

extern "C" int kb_csr_destroy(kb_csr A) {
    if (!A) return KB_OK;
    kb_csr_unref(A);
    return KB_OK;
}
void kb_csr_unref(kb_csr_s* A) {
    if (--A->refs > 0) return;
    cudaSetDevice(A->ctx->device);
    cudaStreamSynchronize(A->ctx->stream);
    kb_pcg_ws_free(A->pcg_ws);
    kb_bicg_ws_free(A->bicg_ws);
    kb_gmres_ws_free(A->gmres_ws);
    kb_halo_free(A->halo);
    KB_FREE(A->row_ptr); KB_FREE(A->col); KB_FREE(A->vals); KB_FREE(A->ghosts);
    free_xt_table(A);
    KB_FREE(A->tile_chunk); KB_FREE(A->chunk_row); KB_FREE(A->chunk_nz); KB_FREE(A->tiles_interior); KB_FREE(A->tiles_boundary); KB_FREE(A->tiles_order);
    KB_FREE(A->x_tmp); KB_FREE(A->y_tmp); KB_FREE(A->hist_buf);
    kb_ctx_s* c = A->ctx;
    delete A;
    kb_ctx_unref(c);
}
extern "C" uint64_t kb_csr_nrows(kb_csr A) { return A->n; }
extern "C" uint64_t kb_csr_ncols(kb_csr A) { return A->ncols_global; }
extern "C" uint64_t kb_csr_nnz(kb_csr A) { return A->nnz; }
extern "C" int kb_csr_spmv_kernel_kind(kb_csr A) { return A->kind; }
extern "C" int kb_csr_spmv_x_staged(kb_csr A) { return A->xt; }
extern "C" uint64_t kb_csr_num_ghosts(kb_csr A) { return A->nghost; }
extern "C" int kb_csr_get_ghosts(kb_csr A, uint64_t* out) {
    if (A->nghost == 0) return KB_OK;
    KB_CUDA(cudaSetDevice(A->ctx->device));
    KB_CUDA(cudaMemcpy(out, A->ghosts, A->nghost * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return KB_OK;
}

// ---------------------------------------------------------------------------------------------
// MatVec::matvec
// ---------------------------------------------------------------------------------------------
int kb_csr_spmv_plain(kb_csr_s* A, const double* d_x, double* d_y) {
    return kb_launch_spmv<KbEpiNone, false>(A, d_x, d_y, nullptr, nullptr, nullptr, 0, KbEpiNone{});
}

extern "C" int kb_csr_matvec_device(kb_csr A, const double* d_x, double* d_y) {
    KB_CUDA(cudaSetDevice(A->ctx->device));
    if (A->dist) {
        // x must have room for ghosts: stage into the shard-local operand
        if (!A->x_tmp) KB_TRY(kb_alloc(&A->x_tmp, A->ncols_local));
        KB_CUDA(cudaMemcpyAsync(A->x_tmp, d_x, A->n * sizeof(double), cudaMemcpyDeviceToDevice, A->ctx->stream));
        KB_TRY(kb_halo_exchange(A, A->x_tmp));
        KB_TRY(kb_csr_spmv_plain(A, A->x_tmp, d_y));
    } else {
        KB_TRY(kb_csr_spmv_plain(A, d_x, d_y));
    }
    KB_CUDA(cudaStreamSynchronize(A->ctx->stream));
    return KB_OK;
}

extern "C" int kb_csr_matvec(kb_csr A, const double* x, double* y) {
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    if (!A->x_tmp) KB_TRY(kb_alloc(&A->x_tmp, A->ncols_local));
    if (!A->y_tmp) KB_TRY(kb_alloc(&A->y_tmp, A->n));
    const uint64_t nx = A->dist ? A->n : A->ncols_local;
    KB_CUDA(cudaMemcpyAsync(A->x_tmp, x, nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (A->dist) KB_TRY(kb_halo_exchange(A, A->x_tmp));
    KB_TRY(kb_csr_spmv_plain(A, A->x_tmp, A->y_tmp));
    KB_CUDA(cudaMemcpyAsync(y, A->y_tmp, A->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    return KB_OK;
}

// ---------------------------------------------------------------------------------------------
// InnerProduct::{dot,norm}  (wrappers.rs:90-128) with the canonical tree
// ---------------------------------------------------------------------------------------------
struct DotOp : KbRedBase {
    static constexpr int NRED = 1;
    const double* x; const double* y; double* out;
    __device__ bool skip() const { return false; }
    __device__ void pair(long long i, bool has1, double* red) const {
        double e0 = x[i] * y[i];
        double e1 = has1 ? x[i + 1] * y[i + 1] : 0.0;
        red[0] = e0 + e1;
    }
    __device__ void finish_block(double* s) const { if (threadIdx.x == 0) out[0] = s[0]; }
};

static int dot_host(kb_ctx c, uint64_t n, const double* x, const double* y, double* out) {
    KB_CUDA(cudaSetDevice(c->device));
    if (n == 0) {      // an empty shard still takes part in the collective
        if (c->size > 1) return kb_comm_all_reduce(c, 0.0, out);
        *out = 0.0; return KB_OK;
    }
    double *dx = nullptr, *dy = nullptr, *part = nullptr;
    int P = kb_num_tiles(n);
    KB_TRY(kb_alloc(&dx, n));
    if (y != x) KB_TRY(kb_alloc(&dy, n)); else dy = dx;
    KB_TRY(kb_alloc(&part, (size_t)P + 1));
    KB_CUDA(cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (y != x) KB_CUDA(cudaMemcpyAsync(dy, y, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    DotOp op; op.n = (long long)n; op.partials = part; op.pstride = P; op.ticket = c->ticket; op.x = dx; op.y = dy; op.out = part + P;
    { KbLaunch L(c, KB_K_SMALL); kb_tile_kernel<DotOp><<<P, KB_THREADS, 0, c->stream>>>(op); }
    KB_CUDA(cudaGetLastError());
    double local = 0.0;
    KB_CUDA(cudaMemcpyAsync(&local, part + P, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(dx); if (dy != dx) cudaFree(dy); cudaFree(part);
    if (c->size > 1) return kb_comm_all_reduce(c, local, out);
    *out = local;
    return KB_OK;
}
extern "C" int kb_dot(kb_ctx c, uint64_t n, const double* x, const double* y, double* out) { return dot_host(c, n, x, y, out); }
extern "C" int kb_norm(kb_ctx c, uint64_t n, const double* x, double* out) {
    double d = 0.0;
    KB_TRY(dot_host(c, n, x, x, &d));
    *out = sqrt(d);
    return KB_OK;
}

// ---------------------------------------------------------------------------------------------
// Jacobi (jacobi.rs:53-95): inv_diag[i] = a_ii != 0 ? 1/a_ii : 0, read directly from the CSR row
// ---------------------------------------------------------------------------------------------
__global__ void k_jacobi_setup(const int* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ vals,
                               int n, double* __restrict__ inv_diag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = 0.0;
    for (int p = rp[i]; p < rp[i + 1]; ++p)
        if (col[p] == i) d = d + vals[p];
    inv_diag[i] = (d != 0.0) ? 1.0 / d : 0.0;
}
struct JacobiApplyOp : KbRedBase {
    static constexpr int NRED = 0;
    const double* inv; const double* r; double* z; const KbCtl* sc; int sm;
    __device__ bool skip() const { return kb_skip(sc, sm); }
    __device__ void pair(long long i, bool has1, double*) const {
        if (has1) { double2 a = kb_ld2(inv + i), b = kb_ld2(r + i); kb_st2(z + i, make_double2(a.x * b.x, a.y * b.y)); }
        else z[i] = inv[i] * r[i];
    }
    __device__ void finish_block(double*) const {}
};

extern "C" int kb_pc_create_jacobi(kb_csr A, kb_pc* out) {
    *out = nullptr;
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    kb_pc_s* pc = new kb_pc_s;
    pc->a = A; pc->ctx = c; pc->kind = KB_PC_JACOBI;
    int st = kb_alloc(&pc->inv_diag, A->n + 2);
    if (st != KB_OK) { delete pc; return st; }
    A->refs++;
    if (A->n) {
        KbLaunch L(c, KB_K_OTHER);
        k_jacobi_setup<<<(unsigned)((A->n + 255) / 256), 256, 0, c->stream>>>(A->row_ptr, A->col, A->vals, (int)A->n, pc->inv_diag);
    }
    KB_CUDA(cudaStreamSynchronize(c->stream));
    *out = pc;
    return KB_OK;
}

int kb_ilu0_apply_dev(kb_pc_s* pc, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask);

int kb_pc_apply_dev(kb_pc_s* pc, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask) {
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    if (A->n == 0) return KB_OK;
    if (pc->kind == KB_PC_JACOBI) {
        JacobiApplyOp op; op.n = (long long)A->n; op.partials = nullptr; op.pstride = 0; op.ticket = c->ticket;
        op.inv = pc->inv_diag; op.r = d_r; op.z = d_z; op.sc = skip_ctl; op.sm = skip_mask;
        KbLaunch L(c, KB_K_SMALL);
        kb_tile_kernel<JacobiApplyOp><<<A->ntiles, KB_THREADS, 0, c->stream>>>(op);
        KB_CUDA(cudaGetLastError());
        return KB_OK;
    }
    if (pc->kind == KB_PC_ILU0) return kb_ilu0_apply_dev(pc, d_r, d_z, skip_ctl, skip_mask);
    if (pc->kind == KB_PC_ASM) return kb_asm_apply_dev(pc, d_r, d_z, skip_ctl, skip_mask);
    kb_set_error("unknown preconditioner kind");
    return KB_UNSUPPORTED;
}
extern "C" int kb_pc_apply_device(kb_pc pc, const double* d_r, double* d_z) {
    KB_CUDA(cudaSetDevice(pc->a->ctx->device));
    KB_TRY(kb_pc_apply_dev(pc, d_r, d_z));
    KB_CUDA(cudaStreamSynchronize(pc->a->ctx->stream));
    if (kb_ilu0_error(pc) || kb_asm_error(pc)) { kb_set_error("ilu0: a triangular-solve dependency wait timed out"); return KB_SOLVE_ERROR; }
    return KB_OK;
}
extern "C" int kb_pc_apply(kb_pc pc, const double* r, double* z) {
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    if (!pc->r_tmp) KB_TRY(kb_alloc(&pc->r_tmp, A->n + 2));
    if (!pc->z_tmp) KB_TRY(kb_alloc(&pc->z_tmp, A->n + 2));
    KB_CUDA(cudaMemcpyAsync(pc->r_tmp, r, A->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    KB_TRY(kb_pc_apply_dev(pc, pc->r_tmp, pc->z_tmp));
    KB_CUDA(cudaMemcpyAsync(z, pc->z_tmp, A->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    if (kb_ilu0_error(pc) || kb_asm_error(pc)) { kb_set_error("ilu0: a triangular-solve dependency wait timed out"); return KB_SOLVE_ERROR; }
    return KB_OK;
}
extern "C" int kb_pc_destroy(kb_pc pc) {
    if (!pc) return KB_OK;
    cudaSetDevice(pc->ctx->device);
    cudaStreamSynchronize(pc->ctx->stream);
    kb_ilu0_free(pc);
    kb_asm_free(pc);
    KB_FREE(pc->inv_diag); KB_FREE(pc->r_tmp); KB_FREE(pc->z_tmp);
    kb_csr_s* A = pc->a;
    delete pc;
    kb_csr_unref(A);
    return KB_OK;
}
extern "C" uint64_t kb_pc_bad_row(kb_pc pc) { return pc->bad_row; }
extern "C" int kb_pc_get_inv_diag(kb_pc pc, double* out) {
    KB_CUDA(cudaSetDevice(pc->a->ctx->device));
    if (pc->a->n) KB_CUDA(cudaMemcpy(out, pc->inv_diag, pc->a->n * sizeof(double), cudaMemcpyDeviceToHost));
    return KB_OK;
}
