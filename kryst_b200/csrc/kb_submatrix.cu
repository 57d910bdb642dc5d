// kb_submatrix.cu — SubmatrixExtract on the device and CSR read-back (the step before the path).
//
// Reference: `SubmatrixExtract::submatrix(&self, indices)` (src/core/traits.rs, impl src/matrix/sparse.rs:72-93)
// densifies the whole matrix, builds sub[i][j] = A[indices[i]][indices[j]] and re-sparsifies, dropping zeros;
// AdditiveSchwarz::setup (src/preconditioner/asm.rs:58-65) calls it once per subdomain.  Same result here without
// the dense detour, all on the device (integer work, bit-exact vs the oracle):
//   1. sort the pairs (indices[j], j) by global index (stable radix sort: ties keep ascending j, so repeated
//      indices are handled like the reference's dense definition)
//   2. one thread per output row i: walk row indices[i] of A, and for every stored non-zero with column g count the
//      pairs whose key is g (two binary searches)                                     -> exclusive scan = row_ptr
//   3. same walk again, writing (i << 32 | j, value) records                           -> 64-bit radix sort
//      (row-major, ascending local column = the reference's `for j in 0..n` order)
//   4. unpack into a new operator and run the usual validation / kernel selection.
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>
#include "kb_objects.h"

int kb_csr_alloc(kb_ctx c, uint64_t nrows, uint64_t ncols_global, uint64_t nnz, kb_csr_s** out);
int kb_csr_finalize(kb_csr_s* A, const int* d_err);

__global__ void k_sub_narrow(const unsigned long long* __restrict__ src, int* __restrict__ keys, int* __restrict__ pos, int k,
                             unsigned long long limit, int* err) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    unsigned long long v = src[j];
    if (v >= limit) { atomicExch(err, 1); v = 0; }
    keys[j] = (int)v;
    pos[j] = j;
}
__device__ __forceinline__ int kb_lower_bound(const int* __restrict__ a, int n, int key) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
// FILL == false: cnt[i] = number of output entries of row i;  FILL == true: write the records at off[i]
template <bool FILL>
__global__ void k_sub_rows(const int* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ vals,
                           const int* __restrict__ idx, const int* __restrict__ skey, const int* __restrict__ spos, int k,
                           long long* __restrict__ cnt, const long long* __restrict__ off,
                           unsigned long long* __restrict__ rec_key, double* __restrict__ rec_val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const int row = idx[i];
    long long c = 0, o = FILL ? off[i] : 0;
    for (int p = rp[row]; p < rp[row + 1]; ++p) {
        const double v = vals[p];
        if (!(v != 0.0)) continue;                       // sparse.rs:84: explicit zeros are dropped
        const int g = col[p];
        int q = kb_lower_bound(skey, k, g);
        for (; q < k && skey[q] == g; ++q) {
            if (FILL) { rec_key[o] = ((unsigned long long)i << 32) | (unsigned)spos[q]; rec_val[o] = v; ++o; }
            else ++c;
        }
    }
    if (!FILL) cnt[i] = c;
}
__global__ void k_sub_unpack(const unsigned long long* __restrict__ key, long long nnz, int* __restrict__ col) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) col[p] = (int)(key[p] & 0xffffffffull);
}
__global__ void k_sub_rowptr(const long long* __restrict__ off, int k, int* __restrict__ rp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= k) rp[i] = (int)off[i];
}

extern "C" int kb_csr_submatrix(kb_csr A, const uint64_t* indices, uint64_t k64, kb_csr* out) {
    if (!out) { kb_set_error("kb_csr_submatrix: null argument"); return KB_SOLVE_ERROR; }
    *out = nullptr;
    if (!A || (k64 && !indices)) { kb_set_error("kb_csr_submatrix: null argument"); return KB_SOLVE_ERROR; }
    if (A->dist) { kb_set_error("kb_csr_submatrix: not available on a row-block shard (extract before partitioning)"); return KB_UNSUPPORTED; }
    if (k64 >= (1ull << 31) - 2 * KB_TILE) { kb_set_error("kb_csr_submatrix: index set too large"); return KB_UNSUPPORTED; }
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    const int k = (int)k64;
    const uint64_t limit = std::min<uint64_t>(A->n, A->ncols_global);   // indices address rows and columns
    unsigned long long* d_idx64 = nullptr; int *d_idx = nullptr, *d_pos = nullptr, *d_skey = nullptr, *d_spos = nullptr, *d_err = nullptr;
    long long *d_cnt = nullptr, *d_off = nullptr;
    unsigned long long *d_k0 = nullptr, *d_k1 = nullptr; double *d_v0 = nullptr;
    void* tmp = nullptr;
    kb_csr_s* S = nullptr;
    int st = KB_OK;
    auto fail = [&](const char* what) { kb_set_error("kb_csr_submatrix: %s: %s", what, cudaGetErrorString(cudaGetLastError())); st = KB_SOLVE_ERROR; };
    do {
        if ((st = kb_alloc(&d_idx64, (size_t)k)) != KB_OK || (st = kb_alloc(&d_idx, (size_t)k)) != KB_OK || (st = kb_alloc(&d_pos, (size_t)k)) != KB_OK ||
            (st = kb_alloc(&d_skey, (size_t)k)) != KB_OK || (st = kb_alloc(&d_spos, (size_t)k)) != KB_OK || (st = kb_alloc(&d_err, 1)) != KB_OK ||
            (st = kb_alloc(&d_cnt, (size_t)k + 1)) != KB_OK || (st = kb_alloc(&d_off, (size_t)k + 1)) != KB_OK) break;
        cudaMemsetAsync(d_err, 0, sizeof(int), c->stream);
        cudaMemsetAsync(d_cnt, 0, ((size_t)k + 1) * sizeof(long long), c->stream);
        long long total = 0;
        if (k) {
            if (cudaMemcpyAsync(d_idx64, indices, (size_t)k * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { fail("H2D copy"); break; }
            { KbLaunch L(c, KB_K_OTHER); k_sub_narrow<<<(k + 255) / 256, 256, 0, c->stream>>>(d_idx64, d_idx, d_pos, k, limit, d_err); }
            int h_err = 0;
            if (cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { fail("index check"); break; }
            if (h_err) { kb_set_error("kb_csr_submatrix: index out of range (>= min(nrows, ncols) = %llu)", (unsigned long long)limit); st = KB_SOLVE_ERROR; break; }
            size_t tb = 0, tb2 = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, d_idx, d_skey, d_pos, d_spos, k, 0, 32, c->stream);
            cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_cnt, d_off, k + 1, c->stream);
            tb = std::max(tb, tb2);
            if (cudaMalloc(&tmp, std::max<size_t>(tb, 1)) != cudaSuccess) { fail("scratch allocation"); break; }
            { KbLaunch L(c, KB_K_OTHER); cub::DeviceRadixSort::SortPairs(tmp, tb, d_idx, d_skey, d_pos, d_spos, k, 0, 32, c->stream); }
            { KbLaunch L(c, KB_K_OTHER); k_sub_rows<false><<<(k + 127) / 128, 128, 0, c->stream>>>(A->row_ptr, A->col, A->vals, d_idx, d_skey, d_spos, k, d_cnt, nullptr, nullptr, nullptr); }
            { KbLaunch L(c, KB_K_OTHER); cub::DeviceScan::ExclusiveSum(tmp, tb, d_cnt, d_off, k + 1, c->stream); }
            if (cudaMemcpyAsync(&total, d_off + k, sizeof(long long), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { fail("count pass"); break; }
            cudaFree(tmp); tmp = nullptr;
        } else cudaMemsetAsync(d_off, 0, sizeof(long long), c->stream);
        if ((st = kb_csr_alloc(c, k64, k64, (uint64_t)total, &S)) != KB_OK) break;
        if (k) { KbLaunch L(c, KB_K_OTHER); k_sub_rowptr<<<(k + 256) / 256, 256, 0, c->stream>>>(d_off, k, S->row_ptr); }
        if (total) {
            if ((st = kb_alloc(&d_k0, (size_t)total)) != KB_OK || (st = kb_alloc(&d_k1, (size_t)total)) != KB_OK || (st = kb_alloc(&d_v0, (size_t)total)) != KB_OK) break;
            { KbLaunch L(c, KB_K_OTHER); k_sub_rows<true><<<(k + 127) / 128, 128, 0, c->stream>>>(A->row_ptr, A->col, A->vals, d_idx, d_skey, d_spos, k, nullptr, d_off, d_k0, d_v0); }
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, d_k0, d_k1, d_v0, S->vals, (int)total, 0, 64, c->stream);
            if (cudaMalloc(&tmp, std::max<size_t>(tb, 1)) != cudaSuccess) { fail("scratch allocation"); break; }
            { KbLaunch L(c, KB_K_OTHER); cub::DeviceRadixSort::SortPairs(tmp, tb, d_k0, d_k1, d_v0, S->vals, (int)total, 0, 64, c->stream); }
            { KbLaunch L(c, KB_K_OTHER); k_sub_unpack<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(d_k1, total, S->col); }
            if (cudaGetLastError() != cudaSuccess) { fail("fill pass"); break; }
        }
        st = kb_csr_finalize(S, nullptr);
    } while (0);
    if (tmp) cudaFree(tmp);
    KB_FREE(d_idx64); KB_FREE(d_idx); KB_FREE(d_pos); KB_FREE(d_skey); KB_FREE(d_spos); KB_FREE(d_err); KB_FREE(d_cnt); KB_FREE(d_off);
    KB_FREE(d_k0); KB_FREE(d_k1); KB_FREE(d_v0);
    if (st != KB_OK) { if (S) kb_csr_destroy(S); return st; }
    *out = S;
    return KB_OK;
}

// ---- read an operator back (usize indices, like the arrays handed to CsrMatrix::from_csr) -----------------------
__global__ void k_widen(const int* __restrict__ src, unsigned long long* __restrict__ dst, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = (unsigned long long)src[i];
}
static int download_wide(kb_ctx_s* c, const int* d_src, uint64_t* h_dst, size_t count) {
    const size_t CH = (size_t)1 << 24;
    unsigned long long* stage = nullptr;
    KB_TRY(kb_alloc(&stage, std::min(CH, std::max<size_t>(count, 1))));
    int st = KB_OK;
    for (size_t off = 0; off < count && st == KB_OK; off += CH) {
        const size_t m = std::min(CH, count - off);
        { KbLaunch L(c, KB_K_OTHER); k_widen<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(d_src + off, stage, m); }
        if (cudaMemcpyAsync(h_dst + off, stage, m * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("kb_csr_download: D2H copy failed"); st = KB_SOLVE_ERROR; }
    }
    cudaFree(stage);
    return st;
}
extern "C" int kb_csr_download(kb_csr A, uint64_t* row_ptr, uint64_t* col_idx, double* vals) {
    if (!A || !row_ptr) { kb_set_error("kb_csr_download: null argument"); return KB_SOLVE_ERROR; }
    if (A->dist) { kb_set_error("kb_csr_download: not available on a row-block shard (columns are renumbered)"); return KB_UNSUPPORTED; }
    if (A->nnz && (!col_idx || !vals)) { kb_set_error("kb_csr_download: null argument"); return KB_SOLVE_ERROR; }
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    KB_TRY(download_wide(c, A->row_ptr, row_ptr, (size_t)A->n + 1));
    if (A->nnz) {
        KB_TRY(download_wide(c, A->col, col_idx, (size_t)A->nnz));
        KB_CUDA(cudaMemcpyAsync(vals, A->vals, (size_t)A->nnz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        KB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return KB_OK;
}
