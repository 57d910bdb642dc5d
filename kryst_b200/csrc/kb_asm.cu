// kb_asm.cu — AdditiveSchwarz as a preconditioner object on ONE GPU (src/preconditioner/asm.rs:17-116).
//
//   AdditiveSchwarz::new(overlap, subdomains)            asm.rs:34-36
//   setup: empty `subdomains` -> uniform row chunks, chunk = ceil(n/p) (asm.rs:46-57); per block
//          a_sub = a.submatrix(indices) (asm.rs:59-64) and an inner solver from the factory
//   apply: z = 0; for every block IN ORDER: r_blk = r[indices]; x_blk = inner(a_sub, r_blk); z[indices] += x_blk
//          (asm.rs:76-116: the block results are summed serially in subdomain order)
// Here: the sub-operators come from kb_csr_submatrix (device, bit-exact vs the oracle), the "inner solver" is one
// application of a device preconditioner of the block (KB_ASM_INNER_ILU0: textbook ILU(0) of a_sub, i.e. block-Jacobi
// ILU(0) when the blocks are disjoint; KB_ASM_INNER_JACOBI) — an inner Krylov solve cannot live inside the captured
// iteration graphs of the outer solvers and is not offered.  `overlap` is stored but never used by the reference
// (asm.rs:19,34-36 — the index lists are taken as given); as an extension overlap = k > 0 grows every index set by k
// layers of graph neighbours through A's stored pattern (PETSc PCASM's meaning) and sorts it ascending; overlap = 0
// keeps the caller's order exactly.  gather / block apply / scatter-add run on the library stream in block order, so
// the sum order is the reference's and results are bit-identical to the oracle's restatement.
#include <algorithm>
#include <vector>
#include "kb_objects.h"

struct KbAsmBlock {
    int n = 0;
    int* idx = nullptr;            // device: global row of every block row
    kb_csr_s* sub = nullptr;
    kb_pc_s* inner = nullptr;
    double* r_blk = nullptr;
    double* x_blk = nullptr;
};
struct KbAsm {
    std::vector<KbAsmBlock> blocks;
    uint64_t overlap = 0;
};

__global__ void k_asm_zero(double* z, long long n, const KbCtl* sc, int sm) {
    if (kb_skip(sc, sm)) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = 0.0;
}
__global__ void k_asm_gather(const double* __restrict__ r, const int* __restrict__ idx, double* __restrict__ out, int k, const KbCtl* sc, int sm) {
    if (kb_skip(sc, sm)) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) out[j] = r[idx[j]];
}
__global__ void k_asm_scatter_add(double* z, const int* __restrict__ idx, const double* __restrict__ x, int k, const KbCtl* sc, int sm) {
    if (kb_skip(sc, sm)) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < k) { const int g = idx[j]; z[g] = z[g] + x[j]; }       // indices of one block are distinct (checked at setup)
}
// one layer of graph neighbours: rows with mark == layer give mark layer+1 to their unmarked columns
__global__ void k_asm_grow(const int* __restrict__ rp, const int* __restrict__ col, int n, int* mark, int layer) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || mark[r] != layer) return;
    for (int p = rp[r]; p < rp[r + 1]; ++p) { const int c = col[p]; if (c < n && mark[c] == 0) mark[c] = layer + 1; }
}

void kb_asm_free(kb_pc_s* pc) {
    if (!pc || pc->kind != KB_PC_ASM || !pc->extra) return;
    KbAsm* a = reinterpret_cast<KbAsm*>(pc->extra);
    for (KbAsmBlock& b : a->blocks) {
        if (b.inner) kb_pc_destroy(b.inner);
        if (b.sub) kb_csr_destroy(b.sub);
        KB_FREE(b.idx); KB_FREE(b.r_blk); KB_FREE(b.x_blk);
    }
    delete a;
    pc->extra = nullptr;
}

int kb_asm_apply_dev(kb_pc_s* pc, const double* d_r, double* d_z, const KbCtl* sc, int sm) {
    kb_csr_s* A = pc->a;
    kb_ctx_s* c = A->ctx;
    KbAsm* a = reinterpret_cast<KbAsm*>(pc->extra);
    const long long n = (long long)A->n;
    { KbLaunch L(c, KB_K_SMALL); k_asm_zero<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_z, n, sc, sm); }
    for (KbAsmBlock& b : a->blocks) {
        if (b.n == 0) continue;
        const unsigned g = (unsigned)((b.n + 255) / 256);
        { KbLaunch L(c, KB_K_SMALL); k_asm_gather<<<g, 256, 0, c->stream>>>(d_r, b.idx, b.r_blk, b.n, sc, sm); }
        KB_TRY(kb_pc_apply_dev(b.inner, b.r_blk, b.x_blk, sc, sm));
        { KbLaunch L(c, KB_K_SMALL); k_asm_scatter_add<<<g, 256, 0, c->stream>>>(d_z, b.idx, b.x_blk, b.n, sc, sm); }
    }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}
int kb_asm_error(kb_pc_s* pc) {
    if (!pc || pc->kind != KB_PC_ASM || !pc->extra) return 0;
    int e = 0;
    for (KbAsmBlock& b : reinterpret_cast<KbAsm*>(pc->extra)->blocks) if (b.inner) e |= kb_ilu0_error(b.inner);
    return e;
}

extern "C" int kb_pc_create_asm(kb_csr A, uint64_t overlap, uint64_t nsub, const uint64_t* sub_ptr, const uint64_t* sub_idx, int inner, kb_pc* out) {
    *out = nullptr;
    if (!A) { kb_set_error("null operator"); return KB_SOLVE_ERROR; }
    if (A->dist) { kb_set_error("kb_pc_create_asm works on a single-GPU operator; on a shard kb_pc_create_ilu0 is the per-GPU block"); return KB_UNSUPPORTED; }
    if (A->n != A->ncols_global) { kb_set_error("additive Schwarz needs a square operator"); return KB_FACTOR_ERROR; }
    if (inner != KB_ASM_INNER_ILU0 && inner != KB_ASM_INNER_JACOBI) { kb_set_error("unknown inner solver kind"); return KB_UNSUPPORTED; }
    if (nsub == 0) nsub = 1;                                   // asm.rs:48: `capacity().max(1)`
    if (sub_ptr && !sub_idx && sub_ptr[nsub] > 0) { kb_set_error("null subdomain index array"); return KB_SOLVE_ERROR; }
    kb_ctx_s* c = A->ctx;
    KB_CUDA(cudaSetDevice(c->device));
    const uint64_t n = A->n;
    kb_pc_s* pc = new kb_pc_s;
    pc->a = A; pc->ctx = c; pc->kind = KB_PC_ASM;
    A->refs++;
    KbAsm* a = new KbAsm;
    a->overlap = overlap;
    pc->extra = a;
    int st = KB_OK;
    int* mark = nullptr;
    for (uint64_t b = 0; b < nsub && st == KB_OK; ++b) {
        std::vector<uint64_t> idx;
        if (sub_ptr) idx.assign(sub_idx + sub_ptr[b], sub_idx + sub_ptr[b + 1]);
        else {      // uniform chunks (asm.rs:46-57)
            uint64_t lo, hi;
            kb_partition_range(n, nsub, b, &lo, &hi);
            for (uint64_t i = lo; i < hi; ++i) idx.push_back(i);
        }
        for (uint64_t g : idx) if (g >= n) { kb_set_error("subdomain %llu: index %llu out of range", (unsigned long long)b, (unsigned long long)g); st = KB_SOLVE_ERROR; }
        if (st != KB_OK) break;
        {   // a block's rows must be distinct: its results are scatter-added in parallel
            std::vector<uint64_t> s(idx);
            std::sort(s.begin(), s.end());
            if (std::adjacent_find(s.begin(), s.end()) != s.end()) { kb_set_error("subdomain %llu lists a row twice", (unsigned long long)b); st = KB_UNSUPPORTED; break; }
        }
        if (overlap > 0 && !idx.empty()) {      // grow by `overlap` layers of graph neighbours, then ascending order
            if (!mark && (st = kb_alloc(&mark, (size_t)n + 1)) != KB_OK) break;
            std::vector<int> hm((size_t)n, 0);
            for (uint64_t g : idx) hm[(size_t)g] = 1;
            if (cudaMemcpyAsync(mark, hm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
            for (uint64_t l = 1; l <= overlap; ++l) {
                KbLaunch L(c, KB_K_OTHER);
                k_asm_grow<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(A->row_ptr, A->col, (int)n, mark, (int)l);
            }
            if (cudaMemcpyAsync(hm.data(), mark, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaStreamSynchronize(c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
            idx.clear();
            for (uint64_t g = 0; g < n; ++g) if (hm[(size_t)g]) idx.push_back(g);
        }
        a->blocks.emplace_back();
        KbAsmBlock& B = a->blocks.back();
        B.n = (int)idx.size();
        if (B.n == 0) continue;
        std::vector<int> hi32(idx.begin(), idx.end());
        if ((st = kb_alloc(&B.idx, (size_t)B.n)) != KB_OK || (st = kb_alloc(&B.r_blk, (size_t)B.n + 2)) != KB_OK || (st = kb_alloc(&B.x_blk, (size_t)B.n + 2)) != KB_OK) break;
        if (cudaMemcpyAsync(B.idx, hi32.data(), (size_t)B.n * sizeof(int), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { st = KB_SOLVE_ERROR; break; }
        if ((st = kb_csr_submatrix(A, idx.data(), idx.size(), &B.sub)) != KB_OK) break;
        st = inner == KB_ASM_INNER_ILU0 ? kb_pc_create_ilu0(B.sub, &B.inner) : kb_pc_create_jacobi(B.sub, &B.inner);
        if (st == KB_ZERO_PIVOT || st == KB_FACTOR_ERROR) {          // report the row in the caller's numbering
            const uint64_t local = B.inner ? B.inner->bad_row : 0;
            pc->bad_row = local < idx.size() ? idx[(size_t)local] : 0;
        }
    }
    if (mark) cudaFree(mark);
    if (st != KB_OK) {
        if (st == KB_ZERO_PIVOT || st == KB_FACTOR_ERROR) { *out = pc; return st; }     // caller may query kb_pc_bad_row, then destroy
        kb_pc_destroy(pc);
        return st;
    }
    *out = pc;
    return KB_OK;
}
// number of blocks and the (possibly overlap-grown) index list of block b, for parity tests
extern "C" uint64_t kb_pc_asm_num_blocks(kb_pc pc) {
    return (pc && pc->kind == KB_PC_ASM && pc->extra) ? reinterpret_cast<KbAsm*>(pc->extra)->blocks.size() : 0;
}
extern "C" uint64_t kb_pc_asm_block_size(kb_pc pc, uint64_t b) {
    if (!pc || pc->kind != KB_PC_ASM || !pc->extra) return 0;
    KbAsm* a = reinterpret_cast<KbAsm*>(pc->extra);
    return b < a->blocks.size() ? (uint64_t)a->blocks[(size_t)b].n : 0;
}
extern "C" int kb_pc_asm_block_indices(kb_pc pc, uint64_t b, uint64_t* out) {
    if (!pc || pc->kind != KB_PC_ASM || !pc->extra) { kb_set_error("not an additive-Schwarz preconditioner"); return KB_SOLVE_ERROR; }
    KbAsm* a = reinterpret_cast<KbAsm*>(pc->extra);
    if (b >= a->blocks.size()) { kb_set_error("block index out of range"); return KB_SOLVE_ERROR; }
    const KbAsmBlock& B = a->blocks[(size_t)b];
    if (B.n == 0) return KB_OK;
    KB_CUDA(cudaSetDevice(pc->ctx->device));
    std::vector<int> h((size_t)B.n);
    KB_CUDA(cudaMemcpy(h.data(), B.idx, (size_t)B.n * sizeof(int), cudaMemcpyDeviceToHost));
    for (int j = 0; j < B.n; ++j) out[j] = (uint64_t)h[(size_t)j];
    return KB_OK;
}
