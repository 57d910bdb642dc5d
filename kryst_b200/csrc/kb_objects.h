// kb_objects.h — host-side handle layouts (opaque to ABI users).
#pragma once
#include "kb_internal.cuh"

struct KbPcgWs;
struct KbBicgWs;
struct KbGmresWs;
struct KbHalo;

struct kb_csr_s {
    kb_ctx_s* ctx = nullptr;
    int refs = 1;                // the user handle + one per preconditioner set up for this operator
    uint64_t n = 0;              // owned rows (Indexing::nrows)
    uint64_t ncols_global = 0;   // MatShape::ncols
    uint64_t ncols_local = 0;    // owned + ghost columns: length of the x operand of the kernels
    uint64_t nnz = 0;
    int* row_ptr = nullptr;      // i32 after checked narrowing
    int* col = nullptr;          // local column ids, stored order == ascending GLOBAL column
    double* vals = nullptr;
    int ntiles = 0;
    int kind = 0;                // 0 CSR-stream (plain loads), 1 vector-per-row, 2 bulk-async staged CSR-stream
    int* tile_chunk = nullptr;   // kind 2: chunk table (see kb_spmv_bulk.cuh)
    int* chunk_row = nullptr;
    int* chunk_nz = nullptr;
    int nchunks = 0;
    bool prod = false;           // kind 2: per-nonzero product phase (long rows) instead of per-row gathers
    // kind 2 with the x operand staged in shared memory (kb_spmv_xtile.cuh): 0 = off, else configuration + 1
    int xt = 0;
    bool xt_prod = false;        // staged x: per-nonzero product phase (rows longer than 32 on average) instead of a thread per row
    int* xt_tile_chunk = nullptr;
    int* xt_chunk_row = nullptr;
    int* xt_chunk_nz = nullptr;
    int* xt_lo = nullptr;        // [xt_nchunks * 16] x intervals of each chunk
    int* xt_len = nullptr;
    int* xt_tail = nullptr;
    unsigned short* xt_lcol = nullptr;   // [nnz + 16] chunk-local 16-bit column ids
    int xt_nchunks = 0;
    int vec = 8;                 // sub-warp width of the vector kernel
    uint64_t max_row_len = 0;
    uint64_t hist[6] = {0, 0, 0, 0, 0, 0};   // row-length histogram: <=8,<=16,<=32,<=64,<=128,>128
    // row-block shard (src/parallel gains this)
    bool dist = false;
    uint64_t n_global = 0, row_lo = 0, row_hi = 0, nghost = 0;
    uint64_t* ghosts = nullptr;  // device: sorted unique global ids of ghost columns
    KbHalo* halo = nullptr;
    int* tiles_interior = nullptr;   // tiles without ghost columns / with ghost columns (device lists)
    int* tiles_boundary = nullptr;
    int n_interior = 0, n_boundary = 0;
    int* tiles_order = nullptr;      // interior tiles, then boundary tiles: the single-launch SpMV walks it so that only a CTA's LAST tiles wait for the halo
    // scratch for host-slice matvec
    double* x_tmp = nullptr;
    double* y_tmp = nullptr;
    // cached solver workspaces (device vectors, control block, CUDA graphs)
    KbPcgWs* pcg_ws = nullptr;
    KbBicgWs* bicg_ws = nullptr;
    KbGmresWs* gmres_ws = nullptr;
    // observability: host observer + device residual history of the last KB_FLAG_HISTORY / KB_FLAG_MONITOR solve
    kb_monitor_fn monitor = nullptr;
    void* monitor_user = nullptr;
    double* hist_buf = nullptr;
    uint64_t hist_cap = 0, hist_len = 0;
};

enum { KB_PC_JACOBI = 1, KB_PC_ILU0 = 2, KB_PC_ASM = 3 };
uint64_t kb_next_serial();
static inline uint64_t kb_pc_serial(const struct kb_pc_s* pc);

struct kb_pc_s {
    kb_csr_s* a = nullptr;           // operator it was set up for (must outlive every apply/solve using this pc)
    kb_ctx_s* ctx = nullptr;         // kept separately so that destroy is safe after the operator is gone
    int kind = 0;
    double* inv_diag = nullptr;       // Jacobi: 1/a_ii (0 if a_ii == 0); ILU(0): 1/u_ii
    // ILU(0) on the owned diagonal block (pattern = A restricted to owned columns)
    int* l_rp = nullptr;              // block-local CSR (ghost couplings dropped)
    int* l_col = nullptr;
    double* lu = nullptr;
    int* diag_ptr = nullptr;
    uint64_t l_nnz = 0;
    int nlev[2] = {0, 0};             // lower / upper level counts
    int* level_ptr[2] = {nullptr, nullptr};
    int* order[2] = {nullptr, nullptr};     // rows per level, ascending inside a level
    int* sched[2] = {nullptr, nullptr};     // warp-padded execution order for the sync-free solves
    int sched_len[2] = {0, 0};
    double* tmp = nullptr;            // y of L y = r
    uint64_t bad_row = 0;
    double* r_tmp = nullptr;
    double* z_tmp = nullptr;
    void* extra = nullptr;            // ILU(0) bookkeeping (kb_ilu0.cu)
    // Unique for the lifetime of the process.  The solvers' CUDA-graph caches are keyed on it, never on the
    // handle's address: a new preconditioner can be allocated where a destroyed one lived.
    uint64_t serial = kb_next_serial();
};

static inline uint64_t kb_pc_serial(const kb_pc_s* pc) { return pc ? pc->serial : 0; }

// allocation helpers
template <class T>
static inline int kb_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    KB_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    return KB_OK;
}
#define KB_FREE(p)            \
    do {                      \
        if (p) cudaFree(p);   \
        p = nullptr;          \
    } while (0)

void kb_ctx_unref(kb_ctx_s* c);
void kb_csr_unref(kb_csr_s* A);

// ---- internal launch API (implemented across the .cu files) ---------------------------------
int kb_csr_spmv_plain(kb_csr_s* A, const double* d_x, double* d_y);
int kb_halo_exchange(kb_csr_s* A, double* d_x);   // fills ghost entries of d_x (no-op when !dist)
int kb_halo_begin(kb_csr_s* A, double* d_x);      // start the exchange ...
int kb_halo_end(kb_csr_s* A, double* d_x);        // ... ghost tail valid after this
int kb_allreduce_slots(kb_ctx_s* c, double* d_vals, int count);   // rank-ordered sum, in place
int kb_p2p_error(kb_ctx_s* c);                                    // 1 if a peer-memory spin timed out
// z = M^-1 r on the device; kernels are no-ops when skip_ctl->done (or, per skip_mask, early / cycle_break) is set
int kb_pc_apply_dev(kb_pc_s* pc, const double* d_r, double* d_z, const KbCtl* skip_ctl = nullptr, int skip_mask = 0);
void kb_pcg_ws_free(KbPcgWs* w);
void kb_bicg_ws_free(KbBicgWs* w);
void kb_gmres_ws_free(KbGmresWs* w);
void kb_halo_free(KbHalo* h);
int kb_ilu0_build(kb_pc_s* pc);
int kb_ilu0_error(kb_pc_s* pc);                   // 1 if a triangular-solve spin timed out
void kb_ilu0_free(kb_pc_s* pc);
int kb_asm_apply_dev(kb_pc_s* pc, const double* d_r, double* d_z, const KbCtl* skip_ctl, int skip_mask);   // kb_asm.cu
int kb_asm_error(kb_pc_s* pc);
void kb_asm_free(kb_pc_s* pc);
int kb_upload_or_alias(kb_ctx_s* c, const double* src, double* dst, uint64_t n, bool device_ptrs);
// history buffer of the operator (grown on demand, capped at 4 Mi entries); sets h->hist / h->hist_cap when the flags ask for it
int kb_hist_prepare(kb_csr_s* A, uint32_t flags, uint64_t max_entries, KbCtl* h);
