/* kryst_b200.h — C ABI of the B200-native Krylov hot path that replaces kryst's CPU path.
 *
 * Drop-in boundary.  Every entry point cites the reference interface it
 * stands behind (paths relative to the kryst crate root, tmathis720/kryst 0.5.3):
 *
 *   MatVec<V>::matvec(&self,&V,&mut V)              src/core/traits.rs:4-7
 *   CsrMatrix::from_csr(nrows,ncols,row_ptr,col_idx,values)   src/matrix/sparse.rs:26-47
 *   InnerProduct<V>::{dot,norm}                      src/core/traits.rs:16-23, src/core/wrappers.rs:87-129
 *   Preconditioner<M,V>::{setup,apply}               src/preconditioner/mod.rs:8-13
 *   Jacobi<T>                                        src/preconditioner/jacobi.rs:26-95
 *   Ilu0<T>                                          src/preconditioner/ilu.rs:32-122
 *   AdditiveSchwarz (overlap 0, chunk partition)     src/preconditioner/asm.rs:34-57
 *   LinearSolver<M,V>::solve(&mut self,&A,pc,&b,&mut x) -> Result<SolveStats,KError>   src/solver/mod.rs:30-52
 *   PcgSolver / GmresSolver / BiCgStabSolver         src/solver/pcg.rs:114, gmres.rs:216, bicgstab.rs:69
 *   FgmresSolver::solve_flex                         src/solver/fgmres.rs:114-340
 *   SubmatrixExtract::submatrix                      src/matrix/sparse.rs:72-93 (used by asm.rs:58-65)
 *   SolveStats{iterations,final_residual,converged}  src/utils/convergence.rs:9-14
 *   KError                                           src/error.rs:6-19
 *   Comm{rank,size,barrier,scatter,gather,all_reduce,dot}   src/parallel/mod.rs:4-35 ; DistributedInnerProduct  src/core/wrappers.rs:134-156
 *   KspContext::solve_context / SolverKind / PC<T>   src/context/ksp_context.rs:25-148, src/context/pc_context.rs:36-76
 *
 * Conventions: plain pointers and sizes only; no C++/torch types.  Host pointers are borrowed
 * for the duration of the call.  All device memory, streams and communicators are owned by the
 * opaque handles.  A handle is not thread-safe (mirrors `&mut self`).  Calls are synchronous.
 * No exceptions cross the ABI: every call returns a kb_status; details via kb_last_error().
 * There is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 * KB_SOLVE_ERROR.
 */
#ifndef KRYST_B200_H
#define KRYST_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KB_ABI_VERSION 1

/* == KError (src/error.rs:6-19); 0 == Ok(..) */
typedef enum kb_status {
    KB_OK = 0,
    KB_FACTOR_ERROR = 1,        /* KError::FactorError(String)        */
    KB_SOLVE_ERROR = 2,         /* KError::SolveError(String); also CUDA/NCCL failures, bad arguments */
    KB_INDEFINITE_MATRIX = 3,   /* KError::IndefiniteMatrix           (pcg.rs:162-172) */
    KB_INDEFINITE_PC = 4,       /* KError::IndefinitePreconditioner   (pcg.rs:206-213) */
    KB_ZERO_PIVOT = 5,          /* KError::ZeroPivot(row) — row via kb_pc_bad_row()    */
    KB_UNSUPPORTED = 6          /* KError::Unsupported(&'static str)  */
} kb_status;

/* == SolveStats<f64> (src/utils/convergence.rs:9-14) + a breakdown code (0 = none).
 * BiCGStab breakdown `break`s return Ok with converged=false in the reference
 * (bicgstab.rs:117-119,291-292); `breakdown` says which test fired: 1 rho, 2 r^.v, 3 t.t, 4 omega. */
typedef struct kb_stats {
    uint64_t iterations;
    double final_residual;
    int32_t converged;
    int32_t breakdown;
} kb_stats;

typedef struct kb_ctx_s* kb_ctx;   /* one GPU + stream (+ communicator) */
typedef struct kb_csr_s* kb_csr;   /* device-resident CSR operator: MatVec + Indexing + MatShape */
typedef struct kb_pc_s* kb_pc;     /* device-resident preconditioner */

/* ---- context -------------------------------------------------------------------------- */
int kb_abi_version(void);
const char* kb_last_error(void);                       /* thread-local message of the last failure */
int kb_ctx_create(int device, kb_ctx* out);            /* binds one GPU; creates the library stream */
int kb_ctx_destroy(kb_ctx ctx);
void* kb_ctx_stream(kb_ctx ctx);                       /* the cudaStream_t every kernel is launched on */
int kb_ctx_device(kb_ctx ctx);
int kb_ctx_synchronize(kb_ctx ctx);
uint64_t kb_ctx_launch_count(kb_ctx ctx);              /* kernels launched by this library on ctx so far */

/* ---- src/parallel Comm surface (rank,size,barrier,all_reduce: parallel/mod.rs:4-35) ---- */
/* One process per GPU.  The 128-byte id is created on rank 0 and distributed by the host
 * program (torch.distributed / MPI / files); kb_comm_init is collective over all ranks.       */
int kb_comm_unique_id(void* id128);
int kb_comm_init(kb_ctx ctx, int rank, int size, const void* id128);
int kb_comm_rank(kb_ctx ctx);
int kb_comm_size(kb_ctx ctx);
int kb_comm_barrier(kb_ctx ctx);
int kb_comm_all_reduce(kb_ctx ctx, double local, double* global);   /* rank-ordered sum: deterministic */
/* Comm::dot (parallel/mod.rs:19-22) and DistributedInnerProduct::{dot,norm} (src/core/wrappers.rs:134-156): every rank
 * passes its slice; canonical-tree local part on the GPU, rank-ordered sum across ranks.  Collective.                   */
int kb_comm_dot(kb_ctx ctx, uint64_t n_local, const double* a, const double* b, double* out);
int kb_comm_norm(kb_ctx ctx, uint64_t n_local, const double* x, double* out);
/* Comm::scatter / Comm::gather (parallel/mod.rs:9-16, mpi_comm.rs:74-109): equal chunks of bytes_per_rank bytes; `global`
 * (scatter) and `out` (gather) are read / written on `root` only.  Host slices; collective.                             */
int kb_comm_scatter(kb_ctx ctx, const void* global, uint64_t bytes_per_rank, void* out, int root);
int kb_comm_gather(kb_ctx ctx, const void* local, uint64_t bytes_per_rank, void* out, int root);
/* uniform row-chunk partition, chunk = ceil(n/p) (src/preconditioner/asm.rs:46-57); host-only */
void kb_partition_range(uint64_t n, uint64_t p, uint64_t r, uint64_t* lo, uint64_t* hi);

/* ---- operator: CsrMatrix::from_csr (sparse.rs:26-47) + MatVec (traits.rs:4-7) ---------- */
/* Validates like new_checked (monotone row_ptr, in-range strictly ascending columns), narrows
 * indices to i32 on the device.  usize == uint64_t.                                          */
int kb_csr_create(kb_ctx ctx, uint64_t nrows, uint64_t ncols, const uint64_t* row_ptr,
                  const uint64_t* col_idx, const double* vals, kb_csr* out);
/* Row-block shard of a square n_global x n_global matrix: this rank owns rows [row_lo,row_hi)
 * (must equal kb_partition_range for its rank); col_idx are GLOBAL columns.  Builds the ghost
 * list and halo send lists on the device.  Collective.                                        */
int kb_csr_create_dist(kb_ctx ctx, uint64_t n_global, uint64_t row_lo, uint64_t row_hi,
                       const uint64_t* row_ptr, const uint64_t* col_idx, const double* vals, kb_csr* out);
int kb_csr_destroy(kb_csr a);
uint64_t kb_csr_nrows(kb_csr a);                       /* Indexing::nrows / MatShape::nrows (owned rows) */
uint64_t kb_csr_ncols(kb_csr a);                       /* MatShape::ncols (global)                        */
uint64_t kb_csr_nnz(kb_csr a);
int kb_csr_matvec(kb_csr a, const double* x, double* y);            /* host slices: H2D, kernel, D2H   */
int kb_csr_matvec_device(kb_csr a, const double* d_x, double* d_y); /* device pointers, no copies      */
/* partition maps for parity tests (integers, bit-exact vs the oracle) */
uint64_t kb_csr_num_ghosts(kb_csr a);
int kb_csr_get_ghosts(kb_csr a, uint64_t* ghosts_global);
int kb_csr_spmv_kernel_kind(kb_csr a);                 /* 0 = CSR-stream (thread/row from smem), 1 = vector-per-row, 2 = bulk-async staged */
int kb_csr_spmv_x_staged(kb_csr a);                    /* kind 2 only: non-zero when x tiles are staged in shared memory too (stage geometry + 1) */
/* SubmatrixExtract::submatrix(&self, indices) (src/core/traits.rs; impl src/matrix/sparse.rs:72-93), the call
 * AdditiveSchwarz::setup makes per subdomain (src/preconditioner/asm.rs:58-65): out[i][j] = a[indices[i]][indices[j]]
 * for any index order (repeats allowed), stored zeros dropped, built on the device.  Single-GPU operators only.   */
int kb_csr_submatrix(kb_csr a, const uint64_t* indices, uint64_t k, kb_csr* out);
/* read the operator back as the arrays CsrMatrix::from_csr takes (sparse.rs:26-34): row_ptr[nrows+1], col_idx[nnz], vals[nnz] */
int kb_csr_download(kb_csr a, uint64_t* row_ptr, uint64_t* col_idx, double* vals);

/* ---- InnerProduct (wrappers.rs:90-128): canonical-tree dot / norm on device ------------ */
int kb_dot(kb_ctx ctx, uint64_t n, const double* x, const double* y, double* out);   /* host slices */
int kb_norm(kb_ctx ctx, uint64_t n, const double* x, double* out);

/* ---- preconditioners: Preconditioner::setup (create) / apply (mod.rs:8-13) ------------- */
int kb_pc_create_jacobi(kb_csr a, kb_pc* out);         /* jacobi.rs:53-73 values, direct diagonal read */
int kb_pc_create_ilu0(kb_csr a, kb_pc* out);           /* textbook ILU(0); on a shard: block-Jacobi ILU(0) */
/* AdditiveSchwarz::new(overlap, subdomains) + setup (src/preconditioner/asm.rs:34-65), apply = asm.rs:76-116: z = sum over
 * blocks, in order, of R_b^T inner(A_b, R_b r), A_b = a.submatrix(block rows).  sub_ptr == NULL: nsub uniform row chunks
 * (asm.rs:46-57); else block b = sub_idx[sub_ptr[b] .. sub_ptr[b+1]) in the caller's order (rows of a block distinct).
 * inner: one application of the block's ILU(0) (block-Jacobi ILU(0) for disjoint blocks) or of its Jacobi.  overlap = 0 is
 * the reference (which stores `overlap` and never reads it); overlap = k grows each set by k layers of graph neighbours
 * (PETSc PCASM) and orders it ascending.  Single-GPU operators; on a shard kb_pc_create_ilu0 is the per-GPU block.      */
enum { KB_ASM_INNER_ILU0 = 0, KB_ASM_INNER_JACOBI = 1 };
int kb_pc_create_asm(kb_csr a, uint64_t overlap, uint64_t nsub, const uint64_t* sub_ptr, const uint64_t* sub_idx,
                     int inner, kb_pc* out);
uint64_t kb_pc_asm_num_blocks(kb_pc pc);
uint64_t kb_pc_asm_block_size(kb_pc pc, uint64_t b);
int kb_pc_asm_block_indices(kb_pc pc, uint64_t b, uint64_t* out);   /* after overlap growth */
int kb_pc_apply(kb_pc pc, const double* r, double* z); /* host slices */
int kb_pc_apply_device(kb_pc pc, const double* d_r, double* d_z);
int kb_pc_destroy(kb_pc pc);
uint64_t kb_pc_bad_row(kb_pc pc);                      /* row of ZeroPivot / missing diagonal */
int kb_pc_get_inv_diag(kb_pc pc, double* out);         /* Jacobi::inv_diag ; ILU: 1/u_ii */
int kb_pc_ilu0_get_factors(kb_pc pc, double* lu, uint64_t* diag_ptr);
int kb_pc_ilu0_get_levels(kb_pc pc, int upper, uint64_t* nlevels, uint64_t* level_ptr /*n+1*/,
                          uint64_t* order /*n*/);

/* ---- solvers: LinearSolver::solve (solver/mod.rs:43-49) -------------------------------- */
#define KB_FLAG_DEVICE_PTRS 1u    /* b and x are device pointers on ctx's GPU (stay resident)     */
#define KB_FLAG_TEXTBOOK    2u    /* BiCGStab: Tier-T relative-tolerance, preconditioned variant  */
#define KB_FLAG_PROFILE     4u    /* time every kernel class with CUDA events (no graph replay)   */
#define KB_FLAG_NO_GRAPH    8u    /* plain launches instead of CUDA-graph replay                   */
#define KB_FLAG_SINGLE_REDUCTION 16u /* PCG: Chronopoulos-Gear recurrences, ONE fused reduction (one all-reduce on
                                        shards) per iteration - what pcg.rs:36-37's flag is named after; SURVEY 8(f3) */

#define KB_FLAG_BLOCK_ORTH 128u   /* GMRES: block orthogonalisation (the idea of src/solver/pca_gmres.rs:172-229): ONE classical
                                   * Gram-Schmidt pass whose inner products {V^T w, w.w} are reduced together (one all-reduce per
                                   * Arnoldi step instead of three), h_{j+1,j}^2 = w.w - sum h^2; 2 sweeps over the basis instead
                                   * of 3.  Opt-in extension: CGS2 (gmres.rs:65-105 semantics) stays the default.                */
#define KB_FLAG_PIPELINED 256u    /* PCG: pipelined recurrences (Ghysels-Vanroose): u = M^-1 r and w = A u are carried by recurrences, so the
                                   * ONE reduction of an iteration {r.u, w.u, norm} is sent before the iteration's SpMV and received after
                                   * it (the all-reduce overlaps the SpMV on shards).  Jacobi / no preconditioner.  Opt-in extension.     */
#define KB_FLAG_HISTORY 32u       /* record the per-iteration residuals on the device; fetch with kb_get_history        */
#define KB_FLAG_MONITOR 64u       /* slow mode: one iteration (GMRES family: one restart cycle) per launch batch,
                                     the observer set with kb_set_monitor runs on the host for every new history entry */
/* monitor: Option<Box<dyn FnMut(usize, T)>> (pcg.rs:43,81-86,143-145,196-198; fgmres.rs:45-46,92-97,286-289) */
typedef void (*kb_monitor_fn)(uint64_t iteration, double residual, void* user);
int kb_set_monitor(kb_csr a, kb_monitor_fn fn, void* user);        /* fn == NULL clears; used by solves with KB_FLAG_MONITOR */
/* residual_history (pcg.rs:45,146,199; fgmres.rs:48,290) of the last solve on `a` that ran with KB_FLAG_HISTORY or
 * KB_FLAG_MONITOR: PCG = initial norm + one entry per iteration; GMRES / FGMRES = |g[j+1]| per inner iteration;
 * BiCGStab = ||r|| per iteration (the reference keeps no history for those two).  len = entries produced.               */
int kb_get_history(kb_csr a, double* out, uint64_t cap, uint64_t* len);

/* CgNormType (pcg.rs:25) */
enum { KB_NORM_PRECONDITIONED = 0, KB_NORM_UNPRECONDITIONED = 1, KB_NORM_NATURAL = 2, KB_NORM_NONE = 3 };
/* Preconditioning (gmres.rs:28-32) */
enum { KB_SIDE_NONE = 0, KB_SIDE_LEFT = 1, KB_SIDE_RIGHT = 2 };

/* PcgSolver::new(tol,max_iters).with_norm(..).solve (pcg.rs:50-90,114-222).  history receives the
 * residual_history pushes (pcg.rs:146,199); x is written iff the call returns KB_OK.
 * KB_FLAG_SINGLE_REDUCTION (Jacobi / no preconditioner): same conventions, r.u, (Au).u and the norm reduced together;
 * not the reference's arithmetic (its flag is a no-op there) - checked against the oracle's restatement of the variant. */
int kb_pcg_solve(kb_csr a, kb_pc pc, const double* b, double* x, double tol, uint64_t max_iters,
                 int norm_type, uint32_t flags, double* history, uint64_t hist_cap, uint64_t* hist_len,
                 kb_stats* stats);
/* GmresSolver::new(restart,tol,max_iters).with_preconditioning(side).solve (gmres.rs:49-60,216-402);
 * orthogonalisation is CGS2 as a block GEMV (Tier T formulation for Left/Right).     */
int kb_gmres_solve(kb_csr a, kb_pc pc, const double* b, double* x, uint64_t restart, double tol,
                   uint64_t max_iters, int side, uint32_t flags, kb_stats* stats);
/* FgmresSolver::new(tol,max_iters,restart).solve_flex(&a, pc, &b, &mut x) (src/solver/fgmres.rs:52-340):
 * flexible right-preconditioned GMRES, one classical Gram-Schmidt pass, z_j = M^-1 v_j kept; the reference's quirks
 * are reproduced (inner test relative to the current cycle's residual, absolute cycle test, final_residual = ||r0||). */
int kb_fgmres_solve(kb_csr a, kb_pc pc, const double* b, double* x, uint64_t restart, double tol,
                    uint64_t max_iters, uint32_t flags, kb_stats* stats);
/* BiCgStabSolver::new(tol,max_iters).solve (bicgstab.rs:45-47,69-293).  Default = literal
 * (pc ignored, absolute tol); KB_FLAG_TEXTBOOK = right-preconditioned, relative tol.             */
int kb_bicgstab_solve(kb_csr a, kb_pc pc, const double* b, double* x, double tol, uint64_t max_iters,
                      uint32_t flags, kb_stats* stats);

/* ---- the caller above the path: PC<T> factory and KspContext::solve_context ------------------------------ */
/* PC<T> (src/context/pc_context.rs:36-76), same variant order */
enum { KB_PCK_JACOBI = 0, KB_PCK_SSOR = 1, KB_PCK_ILU0 = 2, KB_PCK_ILUP = 3, KB_PCK_ILUT = 4, KB_PCK_CHEBYSHEV = 5, KB_PCK_APPROXINV = 6,
       KB_PCK_BLOCK_JACOBI = 7, KB_PCK_MULTICOLOR = 8, KB_PCK_AMG = 9, KB_PCK_ADDITIVE_SCHWARZ = 10 };
typedef struct kb_pc_spec {
    int32_t kind;                 /* KB_PCK_* */
    uint64_t fill;                /* Ilup { fill } / Ilut { fill, .. } */
    double droptol;               /* Ilut { droptol } */
    uint64_t overlap;             /* AdditiveSchwarz overlap layers (0 = the reference) */
    uint64_t nblocks;             /* BlockJacobi { blocks } / AdditiveSchwarz subdomains (uniform chunks when block_ptr is NULL) */
    const uint64_t* block_ptr;    /* [nblocks + 1] */
    const uint64_t* block_idx;
} kb_pc_spec;
/* Device kinds: Jacobi, Ilu0, Ilup{fill:0}, BlockJacobi{blocks}, AdditiveSchwarz; the rest -> KB_UNSUPPORTED.           */
int kb_pc_create_from_spec(kb_csr a, const kb_pc_spec* spec, kb_pc* out);
/* SolverKind (src/context/ksp_context.rs:25-48), same variant order */
enum { KB_KSP_CG = 0, KB_KSP_PCG = 1, KB_KSP_GMRES_LEFT = 2, KB_KSP_GMRES_RIGHT = 3, KB_KSP_FGMRES = 4, KB_KSP_BICGSTAB = 5,
       KB_KSP_CGS = 6, KB_KSP_QMR = 7, KB_KSP_TFQMR = 8, KB_KSP_MINRES = 9, KB_KSP_CGNR = 10 };
/* KspContext{kind, tol, max_it, restart} (ksp_context.rs:54-69); `a` and `pc` are passed alongside */
typedef struct kb_ksp { int32_t kind; double tol; uint64_t max_it; uint64_t restart; } kb_ksp;
/* KspContext::solve_context(&mut self, b, x, comm) (ksp_context.rs:88-148): builds the solver named by `kind` with
 * (tol, max_it[, restart]) and forwards to its solve; Cg ignores pc (cg.rs:114-115); Fgmres takes pc as the flexible
 * preconditioner; a row-partitioned operator carries its communicator (the reference accepts `comm` and never uses it). */
int kb_ksp_solve(kb_csr a, kb_pc pc, const kb_ksp* ksp, const double* b, double* x, uint32_t flags, kb_stats* stats);

/* ---- measurement hooks (bench.py) ------------------------------------------------------ */
#define KB_PROF_CLASSES 12
typedef struct kb_profile {
    uint64_t launches[KB_PROF_CLASSES];
    double ms[KB_PROF_CLASSES];          /* summed CUDA-event time per class, on ctx's stream */
} kb_profile;
/* classes */
enum { KB_K_SPMV = 0, KB_K_PCG_UPDATE = 1, KB_K_XPAY = 2, KB_K_INIT = 3, KB_K_BICG = 4, KB_K_GS_DOT = 5,
       KB_K_GS_UPDATE = 6, KB_K_TRSV = 7, KB_K_SMALL = 8, KB_K_HALO = 9, KB_K_ALLREDUCE = 10, KB_K_OTHER = 11 };
int kb_profile_reset(kb_ctx ctx);
int kb_profile_get(kb_ctx ctx, kb_profile* out);
const char* kb_profile_class_name(int cls);

#ifdef __cplusplus
}
#endif
#endif /* KRYST_B200_H */
