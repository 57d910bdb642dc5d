// kb_context.cu — the one caller above the hot path, in the C ABI:
//   PC<T> (src/context/pc_context.rs:36-76)            -> kb_pc_create_from_spec : builds the device preconditioner
//   SolverKind + KspContext::solve_context             -> kb_ksp_solve           : constructs the solver with
//   (src/context/ksp_context.rs:25-69,88-148)             (tol, max_it[, restart]) and forwards to its solve
// Pure dispatch, like the reference.  Kinds that are not on the device hot path answer KB_UNSUPPORTED
// (KError::Unsupported), never a CPU fallback.
#include "kb_objects.h"

extern "C" int kb_pc_create_from_spec(kb_csr a, const kb_pc_spec* s, kb_pc* out) {
    if (out) *out = nullptr;
    if (!a || !s || !out) { kb_set_error("kb_pc_create_from_spec: null argument"); return KB_SOLVE_ERROR; }
    switch (s->kind) {
    case KB_PCK_JACOBI: return kb_pc_create_jacobi(a, out);
    case KB_PCK_ILU0: return kb_pc_create_ilu0(a, out);
    case KB_PCK_ILUP:               // Ilup { fill }: level 0 is ILU(0) (ilup.rs:84-98 with fill = 0 keeps A's pattern)
        if (s->fill == 0) return kb_pc_create_ilu0(a, out);
        kb_set_error("PC::Ilup with fill > 0 is not on the device hot path");
        return KB_UNSUPPORTED;
    case KB_PCK_BLOCK_JACOBI:       // BlockJacobi { blocks }: one direct solve per diagonal block (block_jacobi.rs:39-107);
                                    // here the block "solve" is its ILU(0) (exact LU for blocks without fill), overlap 0
        if (!s->block_ptr || s->nblocks == 0) { kb_set_error("PC::BlockJacobi needs its blocks"); return KB_SOLVE_ERROR; }
        return kb_pc_create_asm(a, 0, s->nblocks, s->block_ptr, s->block_idx, KB_ASM_INNER_ILU0, out);
    case KB_PCK_ADDITIVE_SCHWARZ:   // AdditiveSchwarz: asm.rs:34-65 (uniform chunks unless blocks are given)
        return kb_pc_create_asm(a, s->overlap, s->nblocks, s->block_ptr, s->block_idx, KB_ASM_INNER_ILU0, out);
    case KB_PCK_SSOR: case KB_PCK_ILUT: case KB_PCK_CHEBYSHEV: case KB_PCK_APPROXINV: case KB_PCK_MULTICOLOR: case KB_PCK_AMG:
        kb_set_error("this PC variant is not on the device hot path (SURVEY 2: out of scope)");
        return KB_UNSUPPORTED;
    default:
        kb_set_error("unknown PC kind %d", s->kind);
        return KB_SOLVE_ERROR;
    }
}

extern "C" int kb_ksp_solve(kb_csr a, kb_pc pc, const kb_ksp* k, const double* b, double* x, uint32_t flags, kb_stats* stats) {
    if (!k) { kb_set_error("kb_ksp_solve: null context"); return KB_SOLVE_ERROR; }
    switch (k->kind) {
    case KB_KSP_GMRES_LEFT: return kb_gmres_solve(a, pc, b, x, k->restart, k->tol, k->max_it, KB_SIDE_LEFT, flags, stats);     // ksp_context.rs:90-94
    case KB_KSP_GMRES_RIGHT: return kb_gmres_solve(a, pc, b, x, k->restart, k->tol, k->max_it, KB_SIDE_RIGHT, flags, stats);   // :95-99
    case KB_KSP_FGMRES: return kb_fgmres_solve(a, pc, b, x, k->restart, k->tol, k->max_it, flags, stats);                      // :101-107
    case KB_KSP_CG:             // CgSolver ignores the preconditioner (`let _ = pc;`, cg.rs:114-115)
        return kb_pcg_solve(a, nullptr, b, x, k->tol, k->max_it, KB_NORM_UNPRECONDITIONED, flags, nullptr, 0, nullptr, stats);
    case KB_KSP_PCG: return kb_pcg_solve(a, pc, b, x, k->tol, k->max_it, KB_NORM_UNPRECONDITIONED, flags, nullptr, 0, nullptr, stats);   // :113-117
    case KB_KSP_BICGSTAB: return kb_bicgstab_solve(a, pc, b, x, k->tol, k->max_it, flags, stats);                              // :118-122
    case KB_KSP_CGS: case KB_KSP_QMR: case KB_KSP_TFQMR: case KB_KSP_MINRES: case KB_KSP_CGNR:
        kb_set_error("this SolverKind is not on the device hot path (SURVEY 2: out of scope)");
        return KB_UNSUPPORTED;
    default:
        kb_set_error("unknown SolverKind %d", k->kind);
        return KB_SOLVE_ERROR;
    }
}
