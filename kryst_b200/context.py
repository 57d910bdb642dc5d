"""KspContext / SolverKind dispatch: the one caller above the hot path.

Mirror of src/context/ksp_context.rs:25-148: public fields `kind, a, pc, tol, max_it, restart` and
`solve_context(b, x, comm)`, which builds the solver with (tol, max_it[, restart]) and forwards to `solve`.
The kinds on the north-star path (and FGMRES) run on the device; the others (CGS, QMR, TFQMR, MINRES, CGNR) are out of
scope of this build and raise `Unsupported`.  Unlike the reference (ksp_context.rs:88 accepts `comm` and never uses
it), a row-partitioned operator already carries its communicator, so `comm` is only checked for consistency.
"""
import enum

from .api import BiCgStabSolver, FgmresSolver, GmresSolver, PcgSolver, Preconditioning, Unsupported


class SolverKind(enum.Enum):           # ksp_context.rs:25-48
    Cg = "Cg"
    Pcg = "Pcg"
    GmresLeft = "GmresLeft"
    GmresRight = "GmresRight"
    Fgmres = "Fgmres"
    Bicgstab = "Bicgstab"
    Cgs = "Cgs"
    Qmr = "Qmr"
    Tfqmr = "Tfqmr"
    Minres = "Minres"
    Cgnr = "Cgnr"


class KspContext:
    """KspContext{kind, a, pc, flex_pc, tol, max_it, restart} (ksp_context.rs:54-69); fields are public."""

    def __init__(self, kind, a, pc=None, tol=1e-8, max_it=1000, restart=30, flex_pc=None):
        self.kind, self.a, self.pc, self.flex_pc = SolverKind(kind), a, pc, flex_pc
        self.tol, self.max_it, self.restart = tol, max_it, restart

    def solve_context(self, b, x, comm=None):
        """ksp_context.rs:88-148."""
        if comm is not None and hasattr(comm, "size") and comm.size() != self.a.ctx.size():
            raise Unsupported("comm does not match the communicator the operator was partitioned with")
        k = self.kind
        if k is SolverKind.GmresLeft:
            return GmresSolver(self.restart, self.tol, self.max_it).with_preconditioning(Preconditioning.Left).solve(self.a, self.pc, b, x)
        if k is SolverKind.GmresRight:
            return GmresSolver(self.restart, self.tol, self.max_it).with_preconditioning(Preconditioning.Right).solve(self.a, self.pc, b, x)
        if k is SolverKind.Pcg:
            return PcgSolver(self.tol, self.max_it).solve(self.a, self.pc, b, x)
        if k is SolverKind.Cg:
            # CgSolver ignores the preconditioner (`let _ = pc;`, src/solver/cg.rs:114-115)
            return PcgSolver(self.tol, self.max_it).solve(self.a, None, b, x)
        if k is SolverKind.Bicgstab:
            return BiCgStabSolver(self.tol, self.max_it).solve(self.a, self.pc, b, x)
        if k is SolverKind.Fgmres:
            return FgmresSolver(self.tol, self.max_it, self.restart).solve_flex(self.a, self.flex_pc, b, x)
        raise Unsupported("SolverKind::%s is not on the device hot path" % k.value)
