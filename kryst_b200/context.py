"""KspContext / SolverKind dispatch: the one caller above the hot path.

Mirror of src/context/ksp_context.rs:25-148: public fields `kind, a, pc, tol, max_it, restart` and
`solve_context(b, x, comm)`, which builds the solver with (tol, max_it[, restart]) and forwards to `solve`.
The kinds on the north-star path (and FGMRES) run on the device; the others (CGS, QMR, TFQMR, MINRES, CGNR) are out of
scope of this build and raise `Unsupported`.  Unlike the reference (ksp_context.rs:88 accepts `comm` and never uses
it), a row-partitioned operator already carries its communicator, so `comm` is only checked for consistency.
"""
import enum

from .api import PC, Unsupported, ksp_solve  # noqa: F401  (PC: pc_context.rs:36-76 mirror + factory)


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
        # the dispatch itself lives behind the C ABI (kb_ksp_solve, csrc/kb_context.cu), so the Rust / C++ hosts share it;
        # kinds that are not on the device hot path come back as KError::Unsupported
        order = [k for k in SolverKind]
        pc = self.flex_pc if self.kind is SolverKind.Fgmres else self.pc
        return ksp_solve(order.index(self.kind), self.a, pc, self.tol, self.max_it, self.restart, b, x)
