// C++ host-layer tests written like the reference's own tests (tests/preconditioner_integration.rs,
// src/solver/{pcg,gmres,bicgstab}.rs test modules), through include/kryst_b200.hpp -> C ABI -> CUDA kernels.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "kryst_b200.hpp"
using namespace kryst;

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

struct Csr { size_t n; std::vector<uint64_t> rp, ci; std::vector<double> v; };
static Csr tridiag(size_t n, double lo, double d, double up) {
    Csr m; m.n = n; m.rp.push_back(0);
    for (size_t i = 0; i < n; ++i) {
        if (i > 0) { m.ci.push_back(i - 1); m.v.push_back(lo); }
        m.ci.push_back(i); m.v.push_back(d);
        if (i + 1 < n) { m.ci.push_back(i + 1); m.v.push_back(up); }
        m.rp.push_back(m.ci.size());
    }
    return m;
}
static double rel_error(const std::vector<double>& x, double t) {
    double num = 0, den = 0;
    for (double xi : x) { num += (xi - t) * (xi - t); den += t * t; }
    return std::sqrt(num / den);
}

int main() {
    Context ctx(0);
    {   // spd_jacobi_pcg_converges (tests/preconditioner_integration.rs:127-138)
        const size_t n = 10;
        Csr m = tridiag(n, -1, 2, -1);
        DeviceCsr a = DeviceCsr::from_csr(ctx, n, n, m.rp, m.ci, m.v);
        std::vector<double> ones(n, 1.0), b(n), x(n, 0.0);
        a.matvec(ones, b);
        Jacobi pc; pc.setup(a);
        PcgSolver solver(1e-12, n);
        SolveStats st = solver.solve(a, &pc, b, x);
        REQUIRE(st.converged); REQUIRE(rel_error(x, 1.0) < 1e-10); REQUIRE(st.iterations <= n);
        REQUIRE(a.nrows() == n && a.ncols() == n);
    }
    {   // nonsym GMRES(10), none and left+ILU(0) (tests/preconditioner_integration.rs:156-179)
        const size_t n = 10;
        Csr m = tridiag(n, -1, 2, 0.5);
        DeviceCsr a = DeviceCsr::from_csr(ctx, n, n, m.rp, m.ci, m.v);
        std::vector<double> ones(n, 1.0), b(n), x(n, 0.0);
        a.matvec(ones, b);
        GmresSolver s1(10, 1e-12, 100);
        SolveStats st = s1.solve(a, nullptr, b, x);
        REQUIRE(st.converged); REQUIRE(rel_error(x, 1.0) < 1e-10);
        Ilu0 ilu; ilu.setup(a);
        std::fill(x.begin(), x.end(), 0.0);
        GmresSolver s2 = GmresSolver(10, 1e-12, 100).with_preconditioning(Preconditioning::Left);
        st = s2.solve(a, &ilu, b, x);
        REQUIRE(st.converged); REQUIRE(rel_error(x, 1.0) < 1e-10);
    }
    {   // fgmres_equiv_to_gmres_on_fixed_pc (src/solver/fgmres.rs:531-551)
        Csr m; m.n = 2; m.rp = {0, 2, 4}; m.ci = {0, 1, 0, 1}; m.v = {2.0, 1.0, 1.0, 3.0};
        DeviceCsr a = DeviceCsr::from_csr(ctx, 2, 2, m.rp, m.ci, m.v);
        std::vector<double> xt = {1.0, 2.0}, b(2), x(2, 0.0);
        a.matvec(xt, b);
        Jacobi pc; pc.setup(a);
        FgmresSolver solver(1e-10, 100, 25);
        SolveStats st = solver.solve_flex(a, &pc, b, x);
        REQUIRE(st.converged);
        for (int i = 0; i < 2; ++i) REQUIRE(std::fabs(x[i] - xt[i]) < 1e-6);
    }
    {   // SubmatrixExtract::submatrix (src/matrix/sparse.rs:72-93) + read-back; fused single-reduction PCG
        Csr m; m.n = 4; m.rp = {0, 2, 5, 8, 10}; m.ci = {0, 1, 0, 1, 2, 1, 2, 3, 2, 3};
        m.v = {4, -1, -1, 4, -1, -1, 4, -1, -1, 4};
        DeviceCsr a = DeviceCsr::from_csr(ctx, 4, 4, m.rp, m.ci, m.v);
        DeviceCsr s = a.submatrix({3, 1, 2});
        std::vector<uint64_t> rp, ci; std::vector<double> v;
        s.to_csr(rp, ci, v);
        REQUIRE((rp == std::vector<uint64_t>{0, 2, 4, 7}));
        REQUIRE((ci == std::vector<uint64_t>{0, 2, 1, 2, 0, 1, 2}));
        REQUIRE((v == std::vector<double>{4, -1, 4, -1, -1, -1, 4}));
        std::vector<double> xt = {1, 2, 3, 4}, b(4), x(4, 0.0);
        a.matvec(xt, b);
        Jacobi pc; pc.setup(a);
        PcgSolver solver(1e-12, 50);
        SolveStats st = solver.with_fused_reduction(true).solve(a, &pc, b, x);
        REQUIRE(st.converged);
        for (int i = 0; i < 4; ++i) REQUIRE(std::fabs(x[i] - xt[i]) < 1e-9);
    }
    {   // bicgstab_solves_well_conditioned_nonsym (src/solver/bicgstab.rs:303-328)
        Csr m; m.n = 3; m.rp = {0, 3, 6, 9}; m.ci = {0, 1, 2, 0, 1, 2, 0, 1, 2};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.v.push_back(i == j ? 4.0 : double(i + 2 * j + 1));
        DeviceCsr a = DeviceCsr::from_csr(ctx, 3, 3, m.rp, m.ci, m.v);
        std::vector<double> xt = {1, 2, 3}, b(3), x(3, 0.0);
        a.matvec(xt, b);
        BiCgStabSolver solver(1e-10, 100);
        SolveStats st = solver.solve(a, nullptr, b, x);
        REQUIRE(st.converged);
        for (int i = 0; i < 3; ++i) REQUIRE(std::fabs(x[i] - xt[i]) < 1e-8);
    }
    {   // error behaviour: IndefiniteMatrix leaves x untouched (pcg.rs:162-172); dot / norm (tests/core_dense.rs:38-47)
        Csr m; m.n = 3; m.rp = {0, 1, 2, 3}; m.ci = {0, 1, 2}; m.v = {1.0, -1.0, 2.0};
        DeviceCsr a = DeviceCsr::from_csr(ctx, 3, 3, m.rp, m.ci, m.v);
        std::vector<double> b = {1, 1, 1}, x(3, 0.0);
        bool threw = false;
        try { PcgSolver(1e-10, 50).solve(a, nullptr, b, x); } catch (const KError& e) { threw = (e.kind == KError::IndefiniteMatrix); }
        REQUIRE(threw); REQUIRE(x[0] == 0.0 && x[1] == 0.0 && x[2] == 0.0);
        REQUIRE(std::fabs(ctx.dot({1, 2, 3}, {4, -5, 6}) - 12.0) < 1e-12);
        REQUIRE(std::fabs(ctx.norm({1, 2, 3}) - std::sqrt(14.0)) < 1e-12);
    }
    {   // asm_dense_lu_blocks (src/preconditioner/asm.rs:124-136): identity, two blocks -> apply(r) == r; PC factory; KspContext
        Csr m; m.n = 4; m.rp = {0, 1, 2, 3, 4}; m.ci = {0, 1, 2, 3}; m.v = {1, 1, 1, 1};
        DeviceCsr a = DeviceCsr::from_csr(ctx, 4, 4, m.rp, m.ci, m.v);
        AdditiveSchwarz asm_pc(0, std::vector<std::vector<uint64_t>>{{0, 1}, {2, 3}});
        asm_pc.setup(a);
        std::vector<double> r = {1, 2, 3, 4}, z(4, 0.0);
        asm_pc.apply(r, z);
        REQUIRE(z == r);
        Csr t = tridiag(12, -1, 2.5, -0.5);
        DeviceCsr at = DeviceCsr::from_csr(ctx, t.n, t.n, t.rp, t.ci, t.v);
        std::vector<double> ones(t.n, 1.0), b(t.n), x(t.n, 0.0), x2(t.n, 0.0);
        at.matvec(ones, b);
        PC::Built pc = PC::AdditiveSchwarz(0, 3).build(at);                 // pc_context.rs:75 -> device preconditioner
        KspContext ksp{SolverKind::GmresLeft, &at, &pc, 1e-12, 200, 12};    // ksp_context.rs:54-69
        SolveStats st = ksp.solve_context(b, x);
        REQUIRE(st.converged); REQUIRE(rel_error(x, 1.0) < 1e-10);
        SolveStats st2 = GmresSolver(12, 1e-12, 200).with_preconditioning(Preconditioning::Left).solve(at, &pc, b, x2);
        REQUIRE(st2.iterations == st.iterations); REQUIRE(x == x2);
        bool unsupported = false;
        try { PC::AMG().build(at); } catch (const KError& e) { unsupported = (e.kind == KError::Unsupported); }
        REQUIRE(unsupported);
        unsupported = false;
        try { KspContext k2{SolverKind::Minres, &at, nullptr, 1e-8, 10, 5}; k2.solve_context(b, x2); } catch (const KError& e) { unsupported = (e.kind == KError::Unsupported); }
        REQUIRE(unsupported);
    }
    {   // Comm surface on one rank (src/parallel/mod.rs:4-35; rayon_comm.rs:56-78)
        std::vector<double> g = {1, 2, 3}, out(3, 0.0), gathered;
        ctx.scatter(g, out, 0);
        REQUIRE(out == g);
        ctx.gather(out, gathered, 0);
        REQUIRE(gathered == g);
        REQUIRE(std::fabs(ctx.comm_dot({1, 2, 3}, {4, -5, 6}) - 12.0) < 1e-12);
        REQUIRE(std::fabs(ctx.comm_norm({1, 2, 3}) - std::sqrt(14.0)) < 1e-12);
        REQUIRE(ctx.rank() == 0 && ctx.size() == 1);
    }
    std::puts("CPP_HOST_API_OK");
    return 0;
}
