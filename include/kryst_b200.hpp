// kryst_b200.hpp — C++ host-side mirror of kryst's trait API on top of the C ABI (kryst_b200.h).
//
// The reference is Rust; no Rust toolchain exists in the build image, so the host layer above the C ABI is
// written in C++ (and in Python, kryst_b200/api.py) with the reference's names, argument meaning and error
// behaviour, so that code and tests read like the reference's own:
//
//     kryst::DeviceCsr a = kryst::DeviceCsr::from_csr(ctx, nrows, ncols, row_ptr, col_idx, values);
//     kryst::Jacobi pc;  pc.setup(a);                              // Preconditioner::setup
//     kryst::PcgSolver solver(1e-8, 1000);                         // PcgSolver::new(tol, max_iters)
//     kryst::SolveStats st = solver.solve(a, &pc, b, x);           // LinearSolver::solve(&a, Some(&pc), &b, &mut x)
//
// Reference interfaces mirrored (paths relative to the kryst crate root):
//   MatVec / Indexing / MatShape   src/core/traits.rs:4-35         Preconditioner   src/preconditioner/mod.rs:8-13
//   LinearSolver                   src/solver/mod.rs:30-52         SolveStats       src/utils/convergence.rs:9-14
//   KError                         src/error.rs:6-19               Comm             src/parallel/mod.rs:4-35
// Errors: Rust `Result<_, KError>` becomes a thrown kryst::KError carrying the same discriminant.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "kryst_b200.h"

namespace kryst {

struct KError : std::runtime_error {
    enum Kind { FactorError = 1, SolveError = 2, IndefiniteMatrix = 3, IndefinitePreconditioner = 4, ZeroPivot = 5, Unsupported = 6 };
    Kind kind;
    uint64_t row;   // ZeroPivot(row)
    KError(int k, const std::string& msg, uint64_t r = 0) : std::runtime_error(msg), kind(static_cast<Kind>(k)), row(r) {}
};
inline void check(int status, uint64_t row = 0) {
    if (status != KB_OK) throw KError(status, kb_last_error(), row);
}

struct SolveStats {          // src/utils/convergence.rs:9-14
    size_t iterations;
    double final_residual;
    bool converged;
};
inline SolveStats to_stats(const kb_stats& s) { return SolveStats{static_cast<size_t>(s.iterations), s.final_residual, s.converged != 0}; }

// One GPU + stream (+ communicator).  Comm surface: rank/size/barrier/all_reduce (parallel/mod.rs:4-35).
class Context {
public:
    explicit Context(int device = 0) { check(kb_ctx_create(device, &h_)); }
    ~Context() { kb_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    kb_ctx handle() const { return h_; }
    int rank() const { return kb_comm_rank(h_); }
    int size() const { return kb_comm_size(h_); }
    void barrier() const { check(kb_comm_barrier(h_)); }
    double all_reduce(double x) const { double g = 0; check(kb_comm_all_reduce(h_, x, &g)); return g; }
    void comm_init(int rank, int size, const void* id128) { check(kb_comm_init(h_, rank, size, id128)); }
    // InnerProduct<Vec<f64>> for () (src/core/wrappers.rs:87-129)
    double dot(const std::vector<double>& x, const std::vector<double>& y) const { double d = 0; check(kb_dot(h_, x.size(), x.data(), y.data(), &d)); return d; }
    double norm(const std::vector<double>& x) const { double d = 0; check(kb_norm(h_, x.size(), x.data(), &d)); return d; }
private:
    kb_ctx h_ = nullptr;
};

// Device-resident CSR operator: MatVec<Vec<f64>> + Indexing + MatShape (replaces CsrMatrix, sparse.rs:22-68)
class DeviceCsr {
public:
    static DeviceCsr from_csr(const Context& ctx, size_t nrows, size_t ncols, const std::vector<uint64_t>& row_ptr,
                              const std::vector<uint64_t>& col_idx, const std::vector<double>& values) {
        DeviceCsr a;
        check(kb_csr_create(ctx.handle(), nrows, ncols, row_ptr.data(), col_idx.data(), values.data(), &a.h_));
        return a;
    }
    static DeviceCsr from_csr_shard(const Context& ctx, size_t n_global, size_t row_lo, size_t row_hi, const std::vector<uint64_t>& row_ptr,
                                    const std::vector<uint64_t>& col_idx_global, const std::vector<double>& values) {
        DeviceCsr a;
        check(kb_csr_create_dist(ctx.handle(), n_global, row_lo, row_hi, row_ptr.data(), col_idx_global.data(), values.data(), &a.h_));
        return a;
    }
    DeviceCsr(DeviceCsr&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    DeviceCsr& operator=(DeviceCsr&& o) noexcept { if (this != &o) { kb_csr_destroy(h_); h_ = o.h_; o.h_ = nullptr; } return *this; }
    ~DeviceCsr() { kb_csr_destroy(h_); }
    void matvec(const std::vector<double>& x, std::vector<double>& y) const { check(kb_csr_matvec(h_, x.data(), y.data())); }   // MatVec::matvec
    // SubmatrixExtract::submatrix (sparse.rs:72-93; used by AdditiveSchwarz::setup, asm.rs:58-65)
    DeviceCsr submatrix(const std::vector<uint64_t>& indices) const {
        DeviceCsr s;
        check(kb_csr_submatrix(h_, indices.data(), indices.size(), &s.h_));
        return s;
    }
    // the arrays CsrMatrix::from_csr takes, read back from the device
    void to_csr(std::vector<uint64_t>& row_ptr, std::vector<uint64_t>& col_idx, std::vector<double>& values) const {
        row_ptr.assign(nrows() + 1, 0); col_idx.assign(kb_csr_nnz(h_), 0); values.assign(kb_csr_nnz(h_), 0.0);
        check(kb_csr_download(h_, row_ptr.data(), col_idx.data(), values.data()));
    }
    size_t nrows() const { return kb_csr_nrows(h_); }     // Indexing::nrows / MatShape::nrows
    size_t ncols() const { return kb_csr_ncols(h_); }     // MatShape::ncols
    kb_csr handle() const { return h_; }
private:
    DeviceCsr() = default;
    kb_csr h_ = nullptr;
};

// Preconditioner<DeviceCsr, Vec<f64>> (src/preconditioner/mod.rs:8-13)
class Preconditioner {
public:
    virtual ~Preconditioner() { kb_pc_destroy(h_); }
    virtual void setup(const DeviceCsr& a) = 0;
    void apply(const std::vector<double>& r, std::vector<double>& z) const {
        if (!h_) throw KError(KB_SOLVE_ERROR, "preconditioner used before setup()");
        check(kb_pc_apply(h_, r.data(), z.data()));
    }
    kb_pc handle() const { return h_; }
protected:
    void reset(kb_pc h) { kb_pc_destroy(h_); h_ = h; }
    kb_pc h_ = nullptr;
};
class Jacobi : public Preconditioner {      // src/preconditioner/jacobi.rs:26-95
public:
    void setup(const DeviceCsr& a) override { kb_pc h = nullptr; check(kb_pc_create_jacobi(a.handle(), &h)); reset(h); }
};
class Ilu0 : public Preconditioner {        // replaces src/preconditioner/ilu.rs:32-122; block-Jacobi ILU(0) on a shard
public:
    void setup(const DeviceCsr& a) override {
        kb_pc h = nullptr;
        int st = kb_pc_create_ilu0(a.handle(), &h);
        if (st != KB_OK) { uint64_t row = h ? kb_pc_bad_row(h) : 0; std::string msg = kb_last_error(); kb_pc_destroy(h); throw KError(st, msg, row); }
        reset(h);
    }
};

enum class CgNormType { Preconditioned = 0, Unpreconditioned = 1, Natural = 2, None = 3 };   // pcg.rs:25
enum class Preconditioning { None = 0, Left = 1, Right = 2 };                                // gmres.rs:28-32

// LinearSolver<DeviceCsr, Vec<f64>> (src/solver/mod.rs:30-52): x is in/out, written only on Ok
class PcgSolver {            // src/solver/pcg.rs:31-222
public:
    PcgSolver(double tol, size_t max_iters) : tol_(tol), max_iters_(max_iters) {}
    PcgSolver& with_norm(CgNormType t) { norm_ = t; return *this; }
    // pcg.rs:66-80: the reference stores these three and its solve never changes arithmetic on them
    // (single_reduction only swaps the Rayon dot for a serial loop, pcg.rs:151-160); same here.
    PcgSolver& with_single_reduction(bool f) { single_reduction_ = f; return *this; }
    // extension: true single-reduction (Chronopoulos-Gear) recurrences, KB_FLAG_SINGLE_REDUCTION
    PcgSolver& with_fused_reduction(bool f = true) { fused_ = f; return *this; }
    PcgSolver& with_radius(double r) { radius_ = r; has_radius_ = true; return *this; }
    PcgSolver& with_obj_target(double t) { obj_target_ = t; has_obj_target_ = true; return *this; }
    std::vector<double> residual_history;
    SolveStats solve(const DeviceCsr& a, const Preconditioner* pc, const std::vector<double>& b, std::vector<double>& x) {
        std::vector<double> hist(max_iters_ + 1 < (1u << 20) ? max_iters_ + 1 : (1u << 20));
        uint64_t hl = 0;
        kb_stats st{};
        int rc = kb_pcg_solve(a.handle(), pc ? pc->handle() : nullptr, b.data(), x.data(), tol_, max_iters_, static_cast<int>(norm_), fused_ ? KB_FLAG_SINGLE_REDUCTION : 0u,
                              hist.data(), hist.size(), &hl, &st);
        residual_history.insert(residual_history.end(), hist.begin(), hist.begin() + (hl < hist.size() ? hl : hist.size()));
        check(rc);
        return to_stats(st);
    }
private:
    double tol_; size_t max_iters_; CgNormType norm_ = CgNormType::Unpreconditioned;
    bool single_reduction_ = false, fused_ = false, has_radius_ = false, has_obj_target_ = false;
    double radius_ = 0.0, obj_target_ = 0.0;
};
class GmresSolver {          // src/solver/gmres.rs:38-402
public:
    GmresSolver(size_t restart, double tol, size_t max_iters) : restart_(restart), tol_(tol), max_iters_(max_iters) {}
    GmresSolver& with_preconditioning(Preconditioning m) { mode_ = m; return *this; }
    SolveStats solve(const DeviceCsr& a, const Preconditioner* pc, const std::vector<double>& b, std::vector<double>& x) {
        kb_stats st{};
        check(kb_gmres_solve(a.handle(), pc ? pc->handle() : nullptr, b.data(), x.data(), restart_, tol_, max_iters_, static_cast<int>(mode_), 0, &st));
        return to_stats(st);
    }
private:
    size_t restart_; double tol_; size_t max_iters_; Preconditioning mode_ = Preconditioning::Left;   // gmres.rs:53
};
class FgmresSolver {         // src/solver/fgmres.rs:30-340: FgmresSolver::new(tol, max_iters, restart).solve_flex(..)
public:
    FgmresSolver(double tol, size_t max_iters, size_t restart) : tol_(tol), max_iters_(max_iters), restart_(restart) {}
    SolveStats solve_flex(const DeviceCsr& a, const Preconditioner* pc, const std::vector<double>& b, std::vector<double>& x) {
        kb_stats st{};
        check(kb_fgmres_solve(a.handle(), pc ? pc->handle() : nullptr, b.data(), x.data(), restart_, tol_, max_iters_, 0, &st));
        return to_stats(st);
    }
private:
    double tol_; size_t max_iters_; size_t restart_;
};
class BiCgStabSolver {       // src/solver/bicgstab.rs:36-293 (textbook=true: right-preconditioned, relative tol)
public:
    BiCgStabSolver(double tol, size_t max_iters, bool textbook = false) : tol_(tol), max_iters_(max_iters), textbook_(textbook) {}
    SolveStats solve(const DeviceCsr& a, const Preconditioner* pc, const std::vector<double>& b, std::vector<double>& x) {
        kb_stats st{};
        check(kb_bicgstab_solve(a.handle(), pc ? pc->handle() : nullptr, b.data(), x.data(), tol_, max_iters_, textbook_ ? KB_FLAG_TEXTBOOK : 0, &st));
        return to_stats(st);
    }
private:
    double tol_; size_t max_iters_; bool textbook_;
};

}  // namespace kryst
