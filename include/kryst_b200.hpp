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
    // Comm::dot (parallel/mod.rs:19-22), DistributedInnerProduct::{dot,norm} (core/wrappers.rs:134-156): every rank passes its slice
    double comm_dot(const std::vector<double>& a, const std::vector<double>& b) const { double d = 0; check(kb_comm_dot(h_, a.size(), a.data(), b.data(), &d)); return d; }
    double comm_norm(const std::vector<double>& x) const { double d = 0; check(kb_comm_norm(h_, x.size(), x.data(), &d)); return d; }
    // Comm::scatter / Comm::gather (parallel/mod.rs:9-16): equal chunks; `global` is read and `out` written on root only
    template <class T> void scatter(const std::vector<T>& global, std::vector<T>& out, int root) const {
        check(kb_comm_scatter(h_, global.data(), out.size() * sizeof(T), out.data(), root));
    }
    template <class T> void gather(const std::vector<T>& local, std::vector<T>& out, int root) const {
        if (rank() == root) out.assign(local.size() * static_cast<size_t>(size()), T()); else out.clear();
        check(kb_comm_gather(h_, local.data(), local.size() * sizeof(T), out.data(), root));
    }
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

// AdditiveSchwarz::new(overlap, subdomains) (src/preconditioner/asm.rs:17-116) on one GPU: sub-operators from SubmatrixExtract,
// one ILU(0) (or Jacobi) application per block as the inner solve, block results summed in subdomain order.
class AdditiveSchwarz : public Preconditioner {
public:
    enum class Inner { Ilu0 = KB_ASM_INNER_ILU0, Jacobi = KB_ASM_INNER_JACOBI };
    AdditiveSchwarz(size_t overlap, std::vector<std::vector<uint64_t>> subdomains, Inner inner = Inner::Ilu0)
        : overlap_(overlap), nsub_(subdomains.size()), sub_(std::move(subdomains)), inner_(inner) {}
    // `Vec::with_capacity(p)` idiom of the reference: p uniform row chunks (asm.rs:46-57)
    AdditiveSchwarz(size_t overlap, size_t p, Inner inner = Inner::Ilu0) : overlap_(overlap), nsub_(p), inner_(inner) {}
    void setup(const DeviceCsr& a) override {
        std::vector<uint64_t> ptr, idx;
        if (!sub_.empty()) {
            ptr.push_back(0);
            for (const auto& s : sub_) { idx.insert(idx.end(), s.begin(), s.end()); ptr.push_back(idx.size()); }
        }
        kb_pc h = nullptr;
        int st = kb_pc_create_asm(a.handle(), overlap_, nsub_, sub_.empty() ? nullptr : ptr.data(), sub_.empty() ? nullptr : idx.data(), static_cast<int>(inner_), &h);
        if (st != KB_OK) { uint64_t row = h ? kb_pc_bad_row(h) : 0; std::string msg = kb_last_error(); kb_pc_destroy(h); throw KError(st, msg, row); }
        reset(h);
    }
private:
    size_t overlap_, nsub_;
    std::vector<std::vector<uint64_t>> sub_;
    Inner inner_;
};

// PC<T> (src/context/pc_context.rs:36-76) + the factory the reference's enum lacks: PC::Ilu0().build(a)
struct PC {
    int kind = KB_PCK_JACOBI; uint64_t fill = 0; double droptol = 0.0; uint64_t overlap = 0, nblocks = 0;
    std::vector<std::vector<uint64_t>> blocks;
    static PC Jacobi() { PC p; p.kind = KB_PCK_JACOBI; return p; }
    static PC Ssor() { PC p; p.kind = KB_PCK_SSOR; return p; }
    static PC Ilu0() { PC p; p.kind = KB_PCK_ILU0; return p; }
    static PC Ilup(uint64_t fill) { PC p; p.kind = KB_PCK_ILUP; p.fill = fill; return p; }
    static PC Ilut(uint64_t fill, double droptol) { PC p; p.kind = KB_PCK_ILUT; p.fill = fill; p.droptol = droptol; return p; }
    static PC BlockJacobi(std::vector<std::vector<uint64_t>> b) { PC p; p.kind = KB_PCK_BLOCK_JACOBI; p.blocks = std::move(b); return p; }
    static PC AMG() { PC p; p.kind = KB_PCK_AMG; return p; }
    static PC AdditiveSchwarz(uint64_t overlap = 0, uint64_t nblocks = 1) { PC p; p.kind = KB_PCK_ADDITIVE_SCHWARZ; p.overlap = overlap; p.nblocks = nblocks; return p; }
    // -> owning preconditioner handle; variants outside the device path throw KError::Unsupported
    class Built : public Preconditioner { public: explicit Built(kb_pc h) { reset(h); } void setup(const DeviceCsr&) override {} };
    Built build(const DeviceCsr& a) const {
        std::vector<uint64_t> ptr, idx;
        kb_pc_spec s{};
        s.kind = kind; s.fill = fill; s.droptol = droptol; s.overlap = overlap; s.nblocks = nblocks;
        if (!blocks.empty()) {
            ptr.push_back(0);
            for (const auto& b : blocks) { idx.insert(idx.end(), b.begin(), b.end()); ptr.push_back(idx.size()); }
            s.nblocks = blocks.size(); s.block_ptr = ptr.data(); s.block_idx = idx.data();
        }
        kb_pc h = nullptr;
        int st = kb_pc_create_from_spec(a.handle(), &s, &h);
        if (st != KB_OK) { uint64_t row = h ? kb_pc_bad_row(h) : 0; std::string msg = kb_last_error(); kb_pc_destroy(h); throw KError(st, msg, row); }
        return Built(h);
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
    // extension: pipelined (Ghysels-Vanroose) recurrences, KB_FLAG_PIPELINED
    PcgSolver& with_pipelined(bool f = true) { pipelined_ = f; return *this; }
    PcgSolver& with_radius(double r) { radius_ = r; has_radius_ = true; return *this; }
    PcgSolver& with_obj_target(double t) { obj_target_ = t; has_obj_target_ = true; return *this; }
    std::vector<double> residual_history;
    SolveStats solve(const DeviceCsr& a, const Preconditioner* pc, const std::vector<double>& b, std::vector<double>& x) {
        std::vector<double> hist(max_iters_ + 1 < (1u << 20) ? max_iters_ + 1 : (1u << 20));
        uint64_t hl = 0;
        kb_stats st{};
        int rc = kb_pcg_solve(a.handle(), pc ? pc->handle() : nullptr, b.data(), x.data(), tol_, max_iters_, static_cast<int>(norm_), (fused_ ? KB_FLAG_SINGLE_REDUCTION : 0u) | (pipelined_ ? KB_FLAG_PIPELINED : 0u),
                              hist.data(), hist.size(), &hl, &st);
        residual_history.insert(residual_history.end(), hist.begin(), hist.begin() + (hl < hist.size() ? hl : hist.size()));
        check(rc);
        return to_stats(st);
    }
private:
    double tol_; size_t max_iters_; CgNormType norm_ = CgNormType::Unpreconditioned;
    bool single_reduction_ = false, fused_ = false, pipelined_ = false, has_radius_ = false, has_obj_target_ = false;
    double radius_ = 0.0, obj_target_ = 0.0;
};
class GmresSolver {          // src/solver/gmres.rs:38-402
public:
    GmresSolver(size_t restart, double tol, size_t max_iters) : restart_(restart), tol_(tol), max_iters_(max_iters) {}
    GmresSolver& with_preconditioning(Preconditioning m) { mode_ = m; return *this; }
    // extension (block orthogonalisation, the idea of pca_gmres.rs:172-229): one CGS pass, one fused reduction per step
    GmresSolver& with_block_orthogonalisation(bool on = true) { block_ = on; return *this; }
    SolveStats solve(const DeviceCsr& a, const Preconditioner* pc, const std::vector<double>& b, std::vector<double>& x) {
        kb_stats st{};
        check(kb_gmres_solve(a.handle(), pc ? pc->handle() : nullptr, b.data(), x.data(), restart_, tol_, max_iters_, static_cast<int>(mode_),
                             block_ ? KB_FLAG_BLOCK_ORTH : 0u, &st));
        return to_stats(st);
    }
private:
    size_t restart_; double tol_; size_t max_iters_; Preconditioning mode_ = Preconditioning::Left;   // gmres.rs:53
    bool block_ = false;
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

// SolverKind + KspContext::solve_context (src/context/ksp_context.rs:25-69,88-148): pure dispatch, behind kb_ksp_solve
enum class SolverKind { Cg = KB_KSP_CG, Pcg = KB_KSP_PCG, GmresLeft = KB_KSP_GMRES_LEFT, GmresRight = KB_KSP_GMRES_RIGHT, Fgmres = KB_KSP_FGMRES,
                        Bicgstab = KB_KSP_BICGSTAB, Cgs = KB_KSP_CGS, Qmr = KB_KSP_QMR, Tfqmr = KB_KSP_TFQMR, Minres = KB_KSP_MINRES, Cgnr = KB_KSP_CGNR };
struct KspContext {          // public fields like the reference's
    SolverKind kind;
    const DeviceCsr* a;
    const Preconditioner* pc = nullptr;        // Fgmres: the flexible preconditioner (flex_pc)
    double tol = 1e-8;
    size_t max_it = 1000;
    size_t restart = 30;
    SolveStats solve_context(const std::vector<double>& b, std::vector<double>& x) const {
        kb_ksp k{static_cast<int32_t>(kind), tol, max_it, restart};
        kb_stats st{};
        check(kb_ksp_solve(a->handle(), pc ? pc->handle() : nullptr, &k, b.data(), x.data(), 0, &st));
        return to_stats(st);
    }
};

}  // namespace kryst
