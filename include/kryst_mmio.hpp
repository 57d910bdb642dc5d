// kryst_mmio.hpp — Matrix Market coordinate I/O for the arrays CsrMatrix::from_csr takes.
//
// The reference has no on-disk format (its tests and examples build matrices in code, src/matrix/sparse.rs:26-47 is
// the only ingestion point); this header-only reader/writer is the host-side step before the path so that user
// matrices can reach kryst::DeviceCsr::from_csr.  Same rules as kryst_b200/mmio.py (the two are tested against each
// other): `matrix coordinate {real|integer|pattern} {general|symmetric|skew-symmetric}`, entries sorted by
// (row, column), repeated entries summed in file order, explicit zeros kept.  Pure host code: no CUDA, no kryst_b200.h.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace kryst {

struct HostCsr {                       // the argument list of CsrMatrix::from_csr
    size_t nrows = 0, ncols = 0;
    std::vector<uint64_t> row_ptr, col_idx;
    std::vector<double> values;
};

struct MatrixMarketError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace detail {
inline std::string lower(std::string s) {
    for (char& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    return s;
}
// stable sort by (row, column), sum repeats in input order
inline HostCsr coo_to_csr(size_t nrows, size_t ncols, const std::vector<uint64_t>& ri, const std::vector<uint64_t>& ci,
                          const std::vector<double>& v) {
    std::vector<size_t> order(ri.size());
    std::iota(order.begin(), order.end(), size_t{0});
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ri[a] != ri[b] ? ri[a] < ri[b] : ci[a] < ci[b]; });
    HostCsr out;
    out.nrows = nrows; out.ncols = ncols;
    out.row_ptr.assign(nrows + 1, 0);
    for (size_t k = 0; k < order.size(); ++k) {
        const size_t e = order[k];
        if (k > 0 && ri[e] == ri[order[k - 1]] && ci[e] == ci[order[k - 1]]) { out.values.back() = out.values.back() + v[e]; continue; }
        out.col_idx.push_back(ci[e]);
        out.values.push_back(v[e]);
        out.row_ptr[ri[e] + 1] += 1;
    }
    for (size_t i = 0; i < nrows; ++i) out.row_ptr[i + 1] += out.row_ptr[i];
    return out;
}
}  // namespace detail

inline HostCsr read_matrix_market(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw MatrixMarketError("cannot open " + path);
    std::string line;
    if (!std::getline(f, line)) throw MatrixMarketError("empty file");
    std::istringstream hs(line);
    std::string banner, object, format, field, symmetry;
    hs >> banner >> object >> format >> field >> symmetry;
    if (banner != "%%MatrixMarket" || detail::lower(object) != "matrix") throw MatrixMarketError("not a Matrix Market matrix file");
    format = detail::lower(format); field = detail::lower(field); symmetry = detail::lower(symmetry);
    if (format != "coordinate") throw MatrixMarketError("only the coordinate format is supported, got '" + format + "'");
    if (field != "real" && field != "integer" && field != "pattern") throw MatrixMarketError("unsupported field '" + field + "'");
    if (symmetry != "general" && symmetry != "symmetric" && symmetry != "skew-symmetric") throw MatrixMarketError("unsupported symmetry '" + symmetry + "'");
    do {
        if (!std::getline(f, line)) throw MatrixMarketError("missing size line");
    } while (line.empty() || line[0] == '%' || line.find_first_not_of(" \t\r") == std::string::npos);
    unsigned long long nrows = 0, ncols = 0, nent = 0;
    {
        std::istringstream ss(line);
        if (!(ss >> nrows >> ncols >> nent)) throw MatrixMarketError("bad size line '" + line + "'");
    }
    const bool pattern = field == "pattern", general = symmetry == "general", skew = symmetry == "skew-symmetric";
    if (!general && nrows != ncols) throw MatrixMarketError("symmetric storage needs a square matrix");
    std::vector<uint64_t> ri, ci;
    std::vector<double> v;
    ri.reserve(nent); ci.reserve(nent); v.reserve(nent);
    std::vector<uint64_t> mi, mj;          // mirrored off-diagonal entries, appended after the stored ones (as mmio.py does)
    std::vector<double> mv;
    for (unsigned long long k = 0; k < nent; ++k) {
        unsigned long long i = 0, j = 0;
        double x = 1.0;
        if (!(f >> i >> j)) throw MatrixMarketError("expected " + std::to_string(nent) + " entries, file ends after " + std::to_string(k));
        if (!pattern && !(f >> x)) throw MatrixMarketError("entry " + std::to_string(k + 1) + " has no value");
        if (i < 1 || j < 1 || i > nrows || j > ncols) throw MatrixMarketError("entry index out of range");
        ri.push_back(i - 1); ci.push_back(j - 1); v.push_back(x);
        if (!general) {
            if (i != j) { mi.push_back(j - 1); mj.push_back(i - 1); mv.push_back(skew ? -x : x); }
            else if (skew) throw MatrixMarketError("skew-symmetric storage cannot hold diagonal entries");
        }
    }
    ri.insert(ri.end(), mi.begin(), mi.end()); ci.insert(ci.end(), mj.begin(), mj.end()); v.insert(v.end(), mv.begin(), mv.end());
    return detail::coo_to_csr(nrows, ncols, ri, ci, v);
}

// `matrix coordinate real general`, 17 significant digits (round-trips f64 exactly)
inline void write_matrix_market(const std::string& path, const HostCsr& a, const std::string& comment = "") {
    std::FILE* f = std::fopen(path.c_str(), "w");
    if (!f) throw MatrixMarketError("cannot open " + path + " for writing");
    std::fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n");
    if (!comment.empty()) std::fprintf(f, "%% %s\n", comment.c_str());
    std::fprintf(f, "%zu %zu %zu\n", a.nrows, a.ncols, a.col_idx.size());
    for (size_t i = 0; i < a.nrows; ++i)
        for (uint64_t p = a.row_ptr[i]; p < a.row_ptr[i + 1]; ++p)
            std::fprintf(f, "%zu %llu %.17g\n", i + 1, static_cast<unsigned long long>(a.col_idx[p] + 1), a.values[p]);
    std::fclose(f);
}

}  // namespace kryst
