// Host-only check of include/kryst_mmio.hpp (no CUDA): `test_mmio in.mtx [out.mtx]` prints the CSR arrays in a plain
// text form that tests/test_ingest.py compares with kryst_b200/mmio.py, and optionally rewrites the matrix.
// Exit code 2 + message on a MatrixMarketError.
#include <cstdio>
#include "kryst_mmio.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: test_mmio in.mtx [out.mtx]\n"); return 1; }
    try {
        kryst::HostCsr a = kryst::read_matrix_market(argv[1]);
        std::printf("%zu %zu %zu\n", a.nrows, a.ncols, a.col_idx.size());
        for (uint64_t p : a.row_ptr) std::printf("%llu ", static_cast<unsigned long long>(p));
        std::printf("\n");
        for (uint64_t c : a.col_idx) std::printf("%llu ", static_cast<unsigned long long>(c));
        std::printf("\n");
        for (double v : a.values) std::printf("%a ", v);       // hex floats: exact
        std::printf("\n");
        if (argc > 2) kryst::write_matrix_market(argv[2], a, "rewritten by test_mmio");
    } catch (const kryst::MatrixMarketError& e) {
        std::fprintf(stderr, "MatrixMarketError: %s\n", e.what());
        return 2;
    }
    return 0;
}
