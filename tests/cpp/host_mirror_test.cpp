// Drives the C++ host mirror (halo2_scaffold_b200/host/h2b200.hpp) the way the reference's callers drive
// halo2_proofs: best_multiexp / best_fft / EvaluationDomain / ParamsKZG.  Inputs and outputs are raw little-endian
// files in <dir> so that tests/test_gpu_parity.py can compare against the oracle and the golden fixtures.
//   usage: host_mirror_test <dir> <j> <k>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../halo2_scaffold_b200/host/h2b200.hpp"

using namespace h2b200;

template <class T>
static std::vector<T> read_all(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<T> v((size_t)sz / sizeof(T));
    if (sz && fread(v.data(), 1, (size_t)sz, f) != (size_t)sz) exit(2);
    fclose(f);
    return v;
}
template <class T>
static void write_all(const std::string& path, const T* p, size_t n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f || fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string dir = argv[1];
    const uint32_t j = (uint32_t)atoi(argv[2]), k = (uint32_t)atoi(argv[3]);
    try {
        // best_multiexp
        auto scalars = read_all<Fr>(dir + "/scalars.bin");
        auto bases = read_all<G1Affine>(dir + "/bases.bin");
        G1 r = arithmetic::best_multiexp(scalars, bases);
        write_all(dir + "/msm.bin", &r, 1);
        // the reference asserts equal lengths
        bool panicked = false;
        try {
            std::vector<Fr> shorter(scalars.begin(), scalars.end() - 1);
            arithmetic::best_multiexp(shorter, bases);
        } catch (const Panic&) { panicked = true; }
        if (!panicked) { fprintf(stderr, "length mismatch did not panic\n"); return 1; }
        // best_fft through the domain wrappers
        poly::EvaluationDomain dom(j, k);
        auto lagrange = read_all<Fr>(dir + "/lagrange.bin");
        auto coeff = dom.lagrange_to_coeff(lagrange);
        write_all(dir + "/coeff.bin", coeff.data(), coeff.size());
        auto ext = dom.coeff_to_extended(coeff);
        write_all(dir + "/extended.bin", ext.data(), ext.size());
        auto back = dom.extended_to_coeff(ext);
        write_all(dir + "/back.bin", back.data(), back.size());
        // plain best_fft round trip: forward with omega, then the inverse root
        std::vector<Fr> a = lagrange;
        arithmetic::best_fft(a, dom.get_omega(), k);
        write_all(dir + "/fft.bin", a.data(), a.size());
        panicked = false;
        try { arithmetic::best_fft(a, dom.get_omega(), k + 1); } catch (const Panic&) { panicked = true; }
        if (!panicked) { fprintf(stderr, "best_fft length mismatch did not panic\n"); return 1; }
        // ParamsKZG::commit / commit_lagrange over registered SRS vectors (n = bases.size() must be 2^k here)
        if (bases.size() == ((size_t)1 << k)) {
            poly::kzg::ParamsKZG params(k, bases, bases);
            G1 c[2] = {params.commit(scalars), params.commit_lagrange(std::vector<Fr>(scalars.begin(), scalars.begin() + scalars.size() / 2))};
            write_all(dir + "/commit.bin", c, 2);
        }
    } catch (const Panic& e) {
        fprintf(stderr, "panic: %s\n", e.what());
        return 1;
    }
    printf("HOST_MIRROR_OK\n");
    return 0;
}
