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
            // the same two commitments + a third column through ONE batched call, and a column through the one-upload pipeline
            std::vector<std::vector<Fr>> cols = {scalars, std::vector<Fr>(scalars.begin(), scalars.begin() + scalars.size() / 2), lagrange};
            std::vector<G1> many = params.commit_lagrange_many(cols);
            write_all(dir + "/commit_many.bin", many.data(), many.size());
            auto cc = params.commit_lagrange_and_convert(dom, lagrange);
            write_all(dir + "/pipeline_commit.bin", &cc.commitment, 1);
            write_all(dir + "/pipeline_coeff.bin", cc.coeff.data(), cc.coeff.size());
            write_all(dir + "/pipeline_extended.bin", cc.extended.data(), cc.extended.size());
        }
        // ---- the widened rows: vanishing division, SRS file round trip, GraphEvaluator, grand products, lookup permutation
        auto divided = dom.divide_by_vanishing_poly(ext);
        write_all(dir + "/divided.bin", divided.data(), divided.size());
        if (bases.size() == ((size_t)1 << k)) {
            std::vector<uint8_t> g2(128);
            for (size_t i = 0; i < g2.size(); ++i) g2[i] = (uint8_t)(i * 7 + 1);
            std::vector<G1Affine> reversed(bases.rbegin(), bases.rend());
            poly::kzg::ParamsKZG params(k, bases, reversed, g2);
            params.write_custom(dir + "/params.srs", poly::kzg::SerdeFormat::Processed);
            poly::kzg::ParamsKZG back2(dir + "/params.srs", poly::kzg::SerdeFormat::Processed);
            if (back2.k() != k || std::memcmp(back2.get_g().data(), bases.data(), bases.size() * sizeof(G1Affine)) != 0 ||
                std::memcmp(back2.get_g_lagrange().data(), reversed.data(), reversed.size() * sizeof(G1Affine)) != 0 || back2.g2_bytes() != g2) {
                fprintf(stderr, "SRS round trip differs\n");
                return 1;
            }
            G1 c2 = back2.commit(scalars);
            write_all(dir + "/commit_after_read.bin", &c2, 1);
        }
        {
            using namespace plonk::evaluation;
            // q * (a + b[next] * c - a[prev]) folded into the previous value with y, built the way Evaluator::new builds it
            GraphEvaluator ge;
            const uint32_t r0 = ge.add_rotation(0), r1 = ge.add_rotation(1), rm = ge.add_rotation(-1);
            ValueSource q = ge.add_calculation(H2B_CALC_STORE, ValueSource::Fixed(0, r0));
            ValueSource bc = ge.add_calculation(H2B_CALC_MUL, ValueSource::Advice(1, r1), ValueSource::Advice(2, r0));
            ValueSource sum = ge.add_calculation(H2B_CALC_ADD, ValueSource::Advice(0, r0), bc);
            ValueSource diff = ge.add_calculation(H2B_CALC_SUB, sum, ValueSource::Advice(0, rm));
            ValueSource gate = ge.add_calculation(H2B_CALC_MUL, q, diff);
            ValueSource again = ge.add_calculation(H2B_CALC_MUL, q, diff);              // identical calculation: reused
            if (!(again == gate)) { fprintf(stderr, "add_calculation did not reuse an identical calculation\n"); return 1; }
            ValueSource sq = ge.add_calculation(H2B_CALC_SQUARE, gate);
            ValueSource cst = ge.add_constant(fr::from_u64(0x1234567));
            ValueSource scaled = ge.add_calculation(H2B_CALC_MUL, ValueSource::Instance(0, r0), cst);
            ge.add_calculation(H2B_CALC_HORNER, ValueSource::PreviousValue(), ValueSource::Y(), {gate, sq, scaled, ValueSource::Challenge(0)});
            auto cols = read_all<Fr>(dir + "/eval_cols.bin");          // fixed0 | advice0..2 | instance0 | previous values, `rows` each
            auto sc = read_all<Fr>(dir + "/eval_scalars.bin");         // challenge0, beta, gamma, theta, y, delta, deltaomega, last_z
            const size_t rows = cols.size() / 6;
            auto col = [&](size_t j) { return std::vector<Fr>(cols.begin() + j * rows, cols.begin() + (j + 1) * rows); };
            auto out = ge.evaluate({col(0)}, {col(1), col(2), col(3)}, {col(4)}, {sc[0]}, sc[1], sc[2], sc[3], sc[4], col(5), 2);
            write_all(dir + "/eval_out.bin", out.data(), out.size());
            // the same gate built from Expressions through add_expression / Evaluator::new: q * (a + b[next] * c - a[prev]), its square, 0x1234567 * instance
            {
                namespace E = plonk::evaluation;          // (the driver's locals `sum` and `scaled` shadow the builders)
                E::Expr gate_e = E::product(E::fixed(0), E::sum(E::sum(E::advice(0), E::product(E::advice(1, 1), E::advice(2))), E::negated(E::advice(0, -1))));
                GraphEvaluator from_expr = custom_gates_evaluator({gate_e, E::product(gate_e, gate_e), E::scaled(E::instance(0), fr::from_u64(0x1234567)),
                                                                  E::challenge(0)});
                auto out2 = from_expr.evaluate({col(0)}, {col(1), col(2), col(3)}, {col(4)}, {sc[0]}, sc[1], sc[2], sc[3], sc[4], col(5), 2);
                if (std::memcmp(out2.data(), out.data(), out.size() * sizeof(Fr)) != 0) { fprintf(stderr, "add_expression graph differs from the hand-built one\n"); return 1; }
            }
            // grand products over the same columns (rows is a power of two here; omega of that size)
            uint32_t kr = 0;
            while (((size_t)1 << kr) < rows) ++kr;
            poly::EvaluationDomain dr(3, kr);
            auto z = plonk::permutation::commit_product({col(1), col(2)}, {col(3), col(4)}, sc[1], sc[2], sc[5], sc[6], dr.get_omega(), sc[7]);
            write_all(dir + "/perm_z.bin", z.data(), z.size());
            auto zl = plonk::lookup::commit_product(col(1), col(2), col(3), col(4), sc[1], sc[2]);
            write_all(dir + "/lookup_z.bin", zl.data(), zl.size());
            // permute_expression_pair: input / table files hold a satisfiable pair; a foreign input value must fail
            auto lin = read_all<Fr>(dir + "/lookup_input.bin");
            auto ltab = read_all<Fr>(dir + "/lookup_table.bin");
            auto pr = plonk::lookup::permute_expression_pair(lin, ltab, lin.size() - 6);
            write_all(dir + "/permuted_input.bin", pr.first.data(), pr.first.size());
            write_all(dir + "/permuted_table.bin", pr.second.data(), pr.second.size());
            panicked = false;
            try {
                lin[0] = sc[3];
                plonk::lookup::permute_expression_pair(lin, ltab, lin.size() - 6);
            } catch (const Panic&) { panicked = true; }
            if (!panicked) { fprintf(stderr, "missing table value did not fail\n"); return 1; }
        }
    } catch (const Panic& e) {
        fprintf(stderr, "panic: %s\n", e.what());
        return 1;
    }
    printf("HOST_MIRROR_OK\n");
    return 0;
}
