// C++ host-mirror check, run by tests/test_gpu_parity.py::test_cpp_host_mirror on the GPU box:
// reads the reference's SRS fixture, commits the constant-1 Lagrange polynomial (must equal g[0], the generator (1,2)),
// and round-trips a column through lagrange_to_coeff / coeff_to_lagrange and the extended coset.
#include <cstdio>
#include <memory>
#include "../../include/ezkl_b200_halo2.hpp"
using namespace halo2_b200;
int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: test_mirror <kzg_k6.srs>\n"); return 2; }
    try {
        check(b200_init(-1), "b200_init");
        ParamsKZG params = ParamsKZG::read(argv[1]);
        std::vector<Fr> ones(params.n(), fr::ONE);
        G1 c = params.commit_lagrange(ones);
        if (std::memcmp(&c, &params.get_g()[0], 64) != 0 || std::memcmp(&c.z, &(const b200_fq&)fr::ONE, 0) != 0) { std::puts("FAIL commit_lagrange(1) != g[0]"); return 1; }
        EvaluationDomain dom(9, params.k());
        if (dom.extended_k() != 9) { std::puts("FAIL extended_k"); return 1; }
        std::vector<Fr> col(params.n());
        for (size_t i = 0; i < col.size(); ++i) col[i] = fr::from_u64(i * i + 7);
        std::vector<Fr> coeff = col;
        dom.lagrange_to_coeff(coeff);
        std::vector<Fr> back = coeff;
        dom.coeff_to_lagrange(back);
        if (std::memcmp(back.data(), col.data(), 32 * col.size()) != 0) { std::puts("FAIL ntt round trip"); return 1; }
        std::vector<Fr> ext = dom.coeff_to_extended(coeff);
        std::vector<Fr> rec = dom.extended_to_coeff(ext);
        if (rec.size() != params.n() * 8 || std::memcmp(rec.data(), coeff.data(), 32 * coeff.size()) != 0) { std::puts("FAIL extended round trip"); return 1; }
        // p(x) at x = omega^3 equals the Lagrange value col[3]
        Fr x = fr::pow_u64(dom.get_omega(), 3);
        Fr y = eval_polynomial(coeff, x);
        if (std::memcmp(&y, &col[3], 32) != 0) { std::puts("FAIL eval_polynomial"); return 1; }
        std::puts("OK");
        return 0;
    } catch (const std::exception& e) { std::printf("EXCEPTION %s\n", e.what()); return 1; }
}
