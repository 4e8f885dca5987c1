// C++ two-device check, run by tests/test_multi_device.py on a box with >= 2 GPUs: ONE process owns the devices through
// b200_init_multi and uses only the host-pointer C ABI a Rust shim would bind (include/ezkl_b200.h):
//   - a transform large enough to be sharded (B200_SHARD_MIN_LOGN=14 in the environment) must invert exactly,
//   - a batch of transforms dealt over the devices must equal the same transforms done one by one,
//   - one MSM split by base range must equal the same column committed as part of a dealt batch.
#include <cstdio>
#include <cstdlib>
#include "../../include/ezkl_b200_halo2.hpp"
using namespace halo2_b200;
static uint64_t sm(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
int main(int argc, char** argv) {
    const int nd = argc > 1 ? std::atoi(argv[1]) : 2;
    try {
        check(b200_init_multi(nd), "b200_init_multi");
        if (b200_device_count() != nd) { std::puts("FAIL device count"); return 1; }
        const uint32_t k = 16;
        const size_t n = (size_t)1 << k;
        EvaluationDomain dom(5, k);
        uint64_t seed = 0xE2C1B200;
        std::vector<std::vector<Fr>> cols(2 * nd + 1, std::vector<Fr>(n));
        for (auto& c : cols) for (auto& v : c) { Fr t = {{sm(seed), sm(seed), sm(seed), sm(seed) >> 4}}; v = t; }     // < 2^252: valid Montgomery residues
        // sharded single transform round trip
        std::vector<Fr> a = cols[0];
        dom.lagrange_to_coeff(a);
        std::vector<Fr> dealt0 = a;
        dom.coeff_to_lagrange(a);
        if (std::memcmp(a.data(), cols[0].data(), 32 * n) != 0) { std::puts("FAIL sharded round trip"); return 1; }
        // dealt batch == one by one
        std::vector<std::vector<Fr>> batch = cols;
        std::vector<Fr*> ptrs;
        for (auto& c : batch) ptrs.push_back(c.data());
        Fr winv = fr::inv(dom.get_omega()), div = fr::inv(fr::from_u64(n));
        check(b200_ifft_batch(ptrs.data(), ptrs.size(), k, &winv, &div), "ifft_batch");
        if (std::memcmp(batch[0].data(), dealt0.data(), 32 * n) != 0) { std::puts("FAIL dealt batch vs sharded single"); return 1; }
        // MSM: synthetic bases generated on device 0, downloaded, registered (replicated to every device)
        void* d_b = nullptr;
        check(b200_dev_alloc_on(0, &d_b, 64 * n), "dev_alloc");
        check(b200_g1_generate_dev(7, n, d_b, nullptr), "g1_generate");
        check(b200_sync(), "sync");
        std::vector<G1Affine> hb(n);
        check(b200_dev_download(hb.data(), d_b, 64 * n), "download");
        {
        Bases bases(hb.data(), n);
        G1 split = best_multiexp(cols[1], bases);                           // one column: base-split over the devices
        std::vector<const Fr*> cp;
        for (int i = 0; i < 2 * nd; ++i) cp.push_back(cols[1].data());      // the same column 2*nd times: dealt whole
        std::vector<G1> dealt = best_multiexp_batch(cp, n, bases);
        for (auto& g : dealt) if (std::memcmp(&g, &split, sizeof(G1)) != 0) { std::puts("FAIL base-split MSM != dealt MSM"); return 1; }
        }
        b200_shutdown();
        std::puts("OK");
        return 0;
    } catch (const std::exception& e) { std::printf("EXCEPTION %s\n", e.what()); return 1; }
}
