// fp30.cuh — EXPERIMENT (round-2 candidate, not on the product path yet): carry-less Montgomery multiplication for BN254 Fq
// on 9 x 30-bit limbs with 64-bit column accumulators (R' = 2^270).  Every partial product is a plain IMAD.WIDE.U32 with a
// 64-bit addend — no carry flags — which the pipe probes (profiles/r01_pipe_probes.txt) show issuing ~2x faster than the
// carry-chained IMAD.WIDE.U32.X the 8 x 32-bit multiply relies on.  Host + device; checked against bigints in
// tests/test_host_logic.py, timed by b200_debug_bench variant 5.
#pragma once
#include <stdint.h>
#include "field.cuh"

namespace b200 {

struct Fq30 { uint32_t l[9]; };      // value = sum l[i] * 2^(30 i), l[i] < 2^30 when normalised

HD uint32_t fq30_mod(int i) {
    // p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47 in 30-bit limbs
    constexpr uint32_t P[9] = {0x187cfd47u, 0x3082305bu, 0x071ca8d3u, 0x205aa45au, 0x01585d97u, 0x0116da06u, 0x1a029b85u, 0x139cb84cu, 0x00003064u};
    return P[i];
}
static constexpr uint32_t FQ30_INV = 0x24866389u;   // -p^-1 mod 2^30 (the low 30 bits of the 32-bit constant 0xe4866389)
static constexpr uint32_t M30 = 0x3fffffffu;

// r = a * b * 2^-270 mod p, limbs normalised (< 2^30), value < 2p (lazy: callers may feed it straight back in)
HD void fq30_mul(Fq30& r, const Fq30& a, const Fq30& b) {
    uint64_t t[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) t[j] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j < 9; ++j) t[j] += (uint64_t)a.l[j] * b.l[i];
        const uint32_t m = ((uint32_t)t[0] * FQ30_INV) & M30;
#pragma unroll
        for (int j = 0; j < 9; ++j) t[j] += (uint64_t)m * fq30_mod(j);
        // t[0] is now divisible by 2^30: fold it into t[1] and shift the window down one limb
        t[1] += t[0] >> 30;
#pragma unroll
        for (int j = 0; j < 9; ++j) t[j] = t[j + 1];
        t[9] = 0;
        if (i == 3 || i == 7) {          // bound the column sums: propagate carries twice on the way (each column < 2^64 always)
#pragma unroll
            for (int j = 0; j < 8; ++j) { t[j + 1] += t[j] >> 30; t[j] &= M30; }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { t[j + 1] += t[j] >> 30; t[j] &= M30; }
#pragma unroll
    for (int j = 0; j < 9; ++j) r.l[j] = (uint32_t)t[j];
}

}  // namespace b200
