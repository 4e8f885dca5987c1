// ntt.cu — BN254 Fr number-theoretic transform for sm_100a.
//
// Replaces halo2_proofs arithmetic.rs best_fft and the EvaluationDomain transforms built on it (lagrange_to_coeff,
// coeff_to_extended, extended_to_coeff; UPSTREAM poly/domain.rs — in-tree user /root/reference/src/circuit/modules/
// polycommit.rs:52, and every column transform of create_proof / keygen_pk, src/pfsys/mod.rs:396,456).
// Semantics are best_fft's: natural-order in, natural-order out, out[j] = sum_i a[i] * omega^(i*j); arithmetic is exact,
// so any factorisation gives the reference's bytes.
//
// Factorisation: N = M1 * M2 (* M3), M <= 1024, one kernel pass per factor.
//   k_ntt_pass2 (default): 256-thread CTAs hold 1024 elements; every thread runs TWO stages of the radix-2 DIF network on a
//     register-resident quad per round, the first round loads global -> registers and the last stores registers -> global
//     (bit reversal, the inter-pass twiddle omega^(j * i_rest) from a full or two-level table, and the post-scale fused), the
//     per-stage twiddle table is staged into shared memory with cp.async.bulk (TMA) + mbarrier.
//   k_ntt_pass (v1): one radix-2 stage per shared-memory round trip; kept for tiny passes and as B200_NTT_V=1.
// Coset pre-scaling (zeta^(i mod 3)), zero padding and the 1/N (and zeta^-(i mod 3)) post-scaling of the extended-domain
// transforms are fused into the first load / last store.  Algorithmic HBM traffic: 64 B per element per transform; this
// schedule moves 64 B per element per PASS (2 passes up to 2^20, 3 above).  The kernels are bound by the 254-bit multiply
// (10-15 per element), not by HBM: see DESIGN.md §4.3.
#include <cstdlib>
#include "ntt.cuh"

namespace b200 {

struct PassArgs {
    const Fr* src; Fr* dst;
    size_t src_pstride, dst_pstride;
    uint32_t logm, log_g, inner_cnt, in_r_fast;
    uint64_t in_rs, in_inner_s, in_outer_s;
    uint64_t out_rs, out_inner_s, out_outer_s;
    uint64_t n_in;
    uint32_t tw_on, rest_is_inner, lo_bits, first, last;
    uint64_t tw_mul;
    const Fr* tw_m; const Fr* t_lo; const Fr* t_hi;
    const uint4* tw_staged;      // v2: per-stage twiddles, planar [2][M] (lo plane, hi plane), stage s at offset 2^s - 1
    const Fr* t_full;            // v2: omega^e for every e < N (single-multiply inter-pass twiddle), or null
    NttScale pre, post;
};

DEV Fr sh_get(const uint4* lo, const uint4* hi, uint32_t i) {
    uint4 a = lo[i], b = hi[i];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
DEV void sh_put(uint4* lo, uint4* hi, uint32_t i, const Fr& v) {
    lo[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    hi[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

__global__ void __launch_bounds__(1024, 1) k_ntt_pass(const PassArgs a) {
    extern __shared__ uint4 sh[];
    const uint32_t M = 1u << a.logm, G = 1u << a.log_g, total = M << a.log_g;
    uint4* dlo = sh; uint4* dhi = sh + total; uint4* tlo = sh + 2 * total; uint4* thi = tlo + (M >> 1);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t tiles_per_outer = a.inner_cnt >> a.log_g;
    const uint32_t outer = blockIdx.x / tiles_per_outer, tile = blockIdx.x % tiles_per_outer;
    const uint32_t inner0 = tile << a.log_g;
    const Fr* src = a.src + (size_t)blockIdx.y * a.src_pstride;
    Fr* dst = a.dst + (size_t)blockIdx.y * a.dst_pstride;
    const uint64_t in_base = (uint64_t)outer * a.in_outer_s + (uint64_t)inner0 * a.in_inner_s;
    const uint64_t out_base = (uint64_t)outer * a.out_outer_s + (uint64_t)inner0 * a.out_inner_s;

    for (uint32_t k = tid; k < (M >> 1); k += nt) { Fr w = fp_load(a.tw_m + k); sh_put(tlo, thi, k, w); }
    for (uint32_t e = tid; e < total; e += nt) {
        uint32_t g, r;
        if (a.in_r_fast) { r = e & (M - 1); g = e >> a.logm; } else { g = e & (G - 1); r = e >> a.log_g; }
        const uint64_t idx = in_base + (uint64_t)g * a.in_inner_s + (uint64_t)r * a.in_rs;
        Fr v = fp_zero<FrTag>();
        if (idx < a.n_in) {
            v = fp_load(src + idx);
            if (a.first) {
                if (a.pre.mode == 1) v = v * a.pre.c[0];
                else if (a.pre.mode == 3) { uint32_t m3 = (uint32_t)(idx % 3); if (m3) v = v * a.pre.c[m3]; }
            }
        }
        sh_put(dlo, dhi, (g << a.logm) + r, v);
    }
    __syncthreads();
    // radix-2 decimation-in-frequency: natural order in, bit-reversed order out
    const uint32_t nbf = total >> 1;
    for (int s = (int)a.logm - 1; s >= 0; --s) {
        const uint32_t half = 1u << s;
        for (uint32_t bf = tid; bf < nbf; bf += nt) {
            const uint32_t g = bf >> (a.logm - 1), b = bf & ((M >> 1) - 1);
            const uint32_t j = b & (half - 1);
            const uint32_t i0 = (g << a.logm) + ((b >> s) << (s + 1)) + j, i1 = i0 + half;
            Fr x = sh_get(dlo, dhi, i0), y = sh_get(dlo, dhi, i1);
            Fr sum = x + y, dif = x - y;
            if (s > 0) { Fr w = sh_get(tlo, thi, j << (a.logm - 1 - s)); dif = dif * w; }
            sh_put(dlo, dhi, i0, sum);
            sh_put(dlo, dhi, i1, dif);
        }
        __syncthreads();
    }
    for (uint32_t e = tid; e < total; e += nt) {
        const uint32_t g = e & (G - 1), rp = e >> a.log_g;
        const uint32_t pos = a.logm ? (__brev(rp) >> (32 - a.logm)) : 0;
        Fr v = sh_get(dlo, dhi, (g << a.logm) + pos);
        if (a.tw_on) {
            const uint64_t i_rest = a.rest_is_inner ? (uint64_t)(inner0 + g) : 0;
            const uint64_t ex = a.tw_mul * (uint64_t)rp * i_rest;
            if (ex) {
                const uint32_t elo = (uint32_t)(ex & ((1ull << a.lo_bits) - 1)), ehi = (uint32_t)(ex >> a.lo_bits);
                Fr w = fp_load(a.t_lo + elo);
                if (ehi) w = w * fp_load(a.t_hi + ehi);
                v = v * w;
            }
        }
        const uint64_t idx = out_base + (uint64_t)g * a.out_inner_s + (uint64_t)rp * a.out_rs;
        if (a.last) {
            if (a.post.mode == 1) v = v * a.post.c[0];
            else if (a.post.mode == 3) v = v * a.post.c[(uint32_t)(idx % 3)];
        }
        fp_store(dst + idx, v);
    }
}


// ---- v2 pass: radix-4 butterflies in registers, one shared-memory exchange per TWO stages --------------------------------
// Each thread owns one "quad" {p, p+q, p+2q, p+3q} per round and runs stages s and s-1 of the DIF network on it in
// registers (an odd log M starts with a single-stage round on the same quad shape).  The first round loads straight from
// global memory into registers and the last round stores straight from registers to global memory (bit-reversal, inter-pass
// twiddle and post-scale fused), so an M-point line touches shared memory ceil(log M / 2) - 1 times instead of log M.
// The per-stage twiddle table is staged into shared memory by the TMA engine (cp.async.bulk + mbarrier, SASS UBLKCP) while
// the first round's global loads are in flight.
DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEV void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
DEV void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

__global__ void __launch_bounds__(256, 3) k_ntt_pass2(const PassArgs a) {
    extern __shared__ uint4 sh[];
    __shared__ uint64_t bar;
    const uint32_t M = 1u << a.logm, G = 1u << a.log_g, total = M << a.log_g, Q = M >> 2;
    uint4* dlo = sh; uint4* dhi = sh + total; uint4* tlo = sh + 2 * total; uint4* thi = tlo + M;
    const uint32_t tid = threadIdx.x;                  // blockDim.x == G * Q
    const uint32_t tiles_per_outer = a.inner_cnt >> a.log_g;
    const uint32_t outer = blockIdx.x / tiles_per_outer, tile = blockIdx.x % tiles_per_outer;
    const uint32_t inner0 = tile << a.log_g;
    const Fr* src = a.src + (size_t)blockIdx.y * a.src_pstride;
    Fr* dst = a.dst + (size_t)blockIdx.y * a.dst_pstride;
    const uint64_t in_base = (uint64_t)outer * a.in_outer_s + (uint64_t)inner0 * a.in_inner_s;
    const uint64_t out_base = (uint64_t)outer * a.out_outer_s + (uint64_t)inner0 * a.out_inner_s;

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, M * 32);
        tma_bulk_g2s(tlo, a.tw_staged, M * 16, &bar);
        tma_bulk_g2s(thi, a.tw_staged + M, M * 16, &bar);
    }

    int s = (int)a.logm - 1;
    bool first = true, tw_ready = false;
    while (s >= 0) {
        const bool two = !(first && (a.logm & 1u));            // odd log M: the first round is a single stage
        const int s_next = two ? s - 2 : s - 1;
        const bool last = s_next < 0;
        // thread -> (line g, quad t): g fastest where the global side is contiguous in g, t fastest otherwise
        uint32_t g, t;
        const bool g_fast = last || (first && !a.in_r_fast);
        if (g_fast) { g = tid & (G - 1); t = tid >> a.log_g; } else { t = tid & (Q - 1); g = tid >> (a.logm - 2); }
        const uint32_t q = 1u << (s - 1), p = t & (q - 1), blk = t >> (s - 1);
        const uint32_t base = (blk << (s + 1)) + p;
        Fr x[4];
        const bool sparse = first && a.first && a.n_in < ((uint64_t)M * a.inner_cnt);      // only the first round of a padded pass 1
        if (first) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t idx = in_base + (uint64_t)g * a.in_inner_s + (uint64_t)(base + j * q) * a.in_rs;
                x[j] = fp_zero<FrTag>();
                if (idx < a.n_in) {
                    x[j] = fp_load(src + idx);
                    if (a.first) {
                        if (a.pre.mode == 1) x[j] = x[j] * a.pre.c[0];
                        else if (a.pre.mode == 3) { uint32_t m3 = (uint32_t)(idx % 3); if (m3) x[j] = x[j] * a.pre.c[m3]; }
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = sh_get(dlo, dhi, (g << a.logm) + base + j * q);
        }
        if (!tw_ready) { mbar_wait(&bar, 0); tw_ready = true; }
        // stage s: pairs (0,2) and (1,3), twiddles T_s[p], T_s[p+q]
        {
            const uint32_t o = (1u << s) - 1u;
            Fr u0 = x[0] + x[2], u2 = x[0] - x[2], u1 = x[1] + x[3], u3 = x[1] - x[3];
            if (s > 0) {
                // zero-padded inputs (coeff_to_extended: n of 2^ext_k coefficients): most first-round operands are zero,
                // and whole warps agree on which, so the multiplications are skipped without divergence
                if (p && !(sparse && fp_is_zero(u2))) u2 = u2 * sh_get(tlo, thi, o + p);
                if (!(sparse && fp_is_zero(u3))) u3 = u3 * sh_get(tlo, thi, o + p + q);
            }
            x[0] = u0; x[1] = u1; x[2] = u2; x[3] = u3;
        }
        if (two) {     // stage s-1: pairs (0,1) and (2,3), twiddle T_{s-1}[p]
            const uint32_t o = (1u << (s - 1)) - 1u;
            Fr v0 = x[0] + x[1], v1 = x[0] - x[1], v2 = x[2] + x[3], v3 = x[2] - x[3];
            if (s - 1 > 0 && p) {
                Fr w = sh_get(tlo, thi, o + p);
                if (!(sparse && fp_is_zero(v1))) v1 = v1 * w;
                if (!(sparse && fp_is_zero(v3))) v3 = v3 * w;
            }
            x[0] = v0; x[1] = v1; x[2] = v2; x[3] = v3;
        }
        if (!last) {
#pragma unroll
            for (int j = 0; j < 4; ++j) sh_put(dlo, dhi, (g << a.logm) + base + j * q, x[j]);     // in place: a quad is owned by one thread per round
            __syncthreads();
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t pos = base + j * q;
                const uint32_t rp = __brev(pos) >> (32 - a.logm);
                Fr v = x[j];
                if (a.tw_on) {
                    const uint64_t i_rest = a.rest_is_inner ? (uint64_t)(inner0 + g) : 0;
                    const uint64_t ex = a.tw_mul * (uint64_t)rp * i_rest;
                    if (ex) {
                        if (a.t_full) v = v * fp_load(a.t_full + ex);
                        else {
                            const uint32_t elo = (uint32_t)(ex & ((1ull << a.lo_bits) - 1)), ehi = (uint32_t)(ex >> a.lo_bits);
                            Fr w = fp_load(a.t_lo + elo);
                            if (ehi) w = w * fp_load(a.t_hi + ehi);
                            v = v * w;
                        }
                    }
                }
                const uint64_t idx = out_base + (uint64_t)g * a.out_inner_s + (uint64_t)rp * a.out_rs;
                if (a.last) {
                    if (a.post.mode == 1) v = v * a.post.c[0];
                    else if (a.post.mode == 3) v = v * a.post.c[(uint32_t)(idx % 3)];
                }
                fp_store(dst + idx, v);
            }
        }
        first = false;
        s = s_next;
    }
}

// staged twiddles for one pass: entry (2^s - 1 + p) = w^(p << (logm - 1 - s)), written planar ([lo plane M][hi plane M])
__global__ void k_stage_twiddles(Fr w, uint32_t logm, uint4* __restrict__ out) {
    const uint32_t M = 1u << logm, idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M) return;
    Fr v = fp_zero<FrTag>();
    if (idx + 1 < M) {
        const uint32_t s = 31 - __clz(idx + 1), p = idx + 1 - (1u << s);
        v = fp_pow_u64(w, (uint64_t)p << (logm - 1 - s));
    }
    out[idx] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    out[M + idx] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// out[i] = base^i with one pow per 32-element run
__global__ void k_powers_run(Fr base, uint64_t count, Fr* __restrict__ out) {
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 32;
    if (i0 >= count) return;
    Fr v = fp_pow_u64(base, i0);
    for (uint64_t i = i0; i < i0 + 32 && i < count; ++i) { fp_store(out + i, v); v = v * base; }
}

// out[i] = base^i  (i < count); twiddle-table builder
__global__ void k_powers(Fr base, uint32_t count, Fr* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) fp_store(out + i, fp_pow_u64(base, (uint64_t)i));
}

static void choose_passes(uint32_t log_n, int* npass, int logm[3]) {
    logm[0] = logm[1] = logm[2] = 0;
    if (log_n <= 10) { *npass = 1; logm[0] = (int)log_n; return; }
    if (log_n <= 20) { *npass = 2; logm[0] = (int)(log_n + 1) / 2; logm[1] = (int)log_n - logm[0]; return; }
    *npass = 3;
    logm[0] = (int)(log_n + 2) / 3; logm[1] = (int)(log_n - logm[0] + 1) / 2; logm[2] = (int)log_n - logm[0] - logm[1];
}
int ntt_launches_per_run(uint32_t log_n) { int np, lm[3]; choose_passes(log_n, &np, lm); return np; }

static Fr host_pow(const Fr& b, uint64_t e) { return fp_pow_u64(b, e); }

NttPlan* NttContext::get(uint32_t log_n, const Fr& omega, cudaStream_t st) {
    for (NttPlan* p : plans) if (p->log_n == log_n && fp_eq(p->omega, omega)) return p;
    NttPlan* p = new NttPlan();
    p->log_n = log_n; p->omega = omega;
    choose_passes(log_n, &p->npass, p->logm);
    const uint64_t N = 1ull << log_n;
    for (int i = 0; i < p->npass; ++i) {
        const uint32_t M = 1u << p->logm[i], cnt = M > 1 ? M / 2 : 1;
        if (cudaMalloc(&p->d_tw[i], sizeof(Fr) * cnt) != cudaSuccess) { set_error("ntt plan: cudaMalloc failed"); delete p; return nullptr; }
        k_powers<<<div_up(cnt, 128), 128, 0, st>>>(host_pow(omega, N / M), cnt, p->d_tw[i]);
    }
    p->lo_bits = (log_n + 1) / 2;
    const uint32_t nlo = 1u << p->lo_bits, nhi = (uint32_t)(N >> p->lo_bits);
    if (cudaMalloc(&p->d_lo, sizeof(Fr) * nlo) != cudaSuccess || cudaMalloc(&p->d_hi, sizeof(Fr) * (nhi ? nhi : 1)) != cudaSuccess) {
        set_error("ntt plan: cudaMalloc failed"); delete p; return nullptr;
    }
    for (int i = 0; i < p->npass; ++i) {
        const uint32_t M = 1u << p->logm[i];
        if (cudaMalloc(&p->d_staged[i], sizeof(uint4) * 2 * M) != cudaSuccess) { set_error("ntt plan: cudaMalloc failed"); delete p; return nullptr; }
        k_stage_twiddles<<<div_up(M, 128), 128, 0, st>>>(host_pow(omega, N / M), (uint32_t)p->logm[i], p->d_staged[i]);
    }
    if (p->npass > 1 && log_n <= 25) {        // full single-multiply twiddle table (N * 32 B; falls back to the two-level table if it does not fit)
        if (cudaMalloc(&p->d_full, sizeof(Fr) * N) != cudaSuccess) { cudaGetLastError(); p->d_full = nullptr; }
        else k_powers_run<<<div_up(div_up(N, 32), 128), 128, 0, st>>>(omega, N, p->d_full);
    }
    k_powers<<<div_up(nlo, 128), 128, 0, st>>>(omega, nlo, p->d_lo);
    k_powers<<<div_up(nhi ? nhi : 1, 128), 128, 0, st>>>(host_pow(omega, 1ull << p->lo_bits), nhi ? nhi : 1, p->d_hi);
    if (cudaGetLastError() != cudaSuccess) { set_error("ntt plan: table kernel launch failed"); delete p; return nullptr; }
    plans.push_back(p);
    return p;
}
void NttContext::release() {
    for (NttPlan* p : plans) {
        for (int i = 0; i < 3; ++i) { if (p->d_tw[i]) cudaFree(p->d_tw[i]); if (p->d_staged[i]) cudaFree(p->d_staged[i]); }
        if (p->d_full) cudaFree(p->d_full);
        if (p->d_lo) cudaFree(p->d_lo);
        if (p->d_hi) cudaFree(p->d_hi);
        delete p;
    }
    plans.clear();
}

static int launch_pass_v1(PassArgs& a, uint64_t lines, int batch, cudaStream_t st);

// v2 launch: G lines per CTA chosen so that a CTA holds 1024 elements (256 threads, one quad each; 3 CTAs per SM)
static int launch_pass(PassArgs& a, uint64_t lines, int batch, cudaStream_t st) {
    const char* ver = getenv("B200_NTT_V");
    if (a.logm < 2 || (ver && atoi(ver) == 1)) return launch_pass_v1(a, lines, batch, st);
    uint32_t log_g = a.logm >= 10 ? 0 : 10 - a.logm;
    if (const char* e = getenv("B200_NTT_LOGG")) log_g = (uint32_t)atoi(e);
    while (log_g > 0 && ((1u << log_g) > a.inner_cnt || (lines >> log_g) * (uint64_t)batch < 296)) --log_g;
    while (log_g > 0 && (a.logm + log_g > 10)) --log_g;
    if (a.logm + log_g < 7) return launch_pass_v1(a, lines, batch, st);       // fewer than 32 quads: not worth a CTA
    a.log_g = log_g;
    const uint32_t threads = 1u << (a.logm + log_g - 2);
    const size_t smem = (((size_t)1 << (a.logm + log_g)) + ((size_t)1 << a.logm)) * 32;
    B200_CHECK(threads <= 256 && smem <= 200 * 1024, -1, "ntt: pass of 2^%u does not fit a CTA", a.logm);
    B200_CUDA(cudaFuncSetAttribute(k_ntt_pass2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid((unsigned)(lines >> log_g), (unsigned)batch);
    k_ntt_pass2<<<grid, threads, smem, st>>>(a);
    B200_CUDA(cudaGetLastError());
    return 0;
}

static int launch_pass_v1(PassArgs& a, uint64_t lines, int batch, cudaStream_t st) {
    // lines per CTA: largest G in {4,2,1} that still yields >= 2 CTAs per SM (and fits shared memory)
    uint32_t log_g = 2;
    while (log_g > 0 && ((1u << log_g) > a.inner_cnt || (lines >> log_g) * (uint64_t)batch < 296)) --log_g;
    while (log_g > 0 && (((size_t)1 << (a.logm + log_g)) + ((size_t)1 << a.logm) / 2) * 32 > 200 * 1024) --log_g;
    if (const char* e = getenv("B200_NTT_LOGG")) { uint32_t v = (uint32_t)atoi(e); while (v > 0 && (1u << v) > a.inner_cnt) --v; log_g = v; }
    a.log_g = log_g;
    const size_t smem = (((size_t)1 << (a.logm + log_g)) + (((size_t)1 << a.logm) >> 1)) * 32;
    const uint32_t nbf = (1u << (a.logm + log_g)) >> 1;
    uint32_t threads = nbf < 32 ? 32 : (nbf > 1024 ? 1024 : nbf);
    if (const char* e = getenv("B200_NTT_THREADS")) { uint32_t v = (uint32_t)atoi(e); if (v >= 32 && v <= 1024 && v < threads) threads = v; }
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA(cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)(lines >> log_g), (unsigned)batch);
    k_ntt_pass<<<grid, threads, smem, st>>>(a);
    B200_CUDA(cudaGetLastError());
    return 0;
}

int ntt_run(NttPlan* p, const Fr* d_src, size_t src_stride, size_t n_in, Fr* d_tmp, size_t tmp_stride, Fr* d_dst, size_t dst_stride,
            uint32_t log_n, const Fr& omega, const NttScale& pre, const NttScale& post, int batch, cudaStream_t st) {
    B200_CHECK(log_n >= 1 && log_n <= 28, -1, "ntt: log_n = %u out of range [1, 28]", log_n);
    B200_CHECK(batch > 0 && batch <= 65535, -1, "ntt: batch %d out of range", batch);
    const uint64_t N = 1ull << log_n;
    B200_CHECK(n_in <= N, -1, "ntt: n_in %zu > N", n_in);
    B200_CHECK(p && p->log_n == log_n && fp_eq(p->omega, omega), -1, "ntt: plan does not match (log_n, omega)");
    ProfScope ps(PROF_NTT, st);
    PassArgs a;
    memset(&a, 0, sizeof a);
    a.pre = pre; a.post = post;
    a.t_lo = p->d_lo; a.t_hi = p->d_hi; a.lo_bits = p->lo_bits; a.t_full = getenv("B200_NTT_NOFULL") ? nullptr : p->d_full;
    if (p->npass == 1) {
        a.src = d_src; a.src_pstride = src_stride; a.dst = d_dst; a.dst_pstride = dst_stride;
        a.logm = p->logm[0]; a.inner_cnt = 1; a.in_r_fast = 1; a.in_rs = 1; a.out_rs = 1; a.n_in = n_in; a.first = a.last = 1;
        a.tw_m = p->d_tw[0]; a.tw_staged = p->d_staged[0];
        return launch_pass(a, 1, batch, st);
    }
    const uint64_t N1 = 1ull << p->logm[0], N2 = 1ull << p->logm[1], N3 = 1ull << p->logm[2];
    if (p->npass == 2) {
        // pass 1: columns i2 (stride 1), transform over i1 (stride N2); twiddle omega^(j1 * i2)
        a.src = d_src; a.src_pstride = src_stride; a.dst = d_tmp; a.dst_pstride = tmp_stride;
        a.logm = p->logm[0]; a.inner_cnt = (uint32_t)N2; a.in_r_fast = 0;
        a.in_rs = N2; a.in_inner_s = 1; a.out_rs = N2; a.out_inner_s = 1; a.n_in = n_in; a.first = 1; a.last = 0;
        a.tw_on = 1; a.rest_is_inner = 1; a.tw_mul = 1; a.tw_m = p->d_tw[0]; a.tw_staged = p->d_staged[0];
        if (int rc = launch_pass(a, N2, batch, st)) return rc;
        // pass 2: rows j1 (stride N2), transform over i2 (stride 1); X[j1 + N1*j2]
        a.src = d_tmp; a.src_pstride = tmp_stride; a.dst = d_dst; a.dst_pstride = dst_stride;
        a.logm = p->logm[1]; a.inner_cnt = (uint32_t)N1; a.in_r_fast = 1;
        a.in_rs = 1; a.in_inner_s = N2; a.out_rs = N1; a.out_inner_s = 1; a.n_in = ~0ull; a.first = 0; a.last = 1;
        a.tw_on = 0; a.tw_m = p->d_tw[1]; a.tw_staged = p->d_staged[1];
        return launch_pass(a, N1, batch, st);
    }
    // three passes: i = i1*N2*N3 + i2*N3 + i3  ->  j = j1 + N1*j2 + N1*N2*j3
    const uint64_t N23 = N2 * N3;
    a.src = d_src; a.src_pstride = src_stride; a.dst = d_tmp; a.dst_pstride = tmp_stride;
    a.logm = p->logm[0]; a.inner_cnt = (uint32_t)N23; a.in_r_fast = 0;
    a.in_rs = N23; a.in_inner_s = 1; a.out_rs = N23; a.out_inner_s = 1; a.n_in = n_in; a.first = 1; a.last = 0;
    a.tw_on = 1; a.rest_is_inner = 1; a.tw_mul = 1; a.tw_m = p->d_tw[0]; a.tw_staged = p->d_staged[0];
    if (int rc = launch_pass(a, N23, batch, st)) return rc;
    // pass 2 (in place on tmp): outer j1 (stride N23), inner i3 (stride 1), transform over i2 (stride N3); twiddle omega^(N1*j2*i3)
    a.src = d_tmp; a.src_pstride = tmp_stride; a.dst = d_tmp; a.dst_pstride = tmp_stride;
    a.logm = p->logm[1]; a.inner_cnt = (uint32_t)N3; a.in_r_fast = 0;
    a.in_rs = N3; a.in_inner_s = 1; a.in_outer_s = N23; a.out_rs = N3; a.out_inner_s = 1; a.out_outer_s = N23; a.n_in = ~0ull; a.first = 0; a.last = 0;
    a.tw_on = 1; a.rest_is_inner = 1; a.tw_mul = N1; a.tw_m = p->d_tw[1]; a.tw_staged = p->d_staged[1];
    if (int rc = launch_pass(a, N1 * N3, batch, st)) return rc;
    // pass 3: outer j2 (stride N3), inner j1 (stride N23), transform over i3 (stride 1)
    a.src = d_tmp; a.src_pstride = tmp_stride; a.dst = d_dst; a.dst_pstride = dst_stride;
    a.logm = p->logm[2]; a.inner_cnt = (uint32_t)N1; a.in_r_fast = 1;
    a.in_rs = 1; a.in_inner_s = N23; a.in_outer_s = N3; a.out_rs = N1 * N2; a.out_inner_s = 1; a.out_outer_s = N1; a.n_in = ~0ull; a.first = 0; a.last = 1;
    a.tw_on = 0; a.tw_m = p->d_tw[2]; a.tw_staged = p->d_staged[2];
    return launch_pass(a, N1 * N2, batch, st);
}

}  // namespace b200
