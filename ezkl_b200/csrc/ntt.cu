// ntt.cu — BN254 Fr number-theoretic transform for sm_100a.
//
// Replaces halo2_proofs arithmetic.rs best_fft and the EvaluationDomain transforms built on it (lagrange_to_coeff,
// coeff_to_extended, extended_to_coeff; UPSTREAM poly/domain.rs — in-tree user /root/reference/src/circuit/modules/
// polycommit.rs:52, and every column transform of create_proof / keygen_pk, src/pfsys/mod.rs:396,456).
// Semantics are best_fft's: natural-order in, natural-order out, out[j] = sum_i a[i] * omega^(i*j); arithmetic is exact,
// so any factorisation gives the reference's bytes.
//
// Factorisation: N = M1 * M2 (* M3), one kernel pass per factor.  Each CTA stages G adjacent "lines" of M elements in
// shared memory (split into two 16-byte planes so warp accesses are conflict-free), runs a radix-2 DIF network on them,
// and writes them back bit-reversal-corrected with the inter-pass twiddle omega^(j * i_rest) (two-level table) fused into
// the store.  Coset pre-scaling (zeta^(i mod 3)), zero padding and the 1/N (and zeta^-(i mod 3)) post-scaling of the
// extended-domain transforms are fused into the first load / last store.  Algorithmic HBM traffic: 64 B per element per
// transform; this schedule moves 64 B per element per PASS (2 passes up to 2^20, 3 above).
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
    k_powers<<<div_up(nlo, 128), 128, 0, st>>>(omega, nlo, p->d_lo);
    k_powers<<<div_up(nhi ? nhi : 1, 128), 128, 0, st>>>(host_pow(omega, 1ull << p->lo_bits), nhi ? nhi : 1, p->d_hi);
    if (cudaGetLastError() != cudaSuccess) { set_error("ntt plan: table kernel launch failed"); delete p; return nullptr; }
    plans.push_back(p);
    return p;
}
void NttContext::release() {
    for (NttPlan* p : plans) {
        for (int i = 0; i < 3; ++i) if (p->d_tw[i]) cudaFree(p->d_tw[i]);
        if (p->d_lo) cudaFree(p->d_lo);
        if (p->d_hi) cudaFree(p->d_hi);
        delete p;
    }
    plans.clear();
}

static int launch_pass(PassArgs& a, uint64_t lines, int batch, cudaStream_t st) {
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

int ntt_run(NttContext& ctx, const Fr* d_src, size_t src_stride, size_t n_in, Fr* d_tmp, size_t tmp_stride, Fr* d_dst, size_t dst_stride,
            uint32_t log_n, const Fr& omega, const NttScale& pre, const NttScale& post, int batch, cudaStream_t st) {
    B200_CHECK(log_n >= 1 && log_n <= 28, -1, "ntt: log_n = %u out of range [1, 28]", log_n);
    B200_CHECK(batch > 0 && batch <= 65535, -1, "ntt: batch %d out of range", batch);
    const uint64_t N = 1ull << log_n;
    B200_CHECK(n_in <= N, -1, "ntt: n_in %zu > N", n_in);
    NttPlan* p = ctx.get(log_n, omega, st);
    if (!p) return -2;
    ProfScope ps(PROF_NTT, st);
    PassArgs a;
    memset(&a, 0, sizeof a);
    a.pre = pre; a.post = post;
    a.t_lo = p->d_lo; a.t_hi = p->d_hi; a.lo_bits = p->lo_bits;
    if (p->npass == 1) {
        a.src = d_src; a.src_pstride = src_stride; a.dst = d_dst; a.dst_pstride = dst_stride;
        a.logm = p->logm[0]; a.inner_cnt = 1; a.in_r_fast = 1; a.in_rs = 1; a.out_rs = 1; a.n_in = n_in; a.first = a.last = 1;
        a.tw_m = p->d_tw[0];
        return launch_pass(a, 1, batch, st);
    }
    const uint64_t N1 = 1ull << p->logm[0], N2 = 1ull << p->logm[1], N3 = 1ull << p->logm[2];
    if (p->npass == 2) {
        // pass 1: columns i2 (stride 1), transform over i1 (stride N2); twiddle omega^(j1 * i2)
        a.src = d_src; a.src_pstride = src_stride; a.dst = d_tmp; a.dst_pstride = tmp_stride;
        a.logm = p->logm[0]; a.inner_cnt = (uint32_t)N2; a.in_r_fast = 0;
        a.in_rs = N2; a.in_inner_s = 1; a.out_rs = N2; a.out_inner_s = 1; a.n_in = n_in; a.first = 1; a.last = 0;
        a.tw_on = 1; a.rest_is_inner = 1; a.tw_mul = 1; a.tw_m = p->d_tw[0];
        if (int rc = launch_pass(a, N2, batch, st)) return rc;
        // pass 2: rows j1 (stride N2), transform over i2 (stride 1); X[j1 + N1*j2]
        a.src = d_tmp; a.src_pstride = tmp_stride; a.dst = d_dst; a.dst_pstride = dst_stride;
        a.logm = p->logm[1]; a.inner_cnt = (uint32_t)N1; a.in_r_fast = 1;
        a.in_rs = 1; a.in_inner_s = N2; a.out_rs = N1; a.out_inner_s = 1; a.n_in = ~0ull; a.first = 0; a.last = 1;
        a.tw_on = 0; a.tw_m = p->d_tw[1];
        return launch_pass(a, N1, batch, st);
    }
    // three passes: i = i1*N2*N3 + i2*N3 + i3  ->  j = j1 + N1*j2 + N1*N2*j3
    const uint64_t N23 = N2 * N3;
    a.src = d_src; a.src_pstride = src_stride; a.dst = d_tmp; a.dst_pstride = tmp_stride;
    a.logm = p->logm[0]; a.inner_cnt = (uint32_t)N23; a.in_r_fast = 0;
    a.in_rs = N23; a.in_inner_s = 1; a.out_rs = N23; a.out_inner_s = 1; a.n_in = n_in; a.first = 1; a.last = 0;
    a.tw_on = 1; a.rest_is_inner = 1; a.tw_mul = 1; a.tw_m = p->d_tw[0];
    if (int rc = launch_pass(a, N23, batch, st)) return rc;
    // pass 2 (in place on tmp): outer j1 (stride N23), inner i3 (stride 1), transform over i2 (stride N3); twiddle omega^(N1*j2*i3)
    a.src = d_tmp; a.src_pstride = tmp_stride; a.dst = d_tmp; a.dst_pstride = tmp_stride;
    a.logm = p->logm[1]; a.inner_cnt = (uint32_t)N3; a.in_r_fast = 0;
    a.in_rs = N3; a.in_inner_s = 1; a.in_outer_s = N23; a.out_rs = N3; a.out_inner_s = 1; a.out_outer_s = N23; a.n_in = ~0ull; a.first = 0; a.last = 0;
    a.tw_on = 1; a.rest_is_inner = 1; a.tw_mul = N1; a.tw_m = p->d_tw[1];
    if (int rc = launch_pass(a, N1 * N3, batch, st)) return rc;
    // pass 3: outer j2 (stride N3), inner j1 (stride N23), transform over i3 (stride 1)
    a.src = d_tmp; a.src_pstride = tmp_stride; a.dst = d_dst; a.dst_pstride = dst_stride;
    a.logm = p->logm[2]; a.inner_cnt = (uint32_t)N1; a.in_r_fast = 1;
    a.in_rs = 1; a.in_inner_s = N23; a.in_outer_s = N3; a.out_rs = N1 * N2; a.out_inner_s = 1; a.out_outer_s = N1; a.n_in = ~0ull; a.first = 0; a.last = 1;
    a.tw_on = 0; a.tw_m = p->d_tw[2];
    return launch_pass(a, N1 * N2, batch, st);
}

}  // namespace b200
