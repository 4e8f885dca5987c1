// debug.cu — small self-test entry points (b200_debug_*) used only by tests/ to localise a failure to one layer
// (field arithmetic, group law, digit recoding) before the composite kernels are blamed.  Not part of the drop-in ABI.
#include "../../include/ezkl_b200.h"
#include <cfenv>
#include <vector>
#include "msm.cuh"
#include "../../tools/experiments/fd.cuh"
#include "../../tools/experiments/msm_affine.cuh"

namespace b200 {
template <class Tag>
__global__ void k_dbg_field(int op, const Fp<Tag>* a, const Fp<Tag>* b, Fp<Tag>* o, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<Tag> x = fp_load(a + i), y = fp_load(b + i), r;
    if (op == 0) r = x + y; else if (op == 1) r = x - y; else if (op == 2) r = x * y; else if (op == 3) r = fp_inv(x); else r = fp_from_mont(x);
    fp_store(o + i, r);
}
// op 0: affine a + affine b; 1: 2*a; 2: k*a (k = b.x.l[0] as small integer); 3: a + a via add_mixed (doubling branch); 4: a + (-a)
__global__ void k_dbg_g1(int op, const G1Affine* a, const G1Affine* b, G1Affine* o, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p = a[i], q = b[i];
    G1Xyzz r;
    if (op == 0) r = g1_add_mixed(g1_to_xyzz(p), q);
    else if (op == 1) r = g1_dbl(g1_to_xyzz(p));
    else if (op == 2) r = g1_mul_small(g1_to_xyzz(p), q.x.l[0]);
    else if (op == 3) r = g1_add(g1_to_xyzz(p), g1_dbl_affine(q));
    else r = g1_add_mixed(g1_to_xyzz(p), g1_neg(p));
    o[i] = g1_to_affine(r);
}
__global__ void k_dbg_digits(const Fr* s, int c, int W, int32_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fp_from_mont(fp_load(s + i));
    uint32_t carry = 0;
    for (int w = 0; w < W; ++w) out[i * W + w] = msm_next_digit(v.l, c, &carry);
}
// ---- throughput microbenchmarks (register-resident loops; results written so nothing is optimised away) ---------
// variant 0: one dependent chain of Fq mulmods (PTX path); 1: two independent chains; 2: portable C++ mul;
// 3: chain of XYZZ mixed additions against a register-resident affine point; 4: Fq add/sub chain
template <int VARIANT>
__global__ void __launch_bounds__(256) k_bench_mul(Fq* out, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = fp_one<FqTag>(), y = fp_one<FqTag>();
    x.l[0] ^= t; y.l[1] ^= (t * 2654435761u);
    if (VARIANT == 0) {
#pragma unroll 1
        for (int i = 0; i < iters; ++i) x = x * y;
    } else if (VARIANT == 1) {
        Fq z = y + y;
#pragma unroll 1
        for (int i = 0; i < iters; i += 2) { x = x * y; z = z * y; }
        x = x + z;
    } else if (VARIANT == 2) {
#pragma unroll 1
        for (int i = 0; i < iters; ++i) { Fq r; fp_mul_portable<FqTag>(r.l, x.l, y.l); x = r; }
    } else if (VARIANT == 3) {
        G1Affine g; g.x = x; g.y = y;
        G1Xyzz acc = g1_dbl_affine(g);
#pragma unroll 1
        for (int i = 0; i < iters; ++i) { acc = g1_add_mixed(acc, g); g.x.l[0] += 1; }
        x = acc.x + acc.y + acc.zz + acc.zzz;
    } else {
#pragma unroll 1
        for (int i = 0; i < iters; ++i) { x = x + y; y = y - x; }
    }
    fp_store(out + t, x);
}
// ---- raw integer-multiply pipe probes: v=0 mad.wide.u32 (no carry), 1 mad.lo.cc/madc.hi.cc pair chains, 2 mad.lo.u32, 3 DFMA
template <int V>
__global__ void __launch_bounds__(256) k_bench_pipe(uint64_t* out, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a = t * 2654435761u + 12345u, b = t ^ 0x9e3779b9u;
    if (V == 0) {
        uint64_t c0 = t, c1 = t + 1, c2 = t + 2, c3 = t + 3, c4 = t + 4, c5 = t + 5, c6 = t + 6, c7 = t + 7;
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
            asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\tmad.wide.u32 %1, %8, %9, %1;\n\tmad.wide.u32 %2, %8, %9, %2;\n\tmad.wide.u32 %3, %8, %9, %3;\n\t"
                         "mad.wide.u32 %4, %8, %9, %4;\n\tmad.wide.u32 %5, %8, %9, %5;\n\tmad.wide.u32 %6, %8, %9, %6;\n\tmad.wide.u32 %7, %8, %9, %7;"
                         : "+l"(c0), "+l"(c1), "+l"(c2), "+l"(c3), "+l"(c4), "+l"(c5), "+l"(c6), "+l"(c7) : "r"(a), "r"(b));
        }
        out[t] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
    } else if (V == 1) {
        uint32_t e0 = t, e1 = 1, e2 = 2, e3 = 3, e4 = 4, e5 = 5, e6 = 6, e7 = 7, o0 = 8, o1 = 9, o2 = 10, o3 = 11, o4 = 12, o5 = 13, o6 = 14, o7 = 15;
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
            asm volatile("mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\tmadc.lo.cc.u32 %2, %16, %17, %2;\n\tmadc.hi.cc.u32 %3, %16, %17, %3;\n\t"
                         "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\tmadc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.u32 %7, %16, %17, %7;\n\t"
                         "mad.lo.cc.u32 %8, %16, %17, %8;\n\tmadc.hi.cc.u32 %9, %16, %17, %9;\n\tmadc.lo.cc.u32 %10, %16, %17, %10;\n\tmadc.hi.cc.u32 %11, %16, %17, %11;\n\t"
                         "madc.lo.cc.u32 %12, %16, %17, %12;\n\tmadc.hi.cc.u32 %13, %16, %17, %13;\n\tmadc.lo.cc.u32 %14, %16, %17, %14;\n\tmadc.hi.u32 %15, %16, %17, %15;"
                         : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "+r"(e6), "+r"(e7), "+r"(o0), "+r"(o1), "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6), "+r"(o7)
                         : "r"(a), "r"(b));
        }
        out[t] = (uint64_t)(e0 ^ e1 ^ e2 ^ e3 ^ e4 ^ e5 ^ e6 ^ e7) << 32 | (o0 ^ o1 ^ o2 ^ o3 ^ o4 ^ o5 ^ o6 ^ o7);
    } else if (V == 2) {
        uint32_t c0 = t, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
            asm volatile("mad.lo.u32 %0, %8, %9, %0;\n\tmad.lo.u32 %1, %8, %9, %1;\n\tmad.lo.u32 %2, %8, %9, %2;\n\tmad.lo.u32 %3, %8, %9, %3;\n\t"
                         "mad.lo.u32 %4, %8, %9, %4;\n\tmad.lo.u32 %5, %8, %9, %5;\n\tmad.lo.u32 %6, %8, %9, %6;\n\tmad.lo.u32 %7, %8, %9, %7;"
                         : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7) : "r"(a), "r"(b));
        }
        out[t] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
    } else {
        double x = (double)a, y = 1.0 + 1e-9 * (double)(b & 1023), c0 = t, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
            c0 = fma(x, y, c0); c1 = fma(x, y, c1); c2 = fma(x, y, c2); c3 = fma(x, y, c3); c4 = fma(x, y, c4); c5 = fma(x, y, c5); c6 = fma(x, y, c6); c7 = fma(x, y, c7);
        }
        out[t] = (uint64_t)(c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7);
    }
}
// ---- FP64-pipe multiplier (fd.cuh) -----------------------------------------------------------------------------------------
// out = fd_mul(a, b) re-sliced back to the 256-bit wire container: a * b * 2^-260 mod N, possibly + N (not canonical)
template <class T>
__global__ void k_dbg_fd_mul(const Fp<typename T::Wire>* a, const Fp<typename T::Wire>* b, Fp<typename T::Wire>* o, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_store(o + i, fd_to_wire(fd_mul(fd_from_wire<T>(fp_load(a + i)), fd_from_wire<T>(fp_load(b + i)))));
}
// throughput: DFMA_WARPS of every 8 warps run a chain of FP64-pipe multiplications, the others the integer-pipe chain
template <int DFMA_WARPS>
__global__ void __launch_bounds__(256) k_bench_hybrid(Fq* out, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = fp_one<FqTag>(), y = fp_one<FqTag>();
    x.l[0] ^= t; y.l[1] ^= (t * 2654435761u);
    if ((int)((threadIdx.x >> 5) & 7) < DFMA_WARPS) {
        Fd<FdFqTag> u = fd_from_wire<FdFqTag>(x), v = fd_from_wire<FdFqTag>(y);
#pragma unroll 1
        for (int i = 0; i < iters; ++i) u = fd_mul(u, v);
        x = fd_to_wire(u);
    } else {
#pragma unroll 1
        for (int i = 0; i < iters; ++i) x = x * y;
    }
    fp_store(out + t, x);
}
}  // namespace b200
using namespace b200;

#pragma GCC visibility push(default)
extern "C" {
// all pointers are HOST pointers; the call stages, runs one kernel and copies back
int b200_debug_field_op(int field, int op, const b200_fr* a, const b200_fr* b, b200_fr* out, size_t n) {
    void *da, *db, *dout;
    B200_CUDA(cudaMalloc(&da, 32 * n)); B200_CUDA(cudaMalloc(&db, 32 * n)); B200_CUDA(cudaMalloc(&dout, 32 * n));
    B200_CUDA(cudaMemcpy(da, a, 32 * n, cudaMemcpyHostToDevice)); B200_CUDA(cudaMemcpy(db, b, 32 * n, cudaMemcpyHostToDevice));
    if (field == 0) k_dbg_field<FrTag><<<div_up(n, 128), 128>>>(op, (const Fr*)da, (const Fr*)db, (Fr*)dout, n);
    else k_dbg_field<FqTag><<<div_up(n, 128), 128>>>(op, (const Fq*)da, (const Fq*)db, (Fq*)dout, n);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpy(out, dout, 32 * n, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return 0;
}
int b200_debug_g1_op(int op, const b200_g1_affine* a, const b200_g1_affine* b, b200_g1_affine* out, size_t n) {
    void *da, *db, *dout;
    B200_CUDA(cudaMalloc(&da, 64 * n)); B200_CUDA(cudaMalloc(&db, 64 * n)); B200_CUDA(cudaMalloc(&dout, 64 * n));
    B200_CUDA(cudaMemcpy(da, a, 64 * n, cudaMemcpyHostToDevice)); B200_CUDA(cudaMemcpy(db, b, 64 * n, cudaMemcpyHostToDevice));
    k_dbg_g1<<<div_up(n, 64), 64>>>(op, (const G1Affine*)da, (const G1Affine*)db, (G1Affine*)dout, n);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpy(out, dout, 64 * n, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return 0;
}
int b200_debug_digits(const b200_fr* s, size_t n, int c, int32_t* out /* n * ceil(255/c) */) {
    const int W = (255 + c - 1) / c;
    void *ds, *dout;
    B200_CUDA(cudaMalloc(&ds, 32 * n)); B200_CUDA(cudaMalloc(&dout, 4 * n * W));
    B200_CUDA(cudaMemcpy(ds, s, 32 * n, cudaMemcpyHostToDevice));
    k_dbg_digits<<<div_up(n, 128), 128>>>((const Fr*)ds, c, W, (int32_t*)dout, n);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpy(out, dout, 4 * n * W, cudaMemcpyDeviceToHost));
    cudaFree(ds); cudaFree(dout);
    return 0;
}
int b200_debug_fd_mul(int field, const b200_fr* a, const b200_fr* b, b200_fr* out, size_t n) {
    void *da, *db, *dout;
    B200_CUDA(cudaMalloc(&da, 32 * n)); B200_CUDA(cudaMalloc(&db, 32 * n)); B200_CUDA(cudaMalloc(&dout, 32 * n));
    B200_CUDA(cudaMemcpy(da, a, 32 * n, cudaMemcpyHostToDevice)); B200_CUDA(cudaMemcpy(db, b, 32 * n, cudaMemcpyHostToDevice));
    if (field == 0) k_dbg_fd_mul<FdFrTag><<<div_up(n, 128), 128>>>((const Fr*)da, (const Fr*)db, (Fr*)dout, n);
    else k_dbg_fd_mul<FdFqTag><<<div_up(n, 128), 128>>>((const Fq*)da, (const Fq*)db, (Fq*)dout, n);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpy(out, dout, 32 * n, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return 0;
}
// host build of the same function (the limb splits are exact in round-toward-zero mode, which this call sets and restores)
int b200_debug_host_fd_mul(int field, const b200_fr* a, const b200_fr* b, b200_fr* out, size_t n) {
    const int old = fegetround();
    fesetround(FE_TOWARDZERO);
    for (size_t i = 0; i < n; ++i) {
        if (field == 0) { Fr x, y; memcpy(&x, &a[i], 32); memcpy(&y, &b[i], 32); Fr r = fd_to_wire(fd_mul(fd_from_wire<FdFrTag>(x), fd_from_wire<FdFrTag>(y))); memcpy(&out[i], &r, 32); }
        else { Fq x, y; memcpy(&x, &a[i], 32); memcpy(&y, &b[i], 32); Fq r = fd_to_wire(fd_mul(fd_from_wire<FdFqTag>(x), fd_from_wire<FdFqTag>(y))); memcpy(&out[i], &r, 32); }
    }
    fesetround(old);
    return 0;
}
// CPU run of the batched-affine accumulation bodies (msm_affine.cuh) on host arrays: out[c] = sum of the chunk's points
int b200_debug_host_affine_chunks(const b200_g1_affine* table, const uint32_t* ents, size_t n_ents, const uint32_t* chunk_start,
                                  const uint32_t* chunk_len, size_t nchunks, b200_g1_affine* out) {
    std::vector<b200::G1Affine> tab, res(nchunks);
    size_t n_pts = 0;
    for (size_t i = 0; i < n_ents; ++i) if ((ents[i] & 0x7fffffffu) + 1 > n_pts) n_pts = (ents[i] & 0x7fffffffu) + 1;
    tab.resize(n_pts);
    memcpy(tab.data(), table, 64 * n_pts);
    int rc = b200::msm_affine_host_chunks(tab.data(), ents, n_ents, chunk_start, chunk_len, nchunks, res.data());
    memcpy(out, res.data(), 64 * nchunks);
    return rc;
}
// returns elapsed ms for `iters` operations per thread on blocks x threads threads
int b200_debug_bench(int variant, int iters, int blocks, int threads, float* ms) {
    Fq* d; B200_CUDA(cudaMalloc(&d, sizeof(Fq) * (size_t)blocks * threads));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        switch (variant) {
            case 0: k_bench_mul<0><<<blocks, threads>>>(d, iters); break;
            case 1: k_bench_mul<1><<<blocks, threads>>>(d, iters); break;
            case 2: k_bench_mul<2><<<blocks, threads>>>(d, iters); break;
            case 3: k_bench_mul<3><<<blocks, threads>>>(d, iters); break;
            case 10: k_bench_hybrid<0><<<blocks, threads>>>(d, iters); break;
            case 11: k_bench_hybrid<1><<<blocks, threads>>>(d, iters); break;
            case 12: k_bench_hybrid<2><<<blocks, threads>>>(d, iters); break;
            case 13: k_bench_hybrid<3><<<blocks, threads>>>(d, iters); break;
            case 14: k_bench_hybrid<4><<<blocks, threads>>>(d, iters); break;
            case 15: k_bench_hybrid<5><<<blocks, threads>>>(d, iters); break;
            case 16: k_bench_hybrid<6><<<blocks, threads>>>(d, iters); break;
            case 17: k_bench_hybrid<7><<<blocks, threads>>>(d, iters); break;
            case 18: k_bench_hybrid<8><<<blocks, threads>>>(d, iters); break;
            default: k_bench_mul<4><<<blocks, threads>>>(d, iters); break;
        }
        cudaEventRecord(e1);
        B200_CUDA(cudaEventSynchronize(e1));
    }
    B200_CUDA(cudaEventElapsedTime(ms, e0, e1));
    cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
// returns elapsed ms; ops per thread per iteration: 8 (v0, v2, v3) or 16 (v1)
int b200_debug_bench_pipe(int variant, int iters, int blocks, int threads, float* ms) {
    uint64_t* d; B200_CUDA(cudaMalloc(&d, 8 * (size_t)blocks * threads));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        switch (variant) {
            case 0: k_bench_pipe<0><<<blocks, threads>>>(d, iters); break;
            case 1: k_bench_pipe<1><<<blocks, threads>>>(d, iters); break;
            case 2: k_bench_pipe<2><<<blocks, threads>>>(d, iters); break;
            default: k_bench_pipe<3><<<blocks, threads>>>(d, iters); break;
        }
        cudaEventRecord(e1);
        B200_CUDA(cudaEventSynchronize(e1));
    }
    B200_CUDA(cudaEventElapsedTime(ms, e0, e1));
    cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
// host-only: the same recoding routine compiled for the CPU (lets the not-gpu tests check it without a device)
int b200_debug_digits_host(const b200_fr* s_canonical, size_t n, int c, int32_t* out) {
    const int W = (255 + c - 1) / c;
    for (size_t i = 0; i < n; ++i) {
        uint32_t v[8]; memcpy(v, &s_canonical[i], 32);
        uint32_t carry = 0;
        for (int w = 0; w < W; ++w) out[i * W + w] = msm_next_digit(v, c, &carry);
    }
    return 0;
}
// host-only: group law / field code compiled for the CPU through the portable path (not-gpu tests)
int b200_debug_host_g1_op(int op, const b200_g1_affine* a, const b200_g1_affine* b, b200_g1_affine* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        G1Affine p, q; memcpy(&p, &a[i], 64); memcpy(&q, &b[i], 64);
        G1Xyzz r;
        if (op == 0) r = g1_add_mixed(g1_to_xyzz(p), q);
        else if (op == 1) r = g1_dbl(g1_to_xyzz(p));
        else if (op == 2) r = g1_mul_small(g1_to_xyzz(p), q.x.l[0]);
        else if (op == 3) r = g1_add(g1_to_xyzz(p), g1_dbl_affine(q));
        else r = g1_add_mixed(g1_to_xyzz(p), g1_neg(p));
        G1Affine o = g1_to_affine(r); memcpy(&out[i], &o, 64);
    }
    return 0;
}
int b200_debug_host_field_op(int field, int op, const b200_fr* a, const b200_fr* b, b200_fr* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (field == 0) { Fr x, y, r; memcpy(&x, &a[i], 32); memcpy(&y, &b[i], 32);
            if (op == 0) r = x + y; else if (op == 1) r = x - y; else if (op == 2) r = x * y; else if (op == 3) r = fp_inv(x); else r = fp_from_mont(x);
            memcpy(&out[i], &r, 32);
        } else { Fq x, y, r; memcpy(&x, &a[i], 32); memcpy(&y, &b[i], 32);
            if (op == 0) r = x + y; else if (op == 1) r = x - y; else if (op == 2) r = x * y; else if (op == 3) r = fp_inv(x); else r = fp_from_mont(x);
            memcpy(&out[i], &r, 32);
        }
    }
    return 0;
}
}
#pragma GCC visibility pop
