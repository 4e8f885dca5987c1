// common.cuh — error plumbing, device scratch arena and block-level helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

namespace b200 {

// Thread-local last-error string behind b200_last_error() (include/ezkl_b200.h).
void set_error(const char* fmt, ...);
const char* get_error();

#define B200_CUDA(call)                                                                                  \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            ::b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));     \
            return -2;                                                                                   \
        }                                                                                                \
    } while (0)

#define B200_CHECK(cond, code, ...)                                                                      \
    do {                                                                                                 \
        if (!(cond)) { ::b200::set_error(__VA_ARGS__); return (code); }                                  \
    } while (0)

// Growable device scratch buffer (one per call-site purpose per context); never shrinks.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); p = nullptr; return -2; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// Small host->device parameter blobs (programs, pointer tables, per-call constants) without a stream synchronisation:
// the blob is copied into a pinned ring and an async H2D copy into the matching slot of a device ring is enqueued on the
// caller's stream.  Slots are only reused after a full lap, and a lap boundary synchronises the device, so neither the
// pinned source nor the device copy can be overwritten while a kernel may still read it.
struct StagingRing {
    uint8_t* h = nullptr;
    uint8_t* d = nullptr;
    size_t cap = 0, head = 0;
    void* push(const void* src, size_t bytes, cudaStream_t st) {
        if (!h) {
            const size_t want = (size_t)8 << 20;
            if (cudaMallocHost((void**)&h, want) != cudaSuccess || cudaMalloc((void**)&d, want) != cudaSuccess) { cudaGetLastError(); h = nullptr; return nullptr; }
            cap = want;
        }
        const size_t need = (bytes + 255) & ~(size_t)255;
        if (need > cap / 4) return nullptr;                       // caller falls back to its synchronous path
        if (head + need > cap) { if (cudaDeviceSynchronize() != cudaSuccess) return nullptr; head = 0; }
        memcpy(h + head, src, bytes);
        if (cudaMemcpyAsync(d + head, h + head, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
        void* out = d + head;
        head += need;
        return out;
    }
};

// Optional device-side timing of kernel classes with CUDA events on the launching stream (bench.py's roofline leg).
enum ProfClass { PROF_MSM_ACCUMULATE = 0, PROF_MSM_TOTAL = 1, PROF_NTT = 2, PROF_POLY = 3, PROF_NCLASS = 4 };
bool prof_enabled();
void prof_mark(int cls, cudaStream_t st, bool begin);
struct ProfScope {
    int cls; cudaStream_t st; bool on;
    ProfScope(int c, cudaStream_t s) : cls(c), st(s), on(prof_enabled()) { if (on) prof_mark(cls, st, true); }
    ~ProfScope() { if (on) prof_mark(cls, st, false); }
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

#if defined(__CUDACC__)
// Exclusive scan of one uint32 per thread across the block (blockDim.x <= 1024, multiple of 32).
// Returns the exclusive prefix; *total gets the block sum (valid in every thread).
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t block_total;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < nw ? warp_sums[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (unsigned)d) wi += t;
        }
        if (lane < nw) warp_sums[lane] = wi - w;
        if (lane == 31) block_total = wi;
    }
    __syncthreads();
    uint32_t r = warp_sums[wid] + incl - v;
    *total = block_total;
    __syncthreads();
    return r;
}
#endif

}  // namespace b200
