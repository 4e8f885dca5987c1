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
// the blob is copied into a slot of a pinned ring and an async H2D copy into the matching slot of a device ring is enqueued
// on the caller's stream, followed by an event.  A slot is reused only after its own event has completed (a host-side wait
// on that one event — never a device-wide synchronisation), so neither the pinned source nor the device copy can be
// overwritten while a kernel may still read it.  Not capturable in a CUDA graph (a replay would re-read the pinned slot).
struct StagingRing {
    static constexpr size_t SLOT = (size_t)256 << 10;
    static constexpr int NSLOT = 32;
    uint8_t* h = nullptr;
    uint8_t* d = nullptr;
    cudaEvent_t ev[NSLOT] = {};
    bool used[NSLOT] = {};
    int head = 0;
    void* push(const void* src, size_t bytes, cudaStream_t st) {
        if (bytes > SLOT) return nullptr;                          // caller falls back to its synchronous path
        if (!h) {
            if (cudaMallocHost((void**)&h, SLOT * NSLOT) != cudaSuccess || cudaMalloc((void**)&d, SLOT * NSLOT) != cudaSuccess) { cudaGetLastError(); release(); return nullptr; }
            for (int i = 0; i < NSLOT; ++i) if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { release(); return nullptr; }
        }
        const int s = head;
        head = (head + 1) % NSLOT;
        if (used[s] && cudaEventSynchronize(ev[s]) != cudaSuccess) return nullptr;
        memcpy(h + s * SLOT, src, bytes);
        if (cudaMemcpyAsync(d + s * SLOT, h + s * SLOT, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
        if (cudaEventRecord(ev[s], st) != cudaSuccess) return nullptr;
        used[s] = true;
        return d + s * SLOT;
    }
    void release() {
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
        for (int i = 0; i < NSLOT; ++i) { if (ev[i]) cudaEventDestroy(ev[i]); ev[i] = nullptr; used[i] = false; }
        h = d = nullptr; head = 0;
    }
};

// Process-wide tuning knobs, read from the environment ONCE in b200_init (never on a launch path).
struct Config {
    size_t ws_budget_call = 0;      // B200_WS_BUDGET_MB: fixed per-call scratch budget (tests use it to force the batch-split paths); 0 = derive
    size_t ws_budget_total = (size_t)48 << 30;   // scratch the library may hold across all calling threads of a device
    int ntt_v1 = 0;                 // B200_NTT_V=1: radix-2 shared-memory pass everywhere
    int ntt_logg = -1;              // B200_NTT_LOGG
    int ntt_threads = 0;            // B200_NTT_THREADS (v1 pass)
    int ntt_nofull = 0;             // B200_NTT_NOFULL: two-level inter-pass twiddles even when the full table exists
    int msm_reduce_m = 0;           // B200_MSM_REDUCE_M
    int msm_reduce2 = 0;            // B200_MSM_REDUCE2=2: four-lane cooperative reduction tail for <= 3 columns (A/B runs)
    int msm_reduce_threads = 0;     // B200_MSM_REDUCE_THREADS: CTA size of the bucket reduction (32 / 64 / 128 / 256), 0 = automatic
    int shard_min_logn = 22;        // B200_SHARD_MIN_LOGN: a single transform of at least this size is sharded across the devices
};
const Config& config();

// Optional device-side timing of kernel classes with CUDA events on the launching stream (bench.py's roofline leg).
enum ProfClass { PROF_MSM_ACCUMULATE = 0, PROF_MSM_TOTAL = 1, PROF_NTT = 2, PROF_POLY = 3, PROF_MSM_RECODE = 4, PROF_MSM_TAIL = 5, PROF_QUOTIENT = 6, PROF_NCLASS = 7 };
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
