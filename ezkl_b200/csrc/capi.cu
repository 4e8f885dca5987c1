// capi.cu — the extern "C" boundary declared in include/ezkl_b200.h: per-thread contexts, the base-table registry,
// host<->device staging and the host-side tail (point normalisation).  No CPU fallback lives here: every compute entry
// point needs an initialised CUDA device and fails with an error code otherwise.
#include <atomic>
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/ezkl_b200.h"
#include "msm.cuh"
#include "ntt.cuh"
#include "poly.cuh"
#include "quotient.cuh"

namespace b200 {

static thread_local char tl_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_err, sizeof tl_err, fmt, ap);
    va_end(ap);
}
const char* get_error() { return tl_err; }

static std::atomic<int> g_device{-1};
static std::atomic<bool> g_inited{false};
static std::atomic<uint64_t> g_launches{0};
static std::mutex g_mu;
static std::unordered_map<uint64_t, MsmTable*> g_tables;
static uint64_t g_next_handle = 1;
static NttContext g_ntt;           // guarded by g_mu (plan creation) — plans are immutable afterwards
// device scratch budget per call, for batch splitting (B200_WS_BUDGET_MB overrides it — the tests use that to force the split paths)
static size_t ws_budget() {
    static const size_t v = getenv("B200_WS_BUDGET_MB") ? (size_t)atol(getenv("B200_WS_BUDGET_MB")) << 20 : (size_t)12 << 30;
    return v;
}
#define WS_BUDGET (ws_budget())

struct Ctx {
    cudaStream_t stream = nullptr;
    MsmWorkspace msm_ws;
    PolyWorkspace poly_ws;
    QuotientWorkspace quot_ws;
    StagingRing ring;
    DevBuf stage_a, stage_b, stage_c, small;
    bool ok = false;
};
static thread_local Ctx tl_ctx;

static int get_ctx(Ctx** out) {
    if (!g_inited.load()) { set_error("b200: not initialised (call b200_init first)"); return -3; }
    Ctx& c = tl_ctx;
    if (!c.ok) {
        B200_CUDA(cudaSetDevice(g_device.load()));
        B200_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        c.ok = true;
    }
    *out = &c;
    return 0;
}
static inline cudaStream_t pick_stream(Ctx* c, void* user) { return user ? (cudaStream_t)user : c->stream; }
// host Fr values arrive with 8-byte alignment (Rust / C callers); Fr is alignas(16), so always copy bytewise
static inline Fr as_fr(const b200_fr* p) { Fr r; memcpy(&r, p, sizeof r); return r; }

// XYZZ (host) -> normalised Jacobian, one shared inversion (Montgomery's trick over zz*zzz)
static void normalize_host(const G1Xyzz* pts, size_t n, b200_g1_jac* out) {
    std::vector<Fq> prod(n), pref(n);
    Fq acc = fp_one<FqTag>();
    for (size_t i = 0; i < n; ++i) {
        pref[i] = acc;
        if (!g1_is_identity(pts[i])) { prod[i] = pts[i].zz * pts[i].zzz; acc = acc * prod[i]; }
    }
    Fq inv = fp_inv(acc);
    for (size_t i = n; i-- > 0;) {
        G1Jac j;
        if (g1_is_identity(pts[i])) {
            j.x = fp_zero<FqTag>(); j.y = fp_one<FqTag>(); j.z = fp_zero<FqTag>();
        } else {
            Fq t = inv * pref[i];            // 1 / (zz * zzz)
            inv = inv * prod[i];
            Fq zz_inv = t * pts[i].zzz, zzz_inv = t * pts[i].zz;
            j.x = pts[i].x * zz_inv; j.y = pts[i].y * zzz_inv; j.z = fp_one<FqTag>();
        }
        memcpy(&out[i], &j, sizeof j);
    }
}

// ---- profiling (CUDA events on the launching stream) ------------------------------------------------------------
static std::atomic<bool> g_prof{false};
struct ProfRec { int cls; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static thread_local std::vector<ProfRec> tl_prof_open;
bool prof_enabled() { return g_prof.load(std::memory_order_relaxed); }
void prof_mark(int cls, cudaStream_t st, bool begin) {
    if (begin) {
        ProfRec r; r.cls = cls;
        cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        tl_prof_open.push_back(r);
    } else {
        for (size_t i = tl_prof_open.size(); i-- > 0;) {
            if (tl_prof_open[i].cls == cls) {
                ProfRec r = tl_prof_open[i];
                tl_prof_open.erase(tl_prof_open.begin() + i);
                cudaEventRecord(r.e1, st);
                std::lock_guard<std::mutex> lk(g_prof_mu);
                g_prof_recs.push_back(r);
                break;
            }
        }
    }
}

static MsmTable* find_table(uint64_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_tables.find(h);
    return it == g_tables.end() ? nullptr : it->second;
}

static NttPlan* warm_plan(uint32_t log_n, const Fr& omega, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_mu);
    size_t before = g_ntt.plans.size();
    NttPlan* p = g_ntt.get(log_n, omega, st);
    if (p && g_ntt.plans.size() != before) cudaStreamSynchronize(st);   // tables complete before other threads use them
    return p;
}

static int ntt_call(Ctx* c, const Fr* src, size_t src_stride, size_t n_in, Fr* tmp, Fr* dst, size_t dst_stride, uint32_t log_n,
                    const Fr& omega, const NttScale& pre, const NttScale& post, int batch, cudaStream_t st) {
    if (log_n < 1 || log_n > 28) { set_error("ntt: log_n = %u out of range [1, 28]", log_n); return -1; }
    NttPlan* plan = warm_plan(log_n, omega, st);         // looked up / built under g_mu; the vector of plans is never touched unlocked
    if (!plan) return -2;
    int rc = ntt_run(plan, src, src_stride, n_in, tmp, (size_t)1 << log_n, dst, dst_stride, log_n, omega, pre, post, batch, st);
    if (!rc) g_launches += (uint64_t)ntt_launches_per_run(log_n);
    (void)c;
    return rc;
}

}  // namespace b200

using namespace b200;

#pragma GCC visibility push(default)
extern "C" {

int b200_version(void) { return 100; }
const char* b200_last_error(void) { return get_error(); }
uint64_t b200_launch_count(void) { return g_launches.load(); }

int b200_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) { set_error("b200_init: no CUDA device (%s)", cudaGetErrorString(e)); return -2; }
    if (device < 0) { B200_CUDA(cudaGetDevice(&device)); }
    B200_CHECK(device < count, -1, "b200_init: device %d >= device count %d", device, count);
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    B200_CHECK(prop.major == 10, -2, "b200_init: device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    B200_CUDA(cudaSetDevice(device));
    B200_CUDA(cudaFree(0));
    g_device.store(device);
    g_inited.store(true);
    return 0;
}

void b200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_tables) { msm_table_free(kv.second); delete kv.second; }
    g_tables.clear();
    g_ntt.release();
    g_inited.store(false);
}

// ---- profiling -------------------------------------------------------------------------------------------------
int b200_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof_recs.clear();
    g_prof.store(on != 0);
    return 0;
}
int b200_profile_read(int cls, double* total_ms, uint64_t* count) {
    B200_CHECK(cls >= 0 && cls < PROF_NCLASS && total_ms && count, -1, "profile_read: bad argument");
    B200_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double ms = 0; uint64_t n = 0;
    for (auto& r : g_prof_recs) if (r.cls == cls) { float t = 0; if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { ms += t; ++n; } }
    *total_ms = ms; *count = n;
    return 0;
}

// ---- memory helpers ---------------------------------------------------------------------------------------
int b200_dev_alloc(void** d_ptr, size_t bytes) { Ctx* c; if (int rc = get_ctx(&c)) return rc; B200_CUDA(cudaMalloc(d_ptr, bytes)); return 0; }
int b200_dev_free(void* d_ptr) { B200_CUDA(cudaFree(d_ptr)); return 0; }
int b200_dev_upload(void* d_dst, const void* h_src, size_t bytes) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_dev_upload_async(void* d_dst, const void* h_src, size_t bytes, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_dst && h_src, -1, "dev_upload_async: null pointer");
    B200_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, pick_stream(c, stream)));     // truly asynchronous only from pinned memory
    return 0;
}
int b200_dev_download(void* h_dst, const void* d_src, size_t bytes) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_host_alloc(void** h_ptr, size_t bytes) { B200_CUDA(cudaMallocHost(h_ptr, bytes)); return 0; }
int b200_host_free(void* h_ptr) { B200_CUDA(cudaFreeHost(h_ptr)); return 0; }
int b200_sync(void) { Ctx* c; if (int rc = get_ctx(&c)) return rc; B200_CUDA(cudaStreamSynchronize(c->stream)); return 0; }

// ---- bases ---------------------------------------------------------------------------------------------------
int b200_bases_register_dev(const void* d_bases, size_t n, int window_bits, uint64_t* handle) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_bases && handle && n > 0, -1, "bases_register: null argument or n == 0");
    B200_CHECK(window_bits == 0 || (window_bits >= 4 && window_bits <= 24), -1, "bases_register: window_bits %d not in {0, 4..24}", window_bits);
    MsmTable* t = new MsmTable();
    int rc = msm_table_build(t, reinterpret_cast<const G1Affine*>(d_bases), n, window_bits, c->stream);
    if (rc) { msm_table_free(t); delete t; return rc; }
    g_launches += (uint64_t)(t->W - 1);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { set_error("bases_register: %s", cudaGetErrorString(e)); msm_table_free(t); delete t; return -2; }
    std::lock_guard<std::mutex> lk(g_mu);
    *handle = g_next_handle++;
    g_tables[*handle] = t;
    return 0;
}
int b200_bases_register(const b200_g1_affine* bases, size_t n, int window_bits, uint64_t* handle) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(bases && handle && n > 0, -1, "bases_register: null argument or n == 0");
    if (c->stage_a.ensure(sizeof(G1Affine) * n)) return -2;
    B200_CUDA(cudaMemcpyAsync(c->stage_a.p, bases, sizeof(G1Affine) * n, cudaMemcpyHostToDevice, c->stream));
    return b200_bases_register_dev(c->stage_a.p, n, window_bits, handle);
}
int b200_bases_release(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_tables.find(handle);
    B200_CHECK(it != g_tables.end(), -1, "bases_release: unknown handle %llu", (unsigned long long)handle);
    msm_table_free(it->second); delete it->second;
    g_tables.erase(it);
    return 0;
}
int b200_bases_info(uint64_t handle, size_t* n, int* window_bits, int* windows) {
    MsmTable* t = find_table(handle);
    B200_CHECK(t, -1, "bases_info: unknown handle %llu", (unsigned long long)handle);
    if (n) *n = t->n;
    if (window_bits) *window_bits = t->c;
    if (windows) *windows = t->W;
    return 0;
}

// ---- MSM -----------------------------------------------------------------------------------------------------
int b200_msm_batch_dev(uint64_t bases, const void* d_scalars, size_t n, size_t stride, size_t batch, void* d_out_xyzz, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    MsmTable* t = find_table(bases);
    B200_CHECK(t, -1, "msm: unknown bases handle %llu", (unsigned long long)bases);
    B200_CHECK(d_scalars && d_out_xyzz, -1, "msm: null pointer");
    B200_CHECK(n <= t->n, -1, "msm: %zu scalars per column, %zu bases registered", n, t->n);
    B200_CHECK(batch <= 1 || stride >= n, -1, "msm: column stride %zu < column length %zu", stride, n);
    if (batch == 0) return 0;
    cudaStream_t st = pick_stream(c, stream);
    const size_t per_col = msm_workspace_per_column(*t, n);
    size_t sub = WS_BUDGET / (per_col ? per_col : 1);
    if (sub < 1) sub = 1;
    if (sub > 4096) sub = 4096;
    const Fr* sc = reinterpret_cast<const Fr*>(d_scalars);
    G1Xyzz* out = reinterpret_cast<G1Xyzz*>(d_out_xyzz);
    for (size_t b0 = 0; b0 < batch; b0 += sub) {
        const size_t nb = batch - b0 < sub ? batch - b0 : sub;
        if (int rc = msm_run(*t, sc + b0 * stride, n, stride, (int)nb, out + b0, c->msm_ws, st)) return rc;
        g_launches += (uint64_t)msm_launches_per_run();
    }
    return 0;
}
int b200_msm_batch(uint64_t bases, const b200_fr* const* scalars, size_t n, size_t batch, b200_g1_jac* out) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(scalars && out, -1, "msm: null pointer");
    if (batch == 0) return 0;
    MsmTable* t = find_table(bases);
    B200_CHECK(t, -1, "msm: unknown bases handle %llu", (unsigned long long)bases);
    B200_CHECK(n <= t->n, -1, "msm: %zu scalars but only %zu bases registered", n, t->n);
    std::vector<G1Xyzz> h(batch);
    if (n == 0) { memset(h.data(), 0, sizeof(G1Xyzz) * batch); normalize_host(h.data(), batch, out); return 0; }
    if (c->stage_a.ensure(sizeof(Fr) * n * batch)) return -2;
    if (c->small.ensure(sizeof(G1Xyzz) * batch)) return -2;
    for (size_t b = 0; b < batch; ++b) {
        B200_CHECK(scalars[b], -1, "msm: scalars[%zu] is null", b);
        B200_CUDA(cudaMemcpyAsync(c->stage_a.as<Fr>() + b * n, scalars[b], sizeof(Fr) * n, cudaMemcpyHostToDevice, c->stream));
    }
    if (int rc = b200_msm_batch_dev(bases, c->stage_a.p, n, n, batch, c->small.p, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(h.data(), c->small.p, sizeof(G1Xyzz) * batch, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    normalize_host(h.data(), batch, out);
    return 0;
}
int b200_msm(uint64_t bases, const b200_fr* scalars, size_t n, b200_g1_jac* out) {
    const b200_fr* cols[1] = {scalars};
    return b200_msm_batch(bases, cols, n, 1, out);
}
int b200_g1_sum_dev(const void* d_points_xyzz, size_t groups, size_t count, void* d_out_xyzz, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_points_xyzz && d_out_xyzz, -1, "g1_sum: null pointer");
    int rc = g1_sum_run(reinterpret_cast<const G1Xyzz*>(d_points_xyzz), groups, count, reinterpret_cast<G1Xyzz*>(d_out_xyzz), pick_stream(c, stream));
    if (!rc) g_launches += 1;
    return rc;
}
int b200_g1_fft_dev(const void* d_in_affine, uint32_t log_n, const b200_fr* omega, const b200_fr* scale, void* d_out_affine, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_in_affine && omega && d_out_affine, -1, "g1_fft: null pointer");
    const Fr w = as_fr(omega);
    Fr sc = fp_one<FrTag>();
    if (scale) sc = as_fr(scale);
    int rc = g1_fft_run(reinterpret_cast<const G1Affine*>(d_in_affine), log_n, w, scale ? &sc : nullptr, reinterpret_cast<G1Affine*>(d_out_affine), c->msm_ws.misc, pick_stream(c, stream));
    if (!rc) g_launches += (uint64_t)g1_fft_launches(log_n);
    return rc;
}
int b200_g1_fft(const b200_g1_affine* in, uint32_t log_n, const b200_fr* omega, const b200_fr* scale, b200_g1_affine* out) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(in && omega && out && log_n <= 26, -1, "g1_fft: bad argument");
    const size_t n = (size_t)1 << log_n;
    if (c->stage_a.ensure(sizeof(G1Affine) * n) || c->stage_b.ensure(sizeof(G1Affine) * n)) return -2;
    B200_CUDA(cudaMemcpyAsync(c->stage_a.p, in, sizeof(G1Affine) * n, cudaMemcpyHostToDevice, c->stream));
    if (int rc = b200_g1_fft_dev(c->stage_a.p, log_n, omega, scale, c->stage_b.p, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(out, c->stage_b.p, sizeof(G1Affine) * n, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_g1_fixed_base_mul_dev(const void* d_scalars, size_t n, const b200_g1_affine* base, void* d_out_affine, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_scalars && base && d_out_affine, -1, "g1_fixed_base_mul: null pointer");
    G1Affine b; memcpy(&b, base, sizeof b);
    int rc = g1_fixed_base_mul_run(reinterpret_cast<const Fr*>(d_scalars), n, b, reinterpret_cast<G1Affine*>(d_out_affine), pick_stream(c, stream));
    if (!rc && n) g_launches += 1;
    return rc;
}
int b200_g1_generate_dev(uint64_t seed, size_t n, void* d_out_affine, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_out_affine, -1, "g1_generate: null pointer");
    int rc = g1_generate_run(seed, n, reinterpret_cast<G1Affine*>(d_out_affine), pick_stream(c, stream));
    if (!rc && n) g_launches += 1;
    return rc;
}
int b200_g1_normalize(const b200_g1_xyzz* points, size_t n, b200_g1_jac* out) {
    B200_CHECK(points && out, -1, "g1_normalize: null pointer");
    if (n == 0) return 0;
    std::vector<G1Xyzz> tmp(n);
    memcpy(tmp.data(), points, sizeof(G1Xyzz) * n);
    normalize_host(tmp.data(), n, out);
    return 0;
}

// ---- NTT -----------------------------------------------------------------------------------------------------
int b200_ntt_dev(const void* d_src, size_t src_stride, size_t n_in, void* d_tmp, void* d_dst, size_t dst_stride, uint32_t log_n,
                 const b200_fr* omega, int pre_mode, const b200_fr* pre, int post_mode, const b200_fr* post, size_t batch, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_src && d_tmp && d_dst && omega, -1, "ntt: null pointer");
    B200_CHECK((pre_mode == 0 || pre_mode == 1 || pre_mode == 3) && (post_mode == 0 || post_mode == 1 || post_mode == 3), -1, "ntt: scale mode must be 0, 1 or 3");
    B200_CHECK((pre_mode == 0 || pre) && (post_mode == 0 || post), -1, "ntt: scale constants missing");
    if (batch == 0) return 0;
    NttScale a, b;
    a.mode = pre_mode; b.mode = post_mode;
    for (int i = 0; i < pre_mode; ++i) a.c[i] = as_fr(pre + i);
    for (int i = 0; i < post_mode; ++i) b.c[i] = as_fr(post + i);
    return ntt_call(c, reinterpret_cast<const Fr*>(d_src), src_stride, n_in, reinterpret_cast<Fr*>(d_tmp), reinterpret_cast<Fr*>(d_dst), dst_stride,
                    log_n, as_fr(omega), a, b, (int)batch, pick_stream(c, stream));
}

// shared host path: batch polynomials src[p] (n_in each) -> dst[p] (2^log_n each)
static int ntt_host(const b200_fr* const* src, b200_fr* const* dst, size_t batch, size_t n_in, uint32_t log_n, const Fr& omega,
                    const NttScale& pre, const NttScale& post) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(log_n >= 1 && log_n <= 28, -1, "ntt: log_n = %u out of range [1, 28]", log_n);
    const size_t N = (size_t)1 << log_n;
    B200_CHECK(n_in <= N, -1, "ntt: %zu input elements > 2^%u", n_in, log_n);
    if (batch == 0) return 0;
    size_t sub = WS_BUDGET / (sizeof(Fr) * N * 3);
    if (sub < 1) sub = 1;
    if (sub > batch) sub = batch;
    if (c->stage_a.ensure(sizeof(Fr) * n_in * sub) || c->stage_b.ensure(sizeof(Fr) * N * sub) || c->stage_c.ensure(sizeof(Fr) * N * sub)) return -2;
    for (size_t b0 = 0; b0 < batch; b0 += sub) {
        const size_t nb = batch - b0 < sub ? batch - b0 : sub;
        for (size_t p = 0; p < nb; ++p) {
            B200_CHECK(src[b0 + p] && dst[b0 + p], -1, "ntt: polynomial %zu is null", b0 + p);
            B200_CUDA(cudaMemcpyAsync(c->stage_a.as<Fr>() + p * n_in, src[b0 + p], sizeof(Fr) * n_in, cudaMemcpyHostToDevice, c->stream));
        }
        if (int rc = ntt_call(c, c->stage_a.as<Fr>(), n_in, n_in, c->stage_b.as<Fr>(), c->stage_c.as<Fr>(), N, log_n, omega, pre, post, (int)nb, c->stream)) return rc;
        for (size_t p = 0; p < nb; ++p)
            B200_CUDA(cudaMemcpyAsync(dst[b0 + p], c->stage_c.as<Fr>() + p * N, sizeof(Fr) * N, cudaMemcpyDeviceToHost, c->stream));
        B200_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}
int b200_fft_batch(b200_fr* const* a, size_t batch, uint32_t log_n, const b200_fr* omega) {
    B200_CHECK(a && omega, -1, "fft: null pointer");
    NttScale none;
    return ntt_host(a, a, batch, (size_t)1 << (log_n <= 28 ? log_n : 0), log_n, as_fr(omega), none, none);
}
int b200_fft(b200_fr* a, uint32_t log_n, const b200_fr* omega) { b200_fr* p[1] = {a}; return b200_fft_batch(p, 1, log_n, omega); }
int b200_ifft_batch(b200_fr* const* a, size_t batch, uint32_t log_n, const b200_fr* omega_inv, const b200_fr* divisor) {
    B200_CHECK(a && omega_inv && divisor, -1, "ifft: null pointer");
    NttScale none, post;
    post.mode = 1; post.c[0] = as_fr(divisor);
    return ntt_host(a, a, batch, (size_t)1 << (log_n <= 28 ? log_n : 0), log_n, as_fr(omega_inv), none, post);
}
int b200_ifft(b200_fr* a, uint32_t log_n, const b200_fr* omega_inv, const b200_fr* divisor) {
    b200_fr* p[1] = {a};
    return b200_ifft_batch(p, 1, log_n, omega_inv, divisor);
}
int b200_coeff_to_extended_batch(const b200_fr* const* coeffs, size_t batch, size_t n_coeffs, uint32_t ext_k, const b200_fr* ext_omega, const b200_fr* zeta, b200_fr* const* out) {
    B200_CHECK(coeffs && out && ext_omega && zeta, -1, "coeff_to_extended: null pointer");
    NttScale pre, none;
    pre.mode = 3; pre.c[0] = fp_one<FrTag>(); pre.c[1] = as_fr(zeta); pre.c[2] = as_fr(zeta) * as_fr(zeta);
    return ntt_host(coeffs, out, batch, n_coeffs, ext_k, as_fr(ext_omega), pre, none);
}
int b200_coeff_to_extended(const b200_fr* coeffs, size_t n_coeffs, uint32_t ext_k, const b200_fr* ext_omega, const b200_fr* zeta, b200_fr* out) {
    const b200_fr* s[1] = {coeffs}; b200_fr* d[1] = {out};
    return b200_coeff_to_extended_batch(s, 1, n_coeffs, ext_k, ext_omega, zeta, d);
}
int b200_extended_to_coeff(b200_fr* a, uint32_t ext_k, const b200_fr* ext_omega_inv, const b200_fr* ext_ifft_divisor, const b200_fr* zeta) {
    B200_CHECK(a && ext_omega_inv && ext_ifft_divisor && zeta, -1, "extended_to_coeff: null pointer");
    NttScale none, post;
    const Fr z = as_fr(zeta), z2 = z * z, d = as_fr(ext_ifft_divisor);
    post.mode = 3; post.c[0] = d; post.c[1] = d * z2; post.c[2] = d * z;     // zeta^-1 = zeta^2
    b200_fr* p[1] = {a};
    return ntt_host(p, p, 1, (size_t)1 << (ext_k <= 28 ? ext_k : 0), ext_k, as_fr(ext_omega_inv), none, post);
}

// ---- polynomial ops --------------------------------------------------------------------------------------------
int b200_poly_op_dev(int op, const void* d_a, const void* d_b, const b200_fr* s, void* d_out, size_t n, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(op >= 0 && op <= 4, -1, "poly_op: unknown op %d", op);
    B200_CHECK(d_a && d_out && (op == POLY_SCALE || d_b) && (op < POLY_SCALE || s), -1, "poly_op: missing operand for op %d", op);
    Fr sv = fp_zero<FrTag>();
    if (s) sv = as_fr(s);
    int rc = poly_binary(op, reinterpret_cast<const Fr*>(d_a), reinterpret_cast<const Fr*>(d_b), s ? &sv : nullptr, reinterpret_cast<Fr*>(d_out), n, pick_stream(c, stream));
    if (!rc && n) g_launches += 1;
    return rc;
}
int b200_poly_op(int op, const b200_fr* a, const b200_fr* b, const b200_fr* s, b200_fr* out, size_t n) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(a && out, -1, "poly_op: null pointer");
    if (n == 0) return 0;
    const bool need_b = op != POLY_SCALE;
    B200_CHECK(!need_b || b, -1, "poly_op: missing operand b");
    if (c->stage_a.ensure(sizeof(Fr) * n) || (need_b && c->stage_b.ensure(sizeof(Fr) * n))) return -2;
    B200_CUDA(cudaMemcpyAsync(c->stage_a.p, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, c->stream));
    if (need_b) B200_CUDA(cudaMemcpyAsync(c->stage_b.p, b, sizeof(Fr) * n, cudaMemcpyHostToDevice, c->stream));
    if (int rc = b200_poly_op_dev(op, c->stage_a.p, need_b ? c->stage_b.p : nullptr, s, c->stage_a.p, n, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(out, c->stage_a.p, sizeof(Fr) * n, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_poly_lincomb_dev(const void* const* d_polys, const b200_fr* scalars, size_t count, size_t n, void* d_out, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_out && (count == 0 || (d_polys && scalars)), -1, "poly_lincomb: null pointer");
    std::vector<Fr> sv(count);
    if (count) memcpy(sv.data(), scalars, sizeof(Fr) * count);
    int rc = poly_lincomb(reinterpret_cast<const Fr* const*>(d_polys), sv.data(), count, reinterpret_cast<Fr*>(d_out), n, c->poly_ws, pick_stream(c, stream));
    if (!rc && n) g_launches += 1;
    return rc;
}
int b200_poly_lincomb(const b200_fr* const* polys, const b200_fr* scalars, size_t count, size_t n, b200_fr* out) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(out && (count == 0 || (polys && scalars)), -1, "poly_lincomb: null pointer");
    if (n == 0) return 0;
    if (c->stage_a.ensure(sizeof(Fr) * n * (count ? count : 1)) || c->stage_b.ensure(sizeof(Fr) * n)) return -2;
    std::vector<const void*> ptrs(count);
    for (size_t j = 0; j < count; ++j) {
        B200_CHECK(polys[j], -1, "poly_lincomb: polys[%zu] is null", j);
        ptrs[j] = c->stage_a.as<Fr>() + j * n;
        B200_CUDA(cudaMemcpyAsync(c->stage_a.as<Fr>() + j * n, polys[j], sizeof(Fr) * n, cudaMemcpyHostToDevice, c->stream));
    }
    if (int rc = b200_poly_lincomb_dev(ptrs.data(), scalars, count, n, c->stage_b.p, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(out, c->stage_b.p, sizeof(Fr) * n, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_poly_scale_cycle_dev(void* d_a, size_t n, const b200_fr* consts, uint32_t period, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_a && consts && period > 0 && period <= 1024, -1, "poly_scale_cycle: bad argument");
    cudaStream_t st = pick_stream(c, stream);
    const Fr* d_consts = reinterpret_cast<const Fr*>(c->ring.push(consts, sizeof(Fr) * period, st));
    if (!d_consts) {
        if (c->small.ensure(sizeof(Fr) * period)) return -2;
        B200_CUDA(cudaMemcpyAsync(c->small.p, consts, sizeof(Fr) * period, cudaMemcpyHostToDevice, st));
        d_consts = c->small.as<Fr>();
    }
    int rc = poly_scale_cycle(reinterpret_cast<const Fr*>(d_a), d_consts, period, reinterpret_cast<Fr*>(d_a), n, st);
    if (!rc && n) g_launches += 1;
    return rc;
}
int b200_poly_scale_cycle(b200_fr* a, size_t n, const b200_fr* consts, uint32_t period) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(a && consts, -1, "poly_scale_cycle: null pointer");
    if (n == 0) return 0;
    if (c->stage_a.ensure(sizeof(Fr) * n)) return -2;
    B200_CUDA(cudaMemcpyAsync(c->stage_a.p, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, c->stream));
    if (int rc = b200_poly_scale_cycle_dev(c->stage_a.p, n, consts, period, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(a, c->stage_a.p, sizeof(Fr) * n, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_poly_eval_batch_dev(const void* d_polys, size_t stride, size_t n, const b200_fr* x, size_t batch, void* d_out, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_polys && x && d_out, -1, "poly_eval: null pointer");
    if (batch == 0) return 0;
    std::vector<Fr> xv(batch);
    memcpy(xv.data(), x, sizeof(Fr) * batch);
    int rc = poly_eval(reinterpret_cast<const Fr*>(d_polys), stride, n, xv.data(), reinterpret_cast<Fr*>(d_out), (int)batch, c->poly_ws, pick_stream(c, stream));
    if (!rc && n) g_launches += 2;
    return rc;
}
int b200_poly_eval_batch(const b200_fr* const* polys, size_t n, const b200_fr* x, size_t batch, b200_fr* out) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(polys && x && out, -1, "poly_eval: null pointer");
    if (batch == 0) return 0;
    if (c->stage_a.ensure(sizeof(Fr) * (n ? n : 1) * batch) || c->small.ensure(sizeof(Fr) * batch)) return -2;
    for (size_t p = 0; p < batch; ++p) {
        B200_CHECK(n == 0 || polys[p], -1, "poly_eval: polys[%zu] is null", p);
        if (n) B200_CUDA(cudaMemcpyAsync(c->stage_a.as<Fr>() + p * n, polys[p], sizeof(Fr) * n, cudaMemcpyHostToDevice, c->stream));
    }
    if (int rc = b200_poly_eval_batch_dev(c->stage_a.p, n, n, x, batch, c->small.p, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(out, c->small.p, sizeof(Fr) * batch, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_poly_eval(const b200_fr* coeffs, size_t n, const b200_fr* x, b200_fr* out) {
    const b200_fr* p[1] = {coeffs};
    return b200_poly_eval_batch(p, n, x, 1, out);
}
int b200_batch_invert_dev(void* d_a, size_t n, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_a, -1, "batch_invert: null pointer");
    int rc = poly_batch_invert(reinterpret_cast<Fr*>(d_a), n, c->poly_ws, pick_stream(c, stream));
    if (!rc && n) g_launches += 1;
    return rc;
}
int b200_batch_invert(b200_fr* a, size_t n) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(a, -1, "batch_invert: null pointer");
    if (n == 0) return 0;
    if (c->stage_a.ensure(sizeof(Fr) * n)) return -2;
    B200_CUDA(cudaMemcpyAsync(c->stage_a.p, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, c->stream));
    if (int rc = b200_batch_invert_dev(c->stage_a.p, n, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(a, c->stage_a.p, sizeof(Fr) * n, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_prefix_scan_dev(int product, const void* d_a, size_t n, const b200_fr* init, void* d_out, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_a && init && d_out, -1, "prefix_scan: null pointer");
    const Fr iv = as_fr(init);
    int rc = poly_prefix_scan(product != 0, reinterpret_cast<const Fr*>(d_a), n, &iv, reinterpret_cast<Fr*>(d_out), c->poly_ws, pick_stream(c, stream));
    if (!rc && n) g_launches += 3;
    return rc;
}
int b200_prefix_scan(int product, const b200_fr* a, size_t n, const b200_fr* init, b200_fr* out) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(a && init && out, -1, "prefix_scan: null pointer");
    if (n == 0) return 0;
    if (c->stage_a.ensure(sizeof(Fr) * n) || c->stage_b.ensure(sizeof(Fr) * n)) return -2;
    B200_CUDA(cudaMemcpyAsync(c->stage_a.p, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, c->stream));
    if (int rc = b200_prefix_scan_dev(product, c->stage_a.p, n, init, c->stage_b.p, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(out, c->stage_b.p, sizeof(Fr) * n, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int b200_kate_division_dev(const void* d_a, size_t n, const b200_fr* b, void* d_q, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_a && b && d_q, -1, "kate_division: null pointer");
    const Fr bv = as_fr(b);
    int rc = poly_kate_division(reinterpret_cast<const Fr*>(d_a), n, &bv, reinterpret_cast<Fr*>(d_q), c->poly_ws, pick_stream(c, stream));
    if (!rc && n > 1) g_launches += 3;
    return rc;
}
int b200_kate_division(const b200_fr* a, size_t n, const b200_fr* b, b200_fr* q) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(a && b && q, -1, "kate_division: null pointer");
    B200_CHECK(n >= 1, -1, "kate_division: empty polynomial");
    if (n == 1) return 0;
    if (c->stage_a.ensure(sizeof(Fr) * n) || c->stage_b.ensure(sizeof(Fr) * n)) return -2;
    B200_CUDA(cudaMemcpyAsync(c->stage_a.p, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, c->stream));
    if (int rc = b200_kate_division_dev(c->stage_a.p, n, b, c->stage_b.p, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(q, c->stage_b.p, sizeof(Fr) * (n - 1), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- quotient numerator (evaluate_h) ------------------------------------------------------------------------------
static_assert(sizeof(b200_instr) == sizeof(QInstr) && sizeof(b200_col_ref) == sizeof(QLoad), "ABI structs must match the kernel's");
int b200_quotient_eval_dev(const void* const* d_columns, size_t n_columns, uint32_t k, uint32_t ext_k, const b200_col_ref* loads, size_t n_loads,
                           const b200_fr* constants, size_t n_constants, const b200_instr* program, size_t n_instr, void* d_out, void* stream) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(d_out && (n_columns == 0 || d_columns) && (n_loads == 0 || loads) && (n_constants == 0 || constants) && (n_instr == 0 || program), -1, "quotient_eval: null pointer");
    B200_CHECK(ext_k >= k && ext_k <= 28, -1, "quotient_eval: need k <= ext_k <= 28");
    const uint64_t N = 1ull << ext_k, scale = 1ull << (ext_k - k);
    std::vector<QLoad> ql(n_loads);
    for (size_t i = 0; i < n_loads; ++i) {
        ql[i].column = loads[i].column;
        const int64_t off = (int64_t)loads[i].rotation * (int64_t)scale;           // Rotation(r) on the extended domain = r * 2^(ext_k - k)
        ql[i].offset = (uint32_t)(((off % (int64_t)N) + (int64_t)N) % (int64_t)N);
    }
    int rc = quotient_eval_run(reinterpret_cast<const Fr* const*>(d_columns), n_columns, ext_k, ql.data(), n_loads, reinterpret_cast<const Fr*>(constants), n_constants,
                               reinterpret_cast<const QInstr*>(program), n_instr, reinterpret_cast<Fr*>(d_out), c->quot_ws, pick_stream(c, stream));
    if (!rc) g_launches += 1;
    return rc;
}
int b200_quotient_eval(const b200_fr* const* columns, size_t n_columns, uint32_t k, uint32_t ext_k, const b200_col_ref* loads, size_t n_loads,
                       const b200_fr* constants, size_t n_constants, const b200_instr* program, size_t n_instr, b200_fr* out) {
    Ctx* c; if (int rc = get_ctx(&c)) return rc;
    B200_CHECK(out && (n_columns == 0 || columns), -1, "quotient_eval: null pointer");
    B200_CHECK(ext_k >= 1 && ext_k <= 28, -1, "quotient_eval: ext_k out of range");
    const size_t N = (size_t)1 << ext_k;
    if (c->stage_a.ensure(sizeof(Fr) * N * (n_columns ? n_columns : 1)) || c->stage_b.ensure(sizeof(Fr) * N)) return -2;
    std::vector<const void*> ptrs(n_columns);
    for (size_t i = 0; i < n_columns; ++i) {
        B200_CHECK(columns[i], -1, "quotient_eval: column %zu is null", i);
        ptrs[i] = c->stage_a.as<Fr>() + i * N;
        B200_CUDA(cudaMemcpyAsync(c->stage_a.as<Fr>() + i * N, columns[i], sizeof(Fr) * N, cudaMemcpyHostToDevice, c->stream));
    }
    if (int rc = b200_quotient_eval_dev(ptrs.data(), n_columns, k, ext_k, loads, n_loads, constants, n_constants, program, n_instr, c->stage_b.p, nullptr)) return rc;
    B200_CUDA(cudaMemcpyAsync(out, c->stage_b.p, sizeof(Fr) * N, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
