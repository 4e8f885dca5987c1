// msm_affine.cu — kernels and orchestration of the batched-affine bucket accumulation (see msm_affine.cuh).
// Selected with B200_MSM_AFFINE=1 (msm.cu); the XYZZ chain k_accumulate stays the default until this path is measured faster.
#include <vector>
#include "msm.cuh"
#include "msm_affine.cuh"

namespace b200 {

__global__ void __launch_bounds__(128) k_aff_phase_a(const AffineArgs a, uint64_t total) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < total) aff_phase_a(a, g);
}
__global__ void __launch_bounds__(128) k_aff_phase_c(const AffineArgs a, uint64_t total) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < total) aff_phase_c(a, g);
}
__global__ void __launch_bounds__(128) k_aff_up(const Fq* vals, uint64_t n_vals, Fq* prefix, Fq* group_prod, uint64_t n_groups) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_groups) aff_up(vals, n_vals, prefix, group_prod, u);
}
__global__ void __launch_bounds__(128) k_aff_down(const Fq* vals, uint64_t n_vals, const Fq* prefix, const Fq* group_inv, Fq* inv_out, uint64_t n_groups) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_groups) aff_down(vals, n_vals, prefix, group_inv, inv_out, u);
}
__global__ void __launch_bounds__(128) k_aff_invert(Fq* vals, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vals[i] = fp_inv(vals[i]);
}
// chunk sums for the downstream combine / reduce kernels: XYZZ view of each chunk's final affine point
__global__ void __launch_bounds__(128) k_aff_finish(const AffineArgs a, const G1Affine* __restrict__ pb_final, G1Xyzz* __restrict__ chunk_sums, uint64_t total) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const uint64_t col = g / a.chunk_stride;
    const uint32_t t = (uint32_t)(g % a.chunk_stride);
    const uint32_t nchunks = a.chunk_offs[col * (a.nbuckets + 1) + a.nbuckets];
    if (t >= nchunks) return;
    const uint32_t ch = a.order[col * a.chunk_stride + t];
    const uint32_t start = a.chunk_start[col * a.chunk_stride + ch];
    chunk_sums[col * a.chunk_stride + ch] = g1_to_xyzz(pb_final[col * a.ent_stride + start]);
}

size_t msm_affine_workspace_bytes(size_t batch, size_t ent_stride, size_t chunk_stride) {
    const size_t tt = batch * chunk_stride, n1 = (tt + AFF_GROUP - 1) / AFF_GROUP, n2 = (n1 + AFF_GROUP - 1) / AFF_GROUP;
    return 2 * batch * ent_stride * sizeof(G1Affine) + (2 * tt + 2 * n1 + n2 + 8) * sizeof(Fq);
}

int msm_accumulate_affine(const MsmTable& t, const uint32_t* ents, size_t ent_stride, const uint32_t* chunk_start, const uint32_t* chunk_len,
                          const uint32_t* order, size_t chunk_stride, const uint32_t* chunk_offs, uint32_t nbuckets, uint32_t cap, int batch,
                          G1Xyzz* chunk_sums, DevBuf& scratch, cudaStream_t st) {
    if (scratch.ensure(msm_affine_workspace_bytes(batch, ent_stride, chunk_stride))) return -2;
    const uint64_t tt = (uint64_t)batch * chunk_stride, n1 = (tt + AFF_GROUP - 1) / AFF_GROUP, n2 = (n1 + AFF_GROUP - 1) / AFF_GROUP;
    G1Affine* pb[2];
    pb[0] = scratch.as<G1Affine>();
    pb[1] = pb[0] + (size_t)batch * ent_stride;
    Fq* thread_prod = reinterpret_cast<Fq*>(pb[1] + (size_t)batch * ent_stride);
    Fq* prefix0 = thread_prod + tt;
    Fq* g1 = prefix0 + tt;
    Fq* prefix1 = g1 + n1;
    Fq* g2 = prefix1 + n1;
    AffineArgs a;
    a.table = t.d_table; a.ents = ents; a.chunk_start = chunk_start; a.chunk_len = chunk_len; a.order = order; a.chunk_offs = chunk_offs;
    a.thread_prod = thread_prod; a.thread_inv = thread_prod;      // inverted in place by the down sweep
    a.ent_stride = ent_stride; a.chunk_stride = chunk_stride; a.nbuckets = nbuckets; a.batch = (uint32_t)batch;
    uint32_t rounds = 0;
    while ((1u << rounds) < cap) ++rounds;
    if (rounds == 0) rounds = 1;                                   // cap == 1: a single copy round
    for (uint32_t r = 0; r < rounds; ++r) {
        a.round = r; a.pb_in = pb[(r + 1) & 1]; a.pb_out = pb[r & 1];
        k_aff_phase_a<<<div_up(tt, 128), 128, 0, st>>>(a, tt);
        k_aff_up<<<div_up(n1, 128), 128, 0, st>>>(thread_prod, tt, prefix0, g1, n1);
        k_aff_up<<<div_up(n2, 128), 128, 0, st>>>(g1, n1, prefix1, g2, n2);
        k_aff_invert<<<div_up(n2, 128), 128, 0, st>>>(g2, n2);
        k_aff_down<<<div_up(n2, 128), 128, 0, st>>>(g1, n1, prefix1, g2, g1, n2);
        k_aff_down<<<div_up(n1, 128), 128, 0, st>>>(thread_prod, tt, prefix0, g1, thread_prod, n1);
        k_aff_phase_c<<<div_up(tt, 128), 128, 0, st>>>(a, tt);
    }
    k_aff_finish<<<div_up(tt, 128), 128, 0, st>>>(a, pb[(rounds - 1) & 1], chunk_sums, tt);
    B200_CUDA(cudaGetLastError());
    return 0;
}

// ---- CPU run of the same bodies on host arrays (single column, identity order): tests/test_host_logic.py -------------
int msm_affine_host_chunks(const G1Affine* table, const uint32_t* ents, size_t n_ents, const uint32_t* chunk_start, const uint32_t* chunk_len,
                           size_t nchunks, G1Affine* out) {
    std::vector<uint32_t> order(nchunks), offs(1, (uint32_t)nchunks);
    uint32_t cap = 1;
    for (size_t i = 0; i < nchunks; ++i) { order[i] = (uint32_t)i; if (chunk_len[i] > cap) cap = chunk_len[i]; }
    std::vector<G1Affine> pba(n_ents), pbb(n_ents);
    const uint64_t tt = nchunks, n1 = (tt + AFF_GROUP - 1) / AFF_GROUP, n2 = (n1 + AFF_GROUP - 1) / AFF_GROUP;
    std::vector<Fq> tp(tt), p0(tt), g1(n1), p1(n1), g2(n2);
    AffineArgs a;
    a.table = table; a.ents = ents; a.chunk_start = chunk_start; a.chunk_len = chunk_len; a.order = order.data(); a.chunk_offs = offs.data();
    a.thread_prod = tp.data(); a.thread_inv = tp.data(); a.ent_stride = n_ents; a.chunk_stride = nchunks; a.nbuckets = 0; a.batch = 1;
    uint32_t rounds = 0;
    while ((1u << rounds) < cap) ++rounds;
    if (rounds == 0) rounds = 1;
    G1Affine* pb[2] = {pba.data(), pbb.data()};
    for (uint32_t r = 0; r < rounds; ++r) {
        a.round = r; a.pb_in = pb[(r + 1) & 1]; a.pb_out = pb[r & 1];
        for (uint64_t g = 0; g < tt; ++g) aff_phase_a(a, g);
        for (uint64_t u = 0; u < n1; ++u) aff_up(tp.data(), tt, p0.data(), g1.data(), u);
        for (uint64_t u = 0; u < n2; ++u) aff_up(g1.data(), n1, p1.data(), g2.data(), u);
        for (uint64_t i = 0; i < n2; ++i) g2[i] = fp_inv(g2[i]);
        for (uint64_t u = 0; u < n2; ++u) aff_down(g1.data(), n1, p1.data(), g2.data(), g1.data(), u);
        for (uint64_t u = 0; u < n1; ++u) aff_down(tp.data(), tt, p0.data(), g1.data(), tp.data(), u);
        for (uint64_t g = 0; g < tt; ++g) aff_phase_c(a, g);
    }
    for (size_t i = 0; i < nchunks; ++i) out[i] = pb[(rounds - 1) & 1][chunk_start[i]];
    return 0;
}

}  // namespace b200
