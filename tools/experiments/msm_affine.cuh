// msm_affine.cuh — EXPERIMENT (not in the product library): batched-affine bucket accumulation, an alternative to k_accumulate's
// XYZZ chain; host + device bodies.
//
// A chunk of L points of one bucket is summed as a binary tree: round r pairs up neighbours (2j, 2j+1) of the chunk's current
// point list and writes the sums back, halving the list.  An affine addition costs one inversion; all pairs of a round, across
// every chunk of every column of the batch, share ONE level of inversions through Montgomery's trick:
//   * a thread owns AFF_CPT chunks and treats all their pairs as one sequence: the producer side multiplies the pair denominators
//     into a running product and parks the product-so-far of every pair; the consumer side starts from the inverse of the
//     thread's total and peels the pairs in REVERSE order (inv_j = run * parked_j; run *= d_j);
//   * the threads of a warp combine their totals with two shuffle scans (each lane gets the product of the OTHER 31 lanes and the
//     warp total), so only one value per warp — a few thousand per round — needs a real (Fermat) inversion, done by one small
//     kernel between rounds;
//   * FUSION: the consumer of round r is also the producer of round r + 1 — as soon as two neighbouring sums exist it forms
//     their denominator — so every point crosses HBM once per round (the round-1 version read the lists twice per round and
//     launched seven kernels per round; this one launches two).  Because peeling reverses the order of accumulation, the
//     direction alternates: the first producer pass runs ascending, round 0 descending, round 1 ascending, ...
// Parked products live in the unused tail of the chunk's region of the list buffers, so the scratch is the two ping-pong point
// lists plus 32 B per thread.  Cost per addition: 1 (parked product) + 2 (peel) + 3 (lambda, x3, y3; one is a squaring) = 5.8
// multiply-equivalents + 12 per thread and round for the warp scans, against 9.46 for the XYZZ mixed addition.
// Every per-thread body is a plain __host__ __device__ function, so the pipeline also runs on the CPU (tests/test_host_logic.py).
#pragma once
#include "../../ezkl_b200/csrc/ec.cuh"

namespace b200 {

struct MsmTable;
struct DevBuf;
size_t msm_affine_workspace_bytes(size_t batch, size_t ent_stride, size_t chunk_stride);
int msm_accumulate_affine(const MsmTable& t, const uint32_t* ents, size_t ent_stride, const uint32_t* chunk_start, const uint32_t* chunk_len,
                          const uint32_t* order, size_t chunk_stride, const uint32_t* chunk_offs, uint32_t nbuckets, uint32_t cap, int batch,
                          G1Xyzz* chunk_sums, DevBuf& scratch, cudaStream_t st);
int msm_affine_launches(uint32_t cap);
int msm_affine_host_chunks(const G1Affine* table, const uint32_t* ents, size_t n_ents, const uint32_t* chunk_start, const uint32_t* chunk_len,
                           size_t nchunks, G1Affine* out);

static constexpr uint32_t AFF_CPT = 8;          // chunks per thread

struct AffineArgs {
    const G1Affine* table;          // precomputed window table (round 0 input), indexed by entry & 0x7fffffff
    const uint32_t* ents;           // [col][ent_stride]
    const uint32_t* chunk_start;    // [col][chunk_stride]
    const uint32_t* chunk_len;
    const uint32_t* order;
    const uint32_t* chunk_offs;     // [col][nbuckets + 1]; last element = number of chunks of the column
    G1Affine* pb[2];                // ping-pong point lists [col][ent_stride]; round r reads pb[(r + 1) & 1] (r > 0), writes pb[r & 1]
    Fq* thr_aux;                    // [threads] product of the other lanes' totals in the thread's warp (producer -> next consumer)
    Fq* warp_prod;                  // [threads / 32] warp totals; inverted in place between the producer and the consumer
    uint64_t ent_stride, chunk_stride;
    uint32_t nbuckets, batch, threads_per_col, round, last_round;
};

// list length of a chunk of `len` points before round r
HD uint32_t aff_len_at(uint32_t len, uint32_t r) {
    for (uint32_t i = 0; i < r; ++i) len = (len + 1u) >> 1;
    return len;
}
// number of rounds a chunk of `len` points takes
HD uint32_t aff_rounds_of(uint32_t len) {
    uint32_t r = 0;
    while (len > 1) { len = (len + 1u) >> 1; ++r; }
    return r;
}

HD G1Affine aff_table_point(const AffineArgs& a, uint64_t col, uint32_t idx) {
    const uint32_t e = a.ents[col * a.ent_stride + idx];
    G1Affine p = a.table[e & 0x7fffffffu];
    if (e >> 31) p = g1_neg(p);
    return p;
}
HD G1Affine aff_input(const AffineArgs& a, uint64_t col, uint32_t start, uint32_t idx) {
    return a.round == 0 ? aff_table_point(a, col, start + idx) : a.pb[(a.round + 1) & 1][col * a.ent_stride + start + idx];
}
// parked products of a chunk: the consumer of round r reads them behind the round's input list (round 0: at the chunk's start of
// the otherwise unused buffer), the producer for round r + 1 writes them behind the round's output list
HD Fq* aff_park_in(const AffineArgs& a, uint64_t col, uint32_t start, uint32_t len_r) {
    return reinterpret_cast<Fq*>(a.pb[(a.round + 1) & 1] + col * a.ent_stride + start + (a.round == 0 ? 0u : len_r));
}
HD Fq* aff_park_out(const AffineArgs& a, uint64_t col, uint32_t start, uint32_t len_next) {
    return reinterpret_cast<Fq*>(a.pb[a.round & 1] + col * a.ent_stride + start + len_next);
}

// Denominator of p + q: x_q - x_p in general, 2*y_p when p == q, and 1 whenever no inversion is needed
// (an identity operand, or p == -q).  Never zero, so products of denominators stay invertible.
HD Fq aff_denominator(const G1Affine& p, const G1Affine& q) {
    if (g1_is_identity(p) || g1_is_identity(q)) return fp_one<FqTag>();
    if (fp_eq(p.x, q.x)) return fp_eq(p.y, q.y) ? fp_dbl(p.y) : fp_one<FqTag>();
    return q.x - p.x;
}
// p + q given inv = 1 / aff_denominator(p, q)
HD G1Affine aff_add_with_inv(const G1Affine& p, const G1Affine& q, const Fq& inv) {
    if (g1_is_identity(p)) return q;
    if (g1_is_identity(q)) return p;
    Fq lambda;
    if (fp_eq(p.x, q.x)) {
        if (!fp_eq(p.y, q.y)) { G1Affine o; o.x = fp_zero<FqTag>(); o.y = fp_zero<FqTag>(); return o; }     // p == -q
        const Fq xx = fp_sqr(p.x);
        lambda = (fp_dbl(xx) + xx) * inv;                                                                       // 3x^2 / 2y
    } else {
        lambda = (q.y - p.y) * inv;
    }
    G1Affine o;
    o.x = fp_sqr(lambda) - p.x - q.x;
    o.y = lambda * (p.x - o.x) - p.y;
    return o;
}

// thread g -> (column, first chunk position); false when the thread lies outside the batch
HD bool aff_thread(const AffineArgs& a, uint64_t g, uint64_t* col, uint32_t* pos0, uint32_t* nchunks) {
    *col = g / a.threads_per_col;
    if (*col >= a.batch) return false;
    *pos0 = (uint32_t)(g % a.threads_per_col) * AFF_CPT;
    *nchunks = a.chunk_offs[*col * (a.nbuckets + 1) + a.nbuckets];
    return *pos0 < *nchunks;
}

// First producer pass (before round 0), ascending: parks the running product of every pair's denominator, returns the thread's total.
// Only the x coordinates are read unless they coincide or vanish (identity operand, doubling, inverse pair).
HD Fq aff_first_pass(const AffineArgs& a, uint64_t g) {
    uint64_t col; uint32_t pos0, nchunks;
    Fq acc = fp_one<FqTag>();
    if (!aff_thread(a, g, &col, &pos0, &nchunks)) return acc;
    for (uint32_t c = 0; c < AFF_CPT && pos0 + c < nchunks; ++c) {
        const uint32_t ch = a.order[col * a.chunk_stride + pos0 + c];
        const uint32_t start = a.chunk_start[col * a.chunk_stride + ch], L = a.chunk_len[col * a.chunk_stride + ch];
        Fq* park = aff_park_in(a, col, start, L);
        for (uint32_t j = 0; j < (L >> 1); ++j) {
            const uint32_t e0 = a.ents[col * a.ent_stride + start + 2 * j], e1 = a.ents[col * a.ent_stride + start + 2 * j + 1];
            const Fq px = a.table[e0 & 0x7fffffffu].x, qx = a.table[e1 & 0x7fffffffu].x;
            Fq d = qx - px;
            if (fp_is_zero(d) || fp_is_zero(px) || fp_is_zero(qx)) d = aff_denominator(aff_table_point(a, col, start + 2 * j), aff_table_point(a, col, start + 2 * j + 1));
            park[j] = acc;
            acc = acc * d;
        }
    }
    return acc;
}

// Round r: consumer of the round's pairs and producer for round r + 1.  `inv_total` = inverse of this thread's denominator total
// for round r; returns the thread's total for round r + 1 (1 on the last round).
HD Fq aff_round(const AffineArgs& a, uint64_t g, Fq inv_total) {
    uint64_t col; uint32_t pos0, nchunks;
    Fq acc2 = fp_one<FqTag>();
    if (!aff_thread(a, g, &col, &pos0, &nchunks)) return acc2;
    const bool desc = (a.round & 1u) == 0;
    const bool produce = !a.last_round;
    Fq run = inv_total;
    uint32_t cnt = nchunks - pos0 < AFF_CPT ? nchunks - pos0 : AFF_CPT;
    for (uint32_t cc = 0; cc < cnt; ++cc) {
        const uint32_t c = desc ? cnt - 1 - cc : cc;
        const uint32_t ch = a.order[col * a.chunk_stride + pos0 + c];
        const uint32_t start = a.chunk_start[col * a.chunk_stride + ch];
        const uint32_t L = aff_len_at(a.chunk_len[col * a.chunk_stride + ch], a.round);
        if (L < 2) continue;                                    // finished chunks rest where they finished
        const uint32_t K = L >> 1, odd = L & 1u, Ln = K + odd;
        const Fq* park = aff_park_in(a, col, start, L);
        Fq* park2 = aff_park_out(a, col, start, Ln);
        G1Affine* out = a.pb[a.round & 1] + col * a.ent_stride + start;
        G1Affine held; held.x = fp_zero<FqTag>(); held.y = fp_zero<FqTag>();
        for (uint32_t ii = 0; ii < Ln; ++ii) {
            const uint32_t i = desc ? Ln - 1 - ii : ii;
            G1Affine e;
            if (odd && i == K) e = aff_input(a, col, start, L - 1);           // the odd leftover passes through
            else {
                const G1Affine p = aff_input(a, col, start, 2 * i), q = aff_input(a, col, start, 2 * i + 1);
                const Fq inv = run * park[i];
                run = run * aff_denominator(p, q);
                e = aff_add_with_inv(p, q, inv);
            }
            out[i] = e;
            if (!produce) continue;
            // next round pairs (2m, 2m+1) of the output list; an odd-length list leaves its last element unpaired
            if ((i & 1u) == (desc ? 1u : 0u)) held = e;                       // first member of its pair in this direction
            else if (!((Ln & 1u) && i == Ln - 1)) {
                const Fq d2 = desc ? aff_denominator(e, held) : aff_denominator(held, e);
                park2[i >> 1] = acc2;
                acc2 = acc2 * d2;
            }
        }
    }
    return acc2;
}
// In the descending direction the unpaired last element (index Ln - 1, even) arrives first and must not be taken for a pair
// member: it is even, so `held` is not touched (desc holds odd indices) and the pairing branch is skipped by the explicit test.
// In the ascending direction it arrives last, is even, and only overwrites `held`.

}  // namespace b200
