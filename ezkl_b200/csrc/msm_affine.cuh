// msm_affine.cuh — batched-affine bucket accumulation (alternative to k_accumulate's XYZZ chain), host + device bodies.
//
// A chunk of L points of one bucket is summed as a binary tree: round r pairs up neighbours (2j, 2j+1) of the chunk's current
// point list and writes the sums back, halving the list.  Affine addition needs one inversion per pair; all pairs of a round,
// across every chunk of every column of the batch, share ONE level of Fermat inversions through Montgomery's trick applied
// hierarchically: per thread (its chunk's pairs) -> groups of 32 threads -> groups of 32 groups -> a few thousand values that
// are inverted in parallel.  Cost per pair: 1 multiply to accumulate the denominator product (phase A), 2 to peel its inverse
// and 3 for lambda / x3 / y3 (phase C) = 6 + ~0.2 amortised, against 10.4 for the XYZZ mixed addition.
// Phase A parks the running prefix product of pair j in the (still unused) output slot of pair j, so no extra scratch is needed.
// Every per-thread body is a plain __host__ __device__ function of the global thread index, so the whole pipeline is also
// runnable on the CPU (tests/test_host_logic.py) — the kernels in msm_affine.cu are loops-free wrappers around them.
#pragma once
#include "ec.cuh"

namespace b200 {

struct AffineArgs {
    const G1Affine* table;          // precomputed window table (round 0 input), indexed by entry & 0x7fffffff
    const uint32_t* ents;           // [col][ent_stride]
    const uint32_t* chunk_start;    // [col][chunk_stride]
    const uint32_t* chunk_len;
    const uint32_t* order;
    const uint32_t* chunk_offs;     // [col][nbuckets + 1]; last element = number of chunks of the column
    G1Affine* pb_in;                // [col][ent_stride]  point list of the previous round (unused in round 0)
    G1Affine* pb_out;               // [col][ent_stride]  point list this round writes (slot start + j)
    Fq* thread_prod;                // [total_threads]    product of a thread's denominators (1 if idle)
    Fq* thread_inv;                 // [total_threads]    its inverse (filled by the down sweep)
    uint64_t ent_stride, chunk_stride;
    uint32_t nbuckets, round, batch;
};

// current list length of a chunk of `len` points before round r
HD uint32_t aff_len_at(uint32_t len, uint32_t r) {
    for (uint32_t i = 0; i < r; ++i) len = (len + 1u) >> 1;
    return len;
}

HD G1Affine aff_table_point(const AffineArgs& a, uint64_t col, uint32_t idx) {
    const uint32_t e = a.ents[col * a.ent_stride + idx];
    G1Affine p = a.table[e & 0x7fffffffu];
    if (e >> 31) p = g1_neg(p);
    return p;
}
HD G1Affine aff_input(const AffineArgs& a, uint64_t col, uint32_t start, uint32_t idx) {
    return a.round == 0 ? aff_table_point(a, col, start + idx) : a.pb_in[col * a.ent_stride + start + idx];
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

// thread -> (col, chunk); returns false when the thread has no pair to add in this round
HD bool aff_thread_chunk(const AffineArgs& a, uint64_t g, uint64_t* col, uint32_t* start, uint32_t* len_r) {
    *col = g / a.chunk_stride;
    const uint32_t t = (uint32_t)(g % a.chunk_stride);
    if (*col >= a.batch) return false;
    const uint32_t nchunks = a.chunk_offs[*col * (a.nbuckets + 1) + a.nbuckets];
    if (t >= nchunks) return false;
    const uint32_t ch = a.order[*col * a.chunk_stride + t];
    *start = a.chunk_start[*col * a.chunk_stride + ch];
    *len_r = aff_len_at(a.chunk_len[*col * a.chunk_stride + ch], a.round);
    return true;
}

// x coordinate (and only it) of an input point: half the bytes of the point, which is all phase A needs in the general case
HD Fq aff_input_x(const AffineArgs& a, uint64_t col, uint32_t start, uint32_t idx) {
    if (a.round == 0) return a.table[a.ents[col * a.ent_stride + start + idx] & 0x7fffffffu].x;
    return a.pb_in[col * a.ent_stride + start + idx].x;
}
// Phase A: denominators of the thread's pairs; prefix product of pair j parked in pb_out[start + j].x.
// Only the x coordinates are read unless they coincide (identity operand, doubling or inverse pair), which needs the y's too.
HD void aff_phase_a(const AffineArgs& a, uint64_t g) {
    uint64_t col; uint32_t start, L;
    Fq acc = fp_one<FqTag>();
    if (aff_thread_chunk(a, g, &col, &start, &L) && L >= 2) {
        const uint32_t K = L >> 1;
        for (uint32_t j = 0; j < K; ++j) {
            const Fq px = aff_input_x(a, col, start, 2 * j), qx = aff_input_x(a, col, start, 2 * j + 1);
            a.pb_out[col * a.ent_stride + start + j].x = acc;
            Fq d = qx - px;
            if (fp_is_zero(d) || fp_is_zero(px) || fp_is_zero(qx))       // rare: fall back to the complete rule (needs y)
                d = aff_denominator(aff_input(a, col, start, 2 * j), aff_input(a, col, start, 2 * j + 1));
            acc = acc * d;
        }
    }
    a.thread_prod[g] = acc;
}
// Phase C: peel the inverses off from the last pair to the first, finish the additions, carry an odd leftover over
HD void aff_phase_c(const AffineArgs& a, uint64_t g) {
    uint64_t col; uint32_t start, L;
    if (!aff_thread_chunk(a, g, &col, &start, &L) || L < 2) {
        // chunks that are already down to one point still have to appear in this round's output list
        if (aff_thread_chunk(a, g, &col, &start, &L) && L == 1) a.pb_out[col * a.ent_stride + start] = aff_input(a, col, start, 0);
        return;
    }
    const uint32_t K = L >> 1;
    G1Affine* out = a.pb_out + col * a.ent_stride + start;
    G1Affine leftover;
    const bool odd = (L & 1u) != 0;
    if (odd) leftover = aff_input(a, col, start, L - 1);
    Fq run = a.thread_inv[g];
    for (uint32_t j = K; j-- > 0;) {
        const G1Affine p = aff_input(a, col, start, 2 * j), q = aff_input(a, col, start, 2 * j + 1);
        const Fq inv = run * out[j].x;               // prefix product of pairs < j was parked here by phase A
        run = run * aff_denominator(p, q);
        out[j] = aff_add_with_inv(p, q, inv);
    }
    if (odd) out[K] = leftover;
}

// Hierarchy: products of groups of AFF_GROUP consecutive values with exclusive prefix products kept for the way down.
static constexpr uint32_t AFF_GROUP = 32;
HD void aff_up(const Fq* vals, uint64_t n_vals, Fq* prefix, Fq* group_prod, uint64_t u) {
    const uint64_t lo = u * AFF_GROUP, hi = lo + AFF_GROUP < n_vals ? lo + AFF_GROUP : n_vals;
    Fq acc = fp_one<FqTag>();
    for (uint64_t i = lo; i < hi; ++i) { prefix[i] = acc; acc = acc * vals[i]; }
    group_prod[u] = acc;
}
HD void aff_down(const Fq* vals, uint64_t n_vals, const Fq* prefix, const Fq* group_inv, Fq* inv_out, uint64_t u) {
    const uint64_t lo = u * AFF_GROUP, hi = lo + AFF_GROUP < n_vals ? lo + AFF_GROUP : n_vals;
    Fq run = group_inv[u];
    for (uint64_t i = hi; i-- > lo;) { const Fq v = vals[i]; inv_out[i] = run * prefix[i]; run = run * v; }
}

}  // namespace b200
