/*
 * bn254_oracle.c — CPU restatement of the reference prover's hot-path arithmetic (TEST INFRASTRUCTURE).
 *
 * This file is the parity ORACLE and the "port" CPU baseline.  It is compiled into oracle/liboracle.so and
 * may be loaded ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product (ezkl_b200/, libezkl_b200.so) never links, loads or calls it.
 *
 * What it restates (none of it is under /root/reference — SURVEY.md §0.2, §8c; algorithms as published in):
 *   halo2_proofs 0.3.0 @ zkonduit/halo2#01c88842   src/arithmetic.rs  : best_multiexp, multiexp_serial,
 *        best_fft, recursive_butterfly_arithmetic, eval_polynomial, kate_division, parallelize
 *                                                   src/poly/domain.rs : EvaluationDomain::{new, lagrange_to_coeff,
 *        coeff_to_extended, extended_to_coeff, distribute_powers_zeta, divide_by_vanishing_poly}
 *                                                   src/poly/kzg/commitment.rs : ParamsKZG::{commit, commit_lagrange}
 *   halo2curves 0.7.0 @ privacy-scaling-explorations/halo2curves#b753a832  src/bn256/{fr,fq,curve}.rs
 *   ezkl call sites that fix the semantics: /root/reference/src/pfsys/mod.rs:390,396,456 (create_keys /
 *        create_proof), src/pfsys/srs.rs:15,36,46, src/circuit/modules/polycommit.rs:52,71 (commit_lagrange).
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against the reference's own fixtures
 * (tests/golden/kzg_k6.srs, pk_k6_subset.npz, extracted from /root/reference/tests/assets by
 * tests/golden/make_golden.py): 64 MSM known answers, NTT / inverse-NTT / coset-NTT known answers.
 *
 * Wire format everywhere (= halo2curves SerdeObject raw bytes):
 *   Fr/Fq  : uint64_t[4] little-endian limbs, Montgomery form (R = 2^256)
 *   G1Aff  : {Fq x, Fq y} 64 B, identity = (0,0)
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef struct { u64 l[4]; } fe;
typedef struct { fe x, y; } g1a;       /* affine, identity = (0,0) */
typedef struct { fe x, y, z; } g1j;    /* Jacobian (X/Z^2, Y/Z^3), identity z = 0 */

/* ---- constants ------------------------------------------------------------- */
static const u64 FQ_M[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 FQ_INV = 0x87d20782e4866389ULL;
static const fe FQ_R1 = {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}};
static const fe FQ_R2 = {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};
static const u64 FR_M[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 FR_INV = 0xc2e1f593efffffffULL;
static const fe FR_R1 = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const fe FR_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
/* canonical (non-Montgomery) values */
static const fe FR_ROOT_OF_UNITY_C = {{0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL}};
static const fe FR_ZETA_C = {{0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL}};
#define FR_S 28

/* ---- generic 4-limb Montgomery field ---------------------------------------------------------------- */
static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe *a, const fe *b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int geq(const u64 a[4], const u64 m[4]) {
    for (int i = 3; i >= 0; --i) { if (a[i] > m[i]) return 1; if (a[i] < m[i]) return 0; }
    return 1;
}
static inline void sub_nb(u64 r[4], const u64 a[4], const u64 b[4], u64 *borrow_out) {
    u64 br = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)a[i] - b[i] - br; r[i] = (u64)d; br = (u64)(d >> 64) & 1; }
    if (borrow_out) *borrow_out = br;
}
static inline void f_add(fe *r, const fe *a, const fe *b, const u64 m[4]) {
    u64 c = 0, t[4];
    for (int i = 0; i < 4; ++i) { u128 s = (u128)a->l[i] + b->l[i] + c; t[i] = (u64)s; c = (u64)(s >> 64); }
    if (c || geq(t, m)) sub_nb(t, t, m, 0);
    memcpy(r->l, t, 32);
}
static inline void f_sub(fe *r, const fe *a, const fe *b, const u64 m[4]) {
    u64 br, t[4];
    sub_nb(t, a->l, b->l, &br);
    if (br) { u64 c = 0; for (int i = 0; i < 4; ++i) { u128 s = (u128)t[i] + m[i] + c; t[i] = (u64)s; c = (u64)(s >> 64); } }
    memcpy(r->l, t, 32);
}
static inline void f_neg(fe *r, const fe *a, const u64 m[4]) {
    if (fe_is_zero(a)) { *r = *a; return; }
    sub_nb(r->l, m, a->l, 0);
}
static inline void f_mul(fe *r, const fe *a, const fe *b, const u64 m[4], u64 inv) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u64 carry = 0; u128 acc;
        for (int j = 0; j < 4; ++j) { acc = (u128)a->l[j] * b->l[i] + t[j] + carry; t[j] = (u64)acc; carry = (u64)(acc >> 64); }
        acc = (u128)t[4] + carry; t[4] = (u64)acc; t[5] = (u64)(acc >> 64);
        u64 mm = t[0] * inv;
        acc = (u128)mm * m[0] + t[0]; carry = (u64)(acc >> 64);
        for (int j = 1; j < 4; ++j) { acc = (u128)mm * m[j] + t[j] + carry; t[j - 1] = (u64)acc; carry = (u64)(acc >> 64); }
        acc = (u128)t[4] + carry; t[3] = (u64)acc; t[4] = t[5] + (u64)(acc >> 64);
    }
    if (t[4] || geq(t, m)) sub_nb(t, t, m, 0);
    memcpy(r->l, t, 32);
}
static void f_pow(fe *r, const fe *a, const u64 e[4], const fe *one, const u64 m[4], u64 inv) {
    fe acc = *one;
    for (int i = 255; i >= 0; --i) {
        f_mul(&acc, &acc, &acc, m, inv);
        if ((e[i / 64] >> (i % 64)) & 1) f_mul(&acc, &acc, a, m, inv);
    }
    *r = acc;
}
static void f_inv(fe *r, const fe *a, const fe *one, const u64 m[4], u64 inv) {
    u64 e[4]; u64 two[4] = {2, 0, 0, 0};
    sub_nb(e, m, two, 0);
    f_pow(r, a, e, one, m, inv);   /* 0 -> 0, as halo2curves' invert().unwrap_or(0) users expect */
}

/* Fq / Fr front ends */
#define Q_ADD(r, a, b) f_add(r, a, b, FQ_M)
#define Q_SUB(r, a, b) f_sub(r, a, b, FQ_M)
#define Q_MUL(r, a, b) f_mul(r, a, b, FQ_M, FQ_INV)
#define R_ADD(r, a, b) f_add(r, a, b, FR_M)
#define R_SUB(r, a, b) f_sub(r, a, b, FR_M)
#define R_MUL(r, a, b) f_mul(r, a, b, FR_M, FR_INV)
static inline void fr_from_mont(fe *r, const fe *a) { fe one = {{1, 0, 0, 0}}; R_MUL(r, a, &one); }
static inline void fr_to_mont(fe *r, const fe *a) { R_MUL(r, a, &FR_R2); }
static inline void fr_inv(fe *r, const fe *a) { f_inv(r, a, &FR_R1, FR_M, FR_INV); }
static inline void fq_inv(fe *r, const fe *a) { f_inv(r, a, &FQ_R1, FQ_M, FQ_INV); }
static void fr_pow_u64(fe *r, const fe *a, u64 e) {
    u64 ee[4] = {e, 0, 0, 0};
    f_pow(r, a, ee, &FR_R1, FR_M, FR_INV);
}

/* ---- G1 (y^2 = x^3 + 3), Jacobian coordinates -------------------------------------------------------- */
static inline void j_set_identity(g1j *p) { memset(p, 0, sizeof *p); }
static inline int j_is_identity(const g1j *p) { return fe_is_zero(&p->z); }
static inline int a_is_identity(const g1a *p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void j_from_affine(g1j *r, const g1a *p) {
    if (a_is_identity(p)) { j_set_identity(r); return; }
    r->x = p->x; r->y = p->y; r->z = FQ_R1;
}
static void j_double(g1j *r, const g1j *p) {   /* dbl-2009-l (a = 0) */
    if (j_is_identity(p)) { *r = *p; return; }
    fe a, b, c, d, e, f, t, x3, y3, z3;
    Q_MUL(&a, &p->x, &p->x); Q_MUL(&b, &p->y, &p->y); Q_MUL(&c, &b, &b);
    Q_ADD(&t, &p->x, &b); Q_MUL(&t, &t, &t); Q_SUB(&t, &t, &a); Q_SUB(&t, &t, &c); Q_ADD(&d, &t, &t);
    Q_ADD(&e, &a, &a); Q_ADD(&e, &e, &a); Q_MUL(&f, &e, &e);
    Q_MUL(&z3, &p->y, &p->z); Q_ADD(&z3, &z3, &z3);
    Q_SUB(&x3, &f, &d); Q_SUB(&x3, &x3, &d);
    Q_SUB(&t, &d, &x3); Q_MUL(&y3, &e, &t);
    Q_ADD(&c, &c, &c); Q_ADD(&c, &c, &c); Q_ADD(&c, &c, &c); Q_SUB(&y3, &y3, &c);
    r->x = x3; r->y = y3; r->z = z3;
}
static void j_add(g1j *r, const g1j *p, const g1j *q) {   /* add-2007-bl with complete case handling */
    if (j_is_identity(p)) { *r = *q; return; }
    if (j_is_identity(q)) { *r = *p; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    Q_MUL(&z1z1, &p->z, &p->z); Q_MUL(&z2z2, &q->z, &q->z);
    Q_MUL(&u1, &p->x, &z2z2); Q_MUL(&u2, &q->x, &z1z1);
    Q_MUL(&s1, &p->y, &q->z); Q_MUL(&s1, &s1, &z2z2);
    Q_MUL(&s2, &q->y, &p->z); Q_MUL(&s2, &s2, &z1z1);
    if (fe_eq(&u1, &u2)) {
        if (fe_eq(&s1, &s2)) { j_double(r, p); return; }
        j_set_identity(r); return;
    }
    Q_SUB(&h, &u2, &u1); Q_ADD(&i, &h, &h); Q_MUL(&i, &i, &i); Q_MUL(&j, &h, &i);
    Q_SUB(&rr, &s2, &s1); Q_ADD(&rr, &rr, &rr); Q_MUL(&v, &u1, &i);
    Q_MUL(&x3, &rr, &rr); Q_SUB(&x3, &x3, &j); Q_SUB(&x3, &x3, &v); Q_SUB(&x3, &x3, &v);
    Q_SUB(&t, &v, &x3); Q_MUL(&y3, &rr, &t); Q_MUL(&t, &s1, &j); Q_ADD(&t, &t, &t); Q_SUB(&y3, &y3, &t);
    Q_ADD(&z3, &p->z, &q->z); Q_MUL(&z3, &z3, &z3); Q_SUB(&z3, &z3, &z1z1); Q_SUB(&z3, &z3, &z2z2); Q_MUL(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void j_add_mixed(g1j *r, const g1j *p, const g1a *q) {   /* madd-2007-bl */
    if (a_is_identity(q)) { *r = *p; return; }
    if (j_is_identity(p)) { j_from_affine(r, q); return; }
    fe z1z1, u2, s2, h, hh, i, j, rr, v, t, x3, y3, z3;
    Q_MUL(&z1z1, &p->z, &p->z); Q_MUL(&u2, &q->x, &z1z1);
    Q_MUL(&s2, &q->y, &p->z); Q_MUL(&s2, &s2, &z1z1);
    if (fe_eq(&p->x, &u2)) {
        if (fe_eq(&p->y, &s2)) { j_double(r, p); return; }
        j_set_identity(r); return;
    }
    Q_SUB(&h, &u2, &p->x); Q_MUL(&hh, &h, &h); Q_ADD(&i, &hh, &hh); Q_ADD(&i, &i, &i); Q_MUL(&j, &h, &i);
    Q_SUB(&rr, &s2, &p->y); Q_ADD(&rr, &rr, &rr); Q_MUL(&v, &p->x, &i);
    Q_MUL(&x3, &rr, &rr); Q_SUB(&x3, &x3, &j); Q_SUB(&x3, &x3, &v); Q_SUB(&x3, &x3, &v);
    Q_SUB(&t, &v, &x3); Q_MUL(&y3, &rr, &t); Q_MUL(&t, &p->y, &j); Q_ADD(&t, &t, &t); Q_SUB(&y3, &y3, &t);
    Q_ADD(&z3, &p->z, &h); Q_MUL(&z3, &z3, &z3); Q_SUB(&z3, &z3, &z1z1); Q_SUB(&z3, &z3, &hh);
    r->x = x3; r->y = y3; r->z = z3;
}
static void j_to_affine(g1a *r, const g1j *p) {
    if (j_is_identity(p)) { memset(r, 0, sizeof *r); return; }
    fe zi, zi2, zi3;
    fq_inv(&zi, &p->z); Q_MUL(&zi2, &zi, &zi); Q_MUL(&zi3, &zi2, &zi);
    Q_MUL(&r->x, &p->x, &zi2); Q_MUL(&r->y, &p->y, &zi3);
}
static void a_neg(g1a *r, const g1a *p) { r->x = p->x; f_neg(&r->y, &p->y, FQ_M); }

/* ---- thread helper ("parallelize" / multicore::scope restated): a persistent pool, like Rayon's, so that an op does not pay
 *      for thread creation — run_threads(fn, arg, T) runs fn(arg, tid, T) for tid < T, tid 0 on the caller ---------------- */
typedef void (*job_fn)(void *arg, int tid, int nthreads);
#define POOL_MAX 512
static pthread_mutex_t pool_call_mu = PTHREAD_MUTEX_INITIALIZER;       /* one parallel region at a time */
static pthread_mutex_t pool_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t pool_go = PTHREAD_COND_INITIALIZER, pool_done = PTHREAD_COND_INITIALIZER;
static pthread_t pool_th[POOL_MAX];
static int pool_size = 0, pool_pending = 0, pool_nthreads = 0;
static unsigned long pool_gen = 0;
static job_fn pool_fn; static void *pool_arg;
static void *pool_main(void *p) {
    const int id = (int)(long)p;                 /* worker id: runs tid = id + 1 */
    unsigned long seen = 0;
    pthread_mutex_lock(&pool_mu);
    for (;;) {
        while (pool_gen == seen) pthread_cond_wait(&pool_go, &pool_mu);
        seen = pool_gen;
        if (id + 1 < pool_nthreads) {
            job_fn fn = pool_fn; void *arg = pool_arg; int nt = pool_nthreads;
            pthread_mutex_unlock(&pool_mu);
            fn(arg, id + 1, nt);
            pthread_mutex_lock(&pool_mu);
            if (--pool_pending == 0) pthread_cond_signal(&pool_done);
        }
    }
    return 0;
}
static void run_threads(job_fn fn, void *arg, int nthreads) {
    if (nthreads <= 1) { fn(arg, 0, 1); return; }
    if (nthreads > POOL_MAX) nthreads = POOL_MAX;
    pthread_mutex_lock(&pool_call_mu);
    pthread_mutex_lock(&pool_mu);
    while (pool_size < nthreads - 1) { pthread_create(&pool_th[pool_size], 0, pool_main, (void *)(long)pool_size); pool_size++; }
    pool_fn = fn; pool_arg = arg; pool_nthreads = nthreads; pool_pending = nthreads - 1; pool_gen++;
    pthread_cond_broadcast(&pool_go);
    pthread_mutex_unlock(&pool_mu);
    fn(arg, 0, nthreads);
    pthread_mutex_lock(&pool_mu);
    while (pool_pending > 0) pthread_cond_wait(&pool_done, &pool_mu);
    pool_nthreads = 0;
    pthread_mutex_unlock(&pool_mu);
    pthread_mutex_unlock(&pool_call_mu);
}

/* ---- MSM: halo2 arithmetic.rs multiexp_serial / best_multiexp ----------------------------------------- */
static inline size_t get_at(size_t segment, size_t c, const uint8_t bytes[32]) {
    size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0};
    size_t avail = 32 - skip_bytes; if (avail > 8) avail = 8;
    memcpy(v, bytes + skip_bytes, avail);
    u64 tmp; memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    tmp %= ((u64)1 << c);
    return (size_t)tmp;
}
typedef struct { uint8_t kind; g1j p; } bucket_t;   /* kind 0 None, 1 Affine (x,y in p), 2 Projective */
static void multiexp_serial(const fe *coeffs, const g1a *bases, size_t n, g1j *acc) {
    fe *repr = (fe *)malloc(sizeof(fe) * (n ? n : 1));
    for (size_t i = 0; i < n; ++i) fr_from_mont(&repr[i], &coeffs[i]);      /* to_repr() */
    size_t c;
    if (n < 4) c = 1; else if (n < 32) c = 3; else c = (size_t)ceil(log((double)n));
    size_t segments = 256 / c + 1, nb = ((size_t)1 << c) - 1;
    bucket_t *buckets = (bucket_t *)malloc(sizeof(bucket_t) * nb);
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t i = 0; i < c; ++i) j_double(acc, acc);
        for (size_t i = 0; i < nb; ++i) buckets[i].kind = 0;
        for (size_t i = 0; i < n; ++i) {
            size_t d = get_at(seg, c, (const uint8_t *)&repr[i]);
            if (!d) continue;
            bucket_t *b = &buckets[d - 1];
            if (b->kind == 0) { b->kind = 1; b->p.x = bases[i].x; b->p.y = bases[i].y; }
            else if (b->kind == 1) {
                g1a a0 = {b->p.x, b->p.y}; g1j t; j_from_affine(&t, &a0);
                j_add_mixed(&b->p, &t, &bases[i]); b->kind = 2;
            } else j_add_mixed(&b->p, &b->p, &bases[i]);
        }
        g1j running; j_set_identity(&running);
        for (size_t i = nb; i-- > 0;) {
            bucket_t *b = &buckets[i];
            if (b->kind == 1) { g1a a0 = {b->p.x, b->p.y}; j_add_mixed(&running, &running, &a0); }
            else if (b->kind == 2) j_add(&running, &running, &b->p);
            j_add(acc, acc, &running);
        }
    }
    free(buckets); free(repr);
}
typedef struct { const fe *coeffs; const g1a *bases; size_t n, chunk, nchunks; g1j *results; } msm_job;
static void msm_worker(void *arg, int tid, int nthreads) {
    msm_job *j = (msm_job *)arg; (void)nthreads;
    if ((size_t)tid >= j->nchunks) return;
    size_t lo = (size_t)tid * j->chunk, hi = lo + j->chunk; if (hi > j->n) hi = j->n;
    j_set_identity(&j->results[tid]);
    multiexp_serial(j->coeffs + lo, j->bases + lo, hi - lo, &j->results[tid]);
}
static void best_multiexp(const fe *coeffs, const g1a *bases, size_t n, int threads, g1j *out) {
    j_set_identity(out);
    if (threads < 1) threads = 1;
    if (n > (size_t)threads) {
        size_t chunk = n / threads, nchunks = (n + chunk - 1) / chunk;
        g1j *results = (g1j *)malloc(sizeof(g1j) * nchunks);
        msm_job j = {coeffs, bases, n, chunk, nchunks, results};
        run_threads(msm_worker, &j, (int)nchunks);
        for (size_t i = 0; i < nchunks; ++i) j_add(out, out, &results[i]);
        free(results);
    } else multiexp_serial(coeffs, bases, n, out);
}

/* ---- NTT: halo2 arithmetic.rs best_fft / recursive_butterfly_arithmetic -------------------------------- */
static inline uint32_t bitreverse32(uint32_t n, uint32_t l) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < l; ++i) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
static void butterfly_level(fe *a, size_t n, size_t twiddle_chunk, const fe *tw) {
    fe *left = a, *right = a + n / 2; fe t;
    t = right[0]; right[0] = left[0]; R_ADD(&left[0], &left[0], &t); R_SUB(&right[0], &right[0], &t);
    for (size_t i = 1; i < n / 2; ++i) {
        R_MUL(&t, &right[i], &tw[i * twiddle_chunk]);
        right[i] = left[i]; R_ADD(&left[i], &left[i], &t); R_SUB(&right[i], &right[i], &t);
    }
}
static void recursive_butterfly(fe *a, size_t n, size_t twiddle_chunk, const fe *tw) {
    if (n == 2) { fe t = a[1]; a[1] = a[0]; R_ADD(&a[0], &a[0], &t); R_SUB(&a[1], &a[1], &t); return; }
    recursive_butterfly(a, n / 2, twiddle_chunk * 2, tw);
    recursive_butterfly(a + n / 2, n / 2, twiddle_chunk * 2, tw);
    butterfly_level(a, n, twiddle_chunk, tw);
}
static int log2_floor(int x) { int l = 0; while ((1 << (l + 1)) <= x) ++l; return l; }
/* best_fft's three parallel regions: the bit-reversal swap and the twiddle table (both `parallelize` loops upstream) and the
 * recursion, whose multicore::join tree is unrolled here into its levels: 2^L independent sub-transforms, then for each of the
 * L levels above them one serial butterfly sweep per tree node, the nodes of a level in parallel — the same work per thread as
 * join gives (the root sweep of n/2 butterflies is serial upstream too). */
typedef struct { fe *a; const fe *omega; fe *tw; uint32_t log_n; int phase; int depth; } fft_job;
static void fft_worker(void *arg, int tid, int nt) {
    fft_job *j = (fft_job *)arg;
    const size_t n = (size_t)1 << j->log_n;
    if (j->phase == 0) {                 /* bit reversal: each k < rk pair is swapped by the thread owning k */
        size_t chunk = (n + nt - 1) / nt, lo = chunk * tid, hi = lo + chunk; if (hi > n) hi = n;
        for (size_t k = lo; k < hi; ++k) { size_t rk = bitreverse32((uint32_t)k, j->log_n); if (k < rk) { fe t = j->a[rk]; j->a[rk] = j->a[k]; j->a[k] = t; } }
    } else if (j->phase == 1) {          /* twiddles omega^i, i < n/2: each thread starts from omega^lo */
        size_t half = n / 2, chunk = (half + nt - 1) / nt, lo = chunk * tid, hi = lo + chunk; if (hi > half) hi = half;
        if (lo >= hi) return;
        fe w; fr_pow_u64(&w, j->omega, (u64)lo);
        for (size_t i = lo; i < hi; ++i) { j->tw[i] = w; R_MUL(&w, &w, j->omega); }
    } else if (j->phase == 2) {          /* the 2^depth sub-transforms below the join levels */
        size_t parts = (size_t)1 << j->depth, sub = n >> j->depth;
        for (size_t p = tid; p < parts; p += nt) recursive_butterfly(j->a + p * sub, sub, parts, j->tw);
    } else {                             /* one level of the join tree: 2^depth nodes of size n >> depth */
        size_t parts = (size_t)1 << j->depth, sub = n >> j->depth;
        for (size_t p = tid; p < parts; p += nt) butterfly_level(j->a + p * sub, sub, parts, j->tw);
    }
}
static void best_fft(fe *a, const fe *omega, uint32_t log_n, int threads) {
    size_t n = (size_t)1 << log_n;
    if (threads < 1) threads = 1;
    int log_threads = log2_floor(threads);
    if (n < 2) return;
    fe *tw = (fe *)malloc(sizeof(fe) * (n / 2));
    fft_job j = {a, omega, tw, log_n, 0, 0};
    const int par = n >= 4096 ? threads : 1;
    run_threads(fft_worker, &j, par);
    j.phase = 1; run_threads(fft_worker, &j, par);
    if ((int)log_n <= log_threads || par == 1) {
        if (par == 1 && (int)log_n > log_threads) recursive_butterfly(a, n, 1, tw);
        else {
            size_t chunk = 2, twiddle_chunk = n / 2;
            for (uint32_t s = 0; s < log_n; ++s) {
                for (size_t off = 0; off < n; off += chunk) butterfly_level(a + off, chunk, twiddle_chunk, tw);
                chunk *= 2; twiddle_chunk /= 2;
            }
        }
    } else {
        j.phase = 2; j.depth = log_threads; run_threads(fft_worker, &j, 1 << log_threads);
        for (int d = log_threads - 1; d >= 0; --d) { j.phase = 3; j.depth = d; run_threads(fft_worker, &j, 1 << d); }
    }
    free(tw);
}

/* ---- parallelize()-style elementwise helpers ----------------------------------------------------------- */
typedef struct { int op; fe *r; const fe *a; const fe *b; const fe *s; size_t n; } ew_job;
static void ew_worker(void *arg, int tid, int nt) {
    ew_job *j = (ew_job *)arg;
    size_t chunk = (j->n + nt - 1) / nt, lo = chunk * tid, hi = lo + chunk; if (hi > j->n) hi = j->n;
    for (size_t i = lo; i < hi; ++i) {
        switch (j->op) {
        case 0: R_ADD(&j->r[i], &j->a[i], &j->b[i]); break;
        case 1: R_SUB(&j->r[i], &j->a[i], &j->b[i]); break;
        case 2: R_MUL(&j->r[i], &j->a[i], &j->b[i]); break;
        case 3: R_MUL(&j->r[i], &j->a[i], j->s); break;                       /* scale */
        case 4: { fe t; R_MUL(&t, &j->b[i], j->s); R_ADD(&j->r[i], &j->a[i], &t); } break; /* a + s*b */
        }
    }
}

/* ======================================= exported C entry points ======================================== */
#define API __attribute__((visibility("default")))

API void orc_fr_constants(fe *modulus, fe *r1, fe *r2, fe *root_of_unity_mont, fe *zeta_mont) {
    memcpy(modulus->l, FR_M, 32); *r1 = FR_R1; *r2 = FR_R2;
    fr_to_mont(root_of_unity_mont, &FR_ROOT_OF_UNITY_C); fr_to_mont(zeta_mont, &FR_ZETA_C);
}
/* omega of the size-2^k domain, Montgomery form: ROOT_OF_UNITY^(2^(S-k)) (domain.rs EvaluationDomain::new) */
API void orc_fr_omega(uint32_t k, fe *out) {
    fe w; fr_to_mont(&w, &FR_ROOT_OF_UNITY_C);
    for (uint32_t i = k; i < FR_S; ++i) R_MUL(&w, &w, &w);
    *out = w;
}
API void orc_fr_zeta(fe *out) { fr_to_mont(out, &FR_ZETA_C); }
API void orc_fr_inv(const fe *a, fe *out, size_t n) { for (size_t i = 0; i < n; ++i) fr_inv(&out[i], &a[i]); }
API void orc_fq_inv(const fe *a, fe *out, size_t n) { for (size_t i = 0; i < n; ++i) fq_inv(&out[i], &a[i]); }
API void orc_fr_pow(const fe *a, u64 e, fe *out) { fr_pow_u64(out, a, e); }
API void orc_fr_from_mont(const fe *a, fe *out, size_t n) { for (size_t i = 0; i < n; ++i) fr_from_mont(&out[i], &a[i]); }
API void orc_fr_to_mont(const fe *a, fe *out, size_t n) { for (size_t i = 0; i < n; ++i) fr_to_mont(&out[i], &a[i]); }
API void orc_fq_from_mont(const fe *a, fe *out, size_t n) { fe one = {{1, 0, 0, 0}}; for (size_t i = 0; i < n; ++i) Q_MUL(&out[i], &a[i], &one); }
API void orc_fq_to_mont(const fe *a, fe *out, size_t n) { for (size_t i = 0; i < n; ++i) Q_MUL(&out[i], &a[i], &FQ_R2); }
/* field: 0 = Fr, 1 = Fq; op: 0 add 1 sub 2 mul */
API void orc_field_op(int field, int op, const fe *a, const fe *b, fe *out, size_t n) {
    const u64 *m = field ? FQ_M : FR_M; u64 inv = field ? FQ_INV : FR_INV;
    for (size_t i = 0; i < n; ++i) {
        if (op == 0) f_add(&out[i], &a[i], &b[i], m); else if (op == 1) f_sub(&out[i], &a[i], &b[i], m);
        else f_mul(&out[i], &a[i], &b[i], m, inv);
    }
}
/* elementwise polynomial ops (Polynomial +,-,* / parallelize): op 0 add,1 sub,2 mul,3 scale by *s,4 a + s*b */
API void orc_poly_op(int op, const fe *a, const fe *b, const fe *s, fe *out, size_t n, int threads) {
    ew_job j = {op, out, a, b, s, n}; run_threads(ew_worker, &j, threads < 1 ? 1 : threads);
}

/* --- G1 --- */
API int orc_g1_is_on_curve(const g1a *p) {
    if (a_is_identity(p)) return 1;
    fe y2, x3, three, t;
    Q_MUL(&y2, &p->y, &p->y); Q_MUL(&x3, &p->x, &p->x); Q_MUL(&x3, &x3, &p->x);
    Q_ADD(&three, &FQ_R1, &FQ_R1); Q_ADD(&three, &three, &FQ_R1); Q_ADD(&t, &x3, &three);
    return fe_eq(&y2, &t);
}
API void orc_g1_add_affine(const g1a *a, const g1a *b, g1a *out, size_t n) {
    for (size_t i = 0; i < n; ++i) { g1j t; j_from_affine(&t, &a[i]); j_add_mixed(&t, &t, &b[i]); j_to_affine(&out[i], &t); }
}
API void orc_g1_jac_to_affine(const g1j *p, g1a *out, size_t n) { for (size_t i = 0; i < n; ++i) j_to_affine(&out[i], &p[i]); }
static void scalar_mul(g1j *r, const g1a *base, const fe *scalar_mont) {
    fe s; fr_from_mont(&s, scalar_mont);
    g1j acc; j_set_identity(&acc);
    for (int i = 255; i >= 0; --i) { j_double(&acc, &acc); if ((s.l[i / 64] >> (i % 64)) & 1) j_add_mixed(&acc, &acc, base); }
    *r = acc;
}
API void orc_g1_scalar_mul(const g1a *bases, const fe *scalars, g1a *out, size_t n) {
    for (size_t i = 0; i < n; ++i) { g1j t; scalar_mul(&t, &bases[i], &scalars[i]); j_to_affine(&out[i], &t); }
}
/* naive sum_i s_i * P_i by double-and-add: the self-check for best_multiexp */
API void orc_msm_naive(const fe *scalars, const g1a *bases, size_t n, g1a *out) {
    g1j acc, t; j_set_identity(&acc);
    for (size_t i = 0; i < n; ++i) { scalar_mul(&t, &bases[i], &scalars[i]); j_add(&acc, &acc, &t); }
    j_to_affine(out, &acc);
}
/* best_multiexp (ParamsKZG::commit / commit_lagrange) followed by to_affine (batch_normalize) */
API void orc_msm(const fe *scalars, const g1a *bases, size_t n, int threads, g1a *out) {
    g1j acc; best_multiexp(scalars, bases, n, threads, &acc); j_to_affine(out, &acc);
}

/* synthetic distinct bases: out[t*chunk + i] = (start_t + i*step) with start_t = [h(seed,t)]G, step = [h(seed)]G.
 * Jacobian chain + Montgomery batch normalisation per thread chunk. */
static u64 splitmix64(u64 *s) { u64 z = (*s += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
typedef struct { g1a *out; size_t n; u64 seed; } gen_job;
static void gen_worker(void *arg, int tid, int nt) {
    gen_job *j = (gen_job *)arg;
    size_t chunk = (j->n + nt - 1) / nt, lo = chunk * tid, hi = lo + chunk; if (hi > j->n) hi = j->n;
    if (lo >= hi) return;
    g1a G; G.x = FQ_R1; Q_ADD(&G.y, &FQ_R1, &FQ_R1);
    u64 s0 = j->seed; fe k = {{splitmix64(&s0), splitmix64(&s0), splitmix64(&s0), splitmix64(&s0) >> 3}};
    g1j stepj; g1a step; scalar_mul(&stepj, &G, &k); j_to_affine(&step, &stepj);
    u64 s1 = j->seed ^ (0x5851f42d4c957f2dULL * (u64)(tid + 1)); fe k1 = {{splitmix64(&s1), splitmix64(&s1), splitmix64(&s1), splitmix64(&s1) >> 3}};
    g1j cur; scalar_mul(&cur, &G, &k1);
    size_t m = hi - lo;
    g1j *pts = (g1j *)malloc(sizeof(g1j) * m); fe *pref = (fe *)malloc(sizeof(fe) * m);
    for (size_t i = 0; i < m; ++i) { pts[i] = cur; j_add_mixed(&cur, &cur, &step); }
    fe acc = FQ_R1;
    for (size_t i = 0; i < m; ++i) { pref[i] = acc; Q_MUL(&acc, &acc, &pts[i].z); }
    fe inv; fq_inv(&inv, &acc);
    for (size_t i = m; i-- > 0;) {
        fe zi, zi2, zi3; Q_MUL(&zi, &inv, &pref[i]); Q_MUL(&inv, &inv, &pts[i].z);
        Q_MUL(&zi2, &zi, &zi); Q_MUL(&zi3, &zi2, &zi);
        Q_MUL(&j->out[lo + i].x, &pts[i].x, &zi2); Q_MUL(&j->out[lo + i].y, &pts[i].y, &zi3);
    }
    free(pts); free(pref);
}
API void orc_gen_bases(g1a *out, size_t n, u64 seed, int threads) { gen_job j = {out, n, seed}; run_threads(gen_worker, &j, threads < 1 ? 1 : threads); }
/* uniform-ish Fr elements in Montgomery wire form (top limb masked below the modulus' top bits) */
API void orc_gen_scalars(fe *out, size_t n, u64 seed) {
    u64 s = seed;
    for (size_t i = 0; i < n; ++i) { fe v = {{splitmix64(&s), splitmix64(&s), splitmix64(&s), splitmix64(&s) >> 3}}; if (geq(v.l, FR_M)) sub_nb(v.l, v.l, FR_M, 0); out[i] = v; }
}

/* --- NTT / EvaluationDomain --- */
API void orc_best_fft(fe *a, uint32_t log_n, const fe *omega, int threads) { best_fft(a, omega, log_n, threads); }
API void orc_lagrange_to_coeff(fe *a, uint32_t k, int threads) {   /* ifft: omega^-1 then * n^-1 */
    fe w, wi, nn = {{(u64)1 << k, 0, 0, 0}}, ninv; orc_fr_omega(k, &w); fr_inv(&wi, &w);
    fr_to_mont(&nn, &nn); fr_inv(&ninv, &nn);
    best_fft(a, &wi, k, threads);
    ew_job j = {3, a, a, 0, &ninv, (size_t)1 << k}; run_threads(ew_worker, &j, threads < 1 ? 1 : threads);
}
API void orc_coeff_to_lagrange(fe *a, uint32_t k, int threads) { fe w; orc_fr_omega(k, &w); best_fft(a, &w, k, threads); }
/* coeff_to_extended: a[i] *= zeta^(i mod 3), zero-pad to 2^ext_k, best_fft(extended_omega) */
API void orc_coeff_to_extended(const fe *coeffs, size_t n_coeffs, uint32_t ext_k, fe *out, int threads) {
    size_t en = (size_t)1 << ext_k; fe z, z2, w; fr_to_mont(&z, &FR_ZETA_C); R_MUL(&z2, &z, &z); orc_fr_omega(ext_k, &w);
    memset(out, 0, en * sizeof(fe));
    for (size_t i = 0; i < n_coeffs; ++i) {
        if (i % 3 == 0) out[i] = coeffs[i]; else if (i % 3 == 1) R_MUL(&out[i], &coeffs[i], &z); else R_MUL(&out[i], &coeffs[i], &z2);
    }
    best_fft(out, &w, ext_k, threads);
}
/* extended_to_coeff: ifft(extended_omega^-1) * (2^ext_k)^-1, then a[i] *= zeta^-(i mod 3); caller truncates */
API void orc_extended_to_coeff(fe *a, uint32_t ext_k, int threads) {
    size_t en = (size_t)1 << ext_k; fe z, z2, w, wi, nn = {{(u64)1 << ext_k, 0, 0, 0}}, ninv;
    fr_to_mont(&z, &FR_ZETA_C); R_MUL(&z2, &z, &z); orc_fr_omega(ext_k, &w); fr_inv(&wi, &w);
    fr_to_mont(&nn, &nn); fr_inv(&ninv, &nn);
    best_fft(a, &wi, ext_k, threads);
    for (size_t i = 0; i < en; ++i) {
        R_MUL(&a[i], &a[i], &ninv);
        if (i % 3 == 1) R_MUL(&a[i], &a[i], &z2); else if (i % 3 == 2) R_MUL(&a[i], &a[i], &z);
    }
}
/* divide_by_vanishing_poly: a[i] *= t_inv[i mod 2^(ext_k-k)], t[i] = (zeta*omega_ext^i)^n - 1 */
API void orc_divide_by_vanishing(fe *a, uint32_t k, uint32_t ext_k) {
    size_t en = (size_t)1 << ext_k, d = (size_t)1 << (ext_k - k);
    fe z, w, cur, *tinv = (fe *)malloc(sizeof(fe) * d);
    fr_to_mont(&z, &FR_ZETA_C); orc_fr_omega(ext_k, &w); cur = z;
    for (size_t i = 0; i < d; ++i) { fe t; fr_pow_u64(&t, &cur, (u64)1 << k); R_SUB(&t, &t, &FR_R1); fr_inv(&tinv[i], &t); R_MUL(&cur, &cur, &w); }
    for (size_t i = 0; i < en; ++i) R_MUL(&a[i], &a[i], &tinv[i % d]);
    free(tinv);
}
API void orc_eval_polynomial(const fe *coeffs, size_t n, const fe *x, fe *out) {
    fe acc = {{0, 0, 0, 0}};
    for (size_t i = n; i-- > 0;) { R_MUL(&acc, &acc, x); R_ADD(&acc, &acc, &coeffs[i]); }
    *out = acc;
}
API void orc_kate_division(const fe *a, size_t n, const fe *b, fe *q /* n-1 */) {
    fe nb, tmp = {{0, 0, 0, 0}}; f_neg(&nb, b, FR_M);
    for (size_t i = n - 1; i >= 1; --i) { fe lead; R_SUB(&lead, &a[i], &tmp); q[i - 1] = lead; R_MUL(&tmp, &lead, &nb); }
}
/* batch_invert (ff::BatchInvert): zeros stay zero */
API void orc_batch_invert(fe *a, size_t n) {
    fe *pref = (fe *)malloc(sizeof(fe) * (n ? n : 1)); fe acc = FR_R1, inv;
    for (size_t i = 0; i < n; ++i) { pref[i] = acc; if (!fe_is_zero(&a[i])) R_MUL(&acc, &acc, &a[i]); }
    fr_inv(&inv, &acc);
    for (size_t i = n; i-- > 0;) { if (fe_is_zero(&a[i])) continue; fe t; R_MUL(&t, &inv, &pref[i]); R_MUL(&inv, &inv, &a[i]); a[i] = t; }
    free(pref);
}
/* running product / running sum, as used for the permutation z(X) and mv-lookup phi(X) columns:
 * out[0] = init, out[i+1] = out[i] (*|+) a[i]  for i < n-1 (exclusive scan) */
API void orc_prefix_scan(int is_product, const fe *a, size_t n, const fe *init, fe *out) {
    fe acc = *init;
    for (size_t i = 0; i < n; ++i) { out[i] = acc; if (is_product) R_MUL(&acc, &acc, &a[i]); else R_ADD(&acc, &acc, &a[i]); }
}

/* ---- evaluate_h: the quotient-numerator interpreter (UPSTREAM plonk/evaluation.rs GraphEvaluator::evaluate, restated for the
 * instruction format of include/ezkl_b200.h: b200_instr / b200_col_ref).  Row-parallel over threads like halo2's parallelize. */
typedef struct { const fe *const *cols; uint32_t k, ext_k; const uint32_t *loads; const fe *consts; const uint32_t *prog; size_t n_instr; fe *out; } qe_job;
static inline const fe *qe_src(const qe_job *j, uint32_t s, const fe *slots, const fe *prev, size_t idx, fe *tmp) {
    uint32_t kind = s >> 30, i = s & 0x3fffffffu;
    if (kind == 3) return prev;
    if (kind == 0) return &slots[i];
    if (kind == 1) return &j->consts[i];
    uint64_t N = (uint64_t)1 << j->ext_k; int64_t scale = (int64_t)1 << (j->ext_k - j->k);
    int64_t rot = (int32_t)j->loads[2 * i + 1];
    int64_t off = ((rot * scale) % (int64_t)N + (int64_t)N) % (int64_t)N;
    *tmp = j->cols[j->loads[2 * i]][(idx + (uint64_t)off) & (N - 1)];
    return tmp;
}
static void qe_worker(void *arg, int tid, int nt) {
    qe_job *j = (qe_job *)arg;
    size_t N = (size_t)1 << j->ext_k, chunk = (N + nt - 1) / nt, lo = chunk * tid, hi = lo + chunk; if (hi > N) hi = N;
    fe slots[256], ta, tb, tc;
    for (size_t idx = lo; idx < hi; ++idx) {
        fe prev; memset(&prev, 0, sizeof prev);
        for (size_t pc = 0; pc < j->n_instr; ++pc) {
            const uint32_t *in = &j->prog[4 * pc];
            uint32_t op = in[0] & 0xff, dst = (in[0] >> 8) & 255;
            const fe *x = qe_src(j, in[1], slots, &prev, idx, &ta);
            fe r;
            if (op <= 2) {
                const fe *y = qe_src(j, in[2], slots, &prev, idx, &tb);
                if (op == 0) R_ADD(&r, x, y); else if (op == 1) R_SUB(&r, x, y); else R_MUL(&r, x, y);
            } else if (op == 7) {
                const fe *y = qe_src(j, in[2], slots, &prev, idx, &tb), *z = qe_src(j, in[3], slots, &prev, idx, &tc);
                fe t; R_MUL(&t, x, y); R_ADD(&r, &t, z);
            } else if (op == 3) f_neg(&r, x, FR_M);
            else if (op == 4) R_ADD(&r, x, x);
            else if (op == 5) R_MUL(&r, x, x);
            else r = *x;
            if (!(in[0] >> 31)) slots[dst] = r;
            prev = r;
        }
        j->out[idx] = prev;
    }
}
API void orc_quotient_eval(const fe *const *cols, uint32_t k, uint32_t ext_k, const uint32_t *loads, const fe *consts, const uint32_t *prog, size_t n_instr,
                           fe *out, int threads) {
    qe_job j = {cols, k, ext_k, loads, consts, prog, n_instr, out};
    run_threads(qe_worker, &j, threads < 1 ? 1 : threads);
}
