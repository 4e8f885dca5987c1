// ec.cuh — BN254 G1 (y^2 = x^3 + 3 over Fq) group law, host + device.
//
// Replaces halo2curves 0.7.0 src/bn256/curve.rs (G1Affine / G1) as used by halo2's best_multiexp
// (UPSTREAM, called from /root/reference/src/circuit/modules/polycommit.rs:71 and every commit in create_proof,
// /root/reference/src/pfsys/mod.rs:456).  Wire formats: G1Affine = {x,y} 64 B, identity (0,0);
// G1Jac = {x,y,z} 96 B with x = X/Z^2, y = Y/Z^3, identity z = 0.
//
// Bucket accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): mixed addition
// of an affine base costs 8M + 2S and needs no inversion.  All formulas handle identity / doubling / inverse inputs
// (ezkl witness columns are full of repeated small scalars, so coincident bucket contents are ordinary inputs).
#pragma once
#include "field.cuh"

namespace b200 {

struct alignas(16) G1Affine { Fq x, y; };
struct alignas(16) G1Xyzz { Fq x, y, zz, zzz; };
struct alignas(16) G1Jac { Fq x, y, z; };

HD bool g1_is_identity(const G1Affine& p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
HD bool g1_is_identity(const G1Xyzz& p) { return fp_is_zero(p.zz); }
HD G1Xyzz g1_xyzz_identity() {
    G1Xyzz r; r.x = fp_zero<FqTag>(); r.y = fp_zero<FqTag>(); r.zz = fp_zero<FqTag>(); r.zzz = fp_zero<FqTag>();
    return r;
}
HD G1Xyzz g1_to_xyzz(const G1Affine& p) {
    G1Xyzz r;
    if (g1_is_identity(p)) return g1_xyzz_identity();
    r.x = p.x; r.y = p.y; r.zz = fp_one<FqTag>(); r.zzz = fp_one<FqTag>();
    return r;
}
HD G1Affine g1_neg(const G1Affine& p) {
    G1Affine r; r.x = p.x; r.y = fp_neg(p.y);   // identity (0,0) stays (0,0)
    return r;
}
// 2 * (affine p) in XYZZ  (mdbl-2008-s-1, a = 0)
HD G1Xyzz g1_dbl_affine(const G1Affine& p) {
    if (g1_is_identity(p)) return g1_xyzz_identity();
    G1Xyzz r;
    Fq u = fp_dbl(p.y), v = fp_sqr(u), w = u * v, s = p.x * v;
    Fq xx = fp_sqr(p.x), m = fp_dbl(xx) + xx;
    r.x = fp_sqr(m) - fp_dbl(s);
    r.y = fp_mulsub2(m, s - r.x, w, p.y);
    r.zz = v; r.zzz = w;
    return r;
}
// 2 * p  (dbl-2008-s-1, a = 0).  A point with y = 0 cannot exist on this curve (b = 3 is not a cube residue issue:
// order is prime), so no special case beyond identity.
HD G1Xyzz g1_dbl(const G1Xyzz& p) {
    if (g1_is_identity(p)) return p;
    G1Xyzz r;
    Fq u = fp_dbl(p.y), v = fp_sqr(u), w = u * v, s = p.x * v;
    Fq xx = fp_sqr(p.x), m = fp_dbl(xx) + xx;
    r.x = fp_sqr(m) - fp_dbl(s);
    r.y = fp_mulsub2(m, s - r.x, w, p.y);
    r.zz = v * p.zz; r.zzz = w * p.zzz;
    return r;
}
// acc + (affine q)  (madd-2008-s), complete.
HD G1Xyzz g1_add_mixed(const G1Xyzz& a, const G1Affine& q) {
    if (g1_is_identity(q)) return a;
    if (g1_is_identity(a)) return g1_to_xyzz(q);
    Fq u2 = q.x * a.zz, s2 = q.y * a.zzz;
    Fq p = u2 - a.x, r = s2 - a.y;
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) return g1_dbl_affine(q);
        return g1_xyzz_identity();
    }
    G1Xyzz o;
    Fq pp = fp_sqr(p), ppp = p * pp, qq = a.x * pp;
    o.x = fp_sqr(r) - ppp - fp_dbl(qq);
    o.y = fp_mulsub2(r, qq - o.x, a.y, ppp);
    o.zz = a.zz * pp; o.zzz = a.zzz * ppp;
    return o;
}
// a + b  (add-2008-s), complete.
HD G1Xyzz g1_add(const G1Xyzz& a, const G1Xyzz& b) {
    if (g1_is_identity(a)) return b;
    if (g1_is_identity(b)) return a;
    Fq u1 = a.x * b.zz, u2 = b.x * a.zz, s1 = a.y * b.zzz, s2 = b.y * a.zzz;
    Fq p = u2 - u1, r = s2 - s1;
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) return g1_dbl(a);
        return g1_xyzz_identity();
    }
    G1Xyzz o;
    Fq pp = fp_sqr(p), ppp = p * pp, qq = u1 * pp;
    o.x = fp_sqr(r) - ppp - fp_dbl(qq);
    o.y = fp_mulsub2(r, qq - o.x, s1, ppp);
    o.zz = a.zz * b.zz * pp; o.zzz = a.zzz * b.zzz * ppp;
    return o;
}
// XYZZ -> affine (one inversion): 1/zzz, then 1/zz = zzz^-1 ... derive from invariant zz^3 = zzz^2:
//   zinv = zz / zzz  (= 1/z),  x = X * zinv^2,  y = Y * zinv^3
HD G1Affine g1_to_affine(const G1Xyzz& p) {
    G1Affine r;
    if (g1_is_identity(p)) { r.x = fp_zero<FqTag>(); r.y = fp_zero<FqTag>(); return r; }
    Fq zzz_inv = fp_inv(p.zzz);
    Fq zinv = p.zz * zzz_inv;
    Fq zinv2 = fp_sqr(zinv);
    r.x = p.x * zinv2;
    r.y = p.y * (zinv2 * zinv);
    return r;
}
// affine -> Jacobian wire form with z = 1 (identity: halo2curves' G1::identity() = (0, 1, 0))
HD G1Jac g1_affine_to_jac(const G1Affine& p) {
    G1Jac r;
    if (g1_is_identity(p)) { r.x = fp_zero<FqTag>(); r.y = fp_one<FqTag>(); r.z = fp_zero<FqTag>(); return r; }
    r.x = p.x; r.y = p.y; r.z = fp_one<FqTag>();
    return r;
}
HD bool g1_is_on_curve(const G1Affine& p) {
    if (g1_is_identity(p)) return true;
    Fq one = fp_one<FqTag>();
    Fq three = one + one + one;
    return fp_eq(fp_sqr(p.y), fp_sqr(p.x) * p.x + three);
}
// k * p for small k; used for the bucket-index weighting in the reduction.  Fixed 2-bit windows over {p, 2p, 3p}: the lanes of a warp
// carry different k, so a bit-by-bit double-and-add executes BOTH the doubling and the addition of every bit for the whole warp
// (divergence); here every window costs two doublings and one addition whatever the digits are.
HD G1Xyzz g1_mul_small(const G1Xyzz& p, uint32_t k) {
    if (k == 0) return g1_xyzz_identity();
    const G1Xyzz p2 = g1_dbl(p), p3 = g1_add(p2, p);
    int top = 30;
    while (top > 0 && !((k >> top) & 3u)) top -= 2;
    uint32_t d = (k >> top) & 3u;
    G1Xyzz acc = d == 1 ? p : (d == 2 ? p2 : p3);
#pragma unroll 1
    for (int i = top - 2; i >= 0; i -= 2) {
        acc = g1_dbl(g1_dbl(acc));
        d = (k >> i) & 3u;
        if (d) acc = g1_add(acc, d == 1 ? p : (d == 2 ? p2 : p3));
    }
    return acc;
}

}  // namespace b200
