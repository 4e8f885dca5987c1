// ezkl_b200_halo2.hpp — C++ host mirror of the halo2 interfaces on the prover hot path, above the C ABI.
//
// The reference is compiled code (Rust), and its toolchain is not available in this image, so the host side is mirrored in
// C++: same names, argument meaning and error behaviour as halo2_proofs 0.3.0 @ zkonduit/halo2#01c88842
//   arithmetic.rs : best_multiexp, best_fft, eval_polynomial, kate_division
//   poly/domain.rs: EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended, extended_to_coeff,
//                   divide_by_vanishing_poly, extended_len}
//   poly/kzg/commitment.rs: ParamsKZG::{read, commit, commit_lagrange}   (loaded via /root/reference/src/pfsys/srs.rs:30-47)
// ezkl call sites: src/pfsys/mod.rs:390,396,456; src/circuit/modules/polycommit.rs:52,71.
// Errors surface as std::runtime_error carrying b200_last_error() (the Rust shim maps them to plonk::Error).
// Header-only; link with -lezkl_b200.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ezkl_b200.h"

namespace halo2_b200 {

using Fr = b200_fr;
using G1Affine = b200_g1_affine;
using G1 = b200_g1_jac;

inline void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + b200_last_error());
}

// ---- minimal host field arithmetic for domain constants (Montgomery, 4 x u64) ------------------------------------
namespace fr {
typedef unsigned __int128 u128;
static const uint64_t M[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t INV = 0xc2e1f593efffffffULL;
static const Fr ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const Fr R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
// canonical values
static const Fr ROOT_OF_UNITY_C = {{0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL}};
static const Fr ZETA_C = {{0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL}};
static const uint32_t S = 28;

inline bool geq(const uint64_t a[4]) {
    for (int i = 3; i >= 0; --i) { if (a[i] > M[i]) return true; if (a[i] < M[i]) return false; }
    return true;
}
inline Fr mul(const Fr& a, const Fr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        uint64_t c = 0; u128 s;
        for (int j = 0; j < 4; ++j) { s = (u128)a.l[j] * b.l[i] + t[j] + c; t[j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        s = (u128)t[4] + c; t[4] = (uint64_t)s; t[5] = (uint64_t)(s >> 64);
        uint64_t m = t[0] * INV;
        s = (u128)m * M[0] + t[0]; c = (uint64_t)(s >> 64);
        for (int j = 1; j < 4; ++j) { s = (u128)m * M[j] + t[j] + c; t[j - 1] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        s = (u128)t[4] + c; t[3] = (uint64_t)s; t[4] = t[5] + (uint64_t)(s >> 64);
    }
    Fr r;
    if (t[4] || geq(t)) { uint64_t br = 0; for (int i = 0; i < 4; ++i) { u128 d = (u128)t[i] - M[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; } }
    std::memcpy(r.l, t, 32);
    return r;
}
inline Fr sub(const Fr& a, const Fr& b) {
    Fr r; uint64_t br = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)a.l[i] - b.l[i] - br; r.l[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    if (br) { uint64_t c = 0; for (int i = 0; i < 4; ++i) { u128 s = (u128)r.l[i] + M[i] + c; r.l[i] = (uint64_t)s; c = (uint64_t)(s >> 64); } }
    return r;
}
inline Fr to_mont(const Fr& a) { return mul(a, R2); }
inline Fr pow(const Fr& a, const uint64_t e[4]) {
    Fr acc = ONE;
    for (int i = 255; i >= 0; --i) { acc = mul(acc, acc); if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, a); }
    return acc;
}
inline Fr pow_u64(const Fr& a, uint64_t e) { uint64_t ee[4] = {e, 0, 0, 0}; return pow(a, ee); }
inline Fr inv(const Fr& a) {
    uint64_t e[4] = {M[0] - 2, M[1], M[2], M[3]};
    return pow(a, e);
}
inline Fr from_u64(uint64_t v) { Fr c = {{v, 0, 0, 0}}; return to_mont(c); }
}  // namespace fr

// ---- arithmetic.rs --------------------------------------------------------------------------------------------------
// Device-resident base vector; halo2 passes `&[C]` every time, here the SRS vector is registered once per ParamsKZG.
class Bases {
public:
    Bases(const G1Affine* points, size_t n, int window_bits = 0) : n_(n) { check(b200_bases_register(points, n, window_bits, &h_), "bases_register"); }
    ~Bases() { if (h_) b200_bases_release(h_); }
    Bases(const Bases&) = delete;
    Bases& operator=(const Bases&) = delete;
    uint64_t handle() const { return h_; }
    size_t len() const { return n_; }
private:
    uint64_t h_ = 0;
    size_t n_;
};

// best_multiexp(coeffs, bases): upstream asserts coeffs.len() == bases.len(); ParamsKZG::commit slices the bases, so
// coeffs.len() <= bases.len() is accepted here and anything longer is an error.
inline G1 best_multiexp(const std::vector<Fr>& coeffs, const Bases& bases) {
    if (coeffs.size() > bases.len()) throw std::runtime_error("best_multiexp: more coefficients than bases");
    G1 out;
    check(b200_msm(bases.handle(), coeffs.data(), coeffs.size(), &out), "best_multiexp");
    return out;
}
inline std::vector<G1> best_multiexp_batch(const std::vector<const Fr*>& columns, size_t n, const Bases& bases) {
    std::vector<G1> out(columns.size());
    check(b200_msm_batch(bases.handle(), columns.data(), n, columns.size(), out.data()), "best_multiexp_batch");
    return out;
}
// best_fft(a, omega, log_n): in place, natural order in and out; a.len() must equal 1 << log_n (upstream assert_eq).
inline void best_fft(std::vector<Fr>& a, const Fr& omega, uint32_t log_n) {
    if (a.size() != ((size_t)1 << log_n)) throw std::runtime_error("best_fft: a.len() != 1 << log_n");
    check(b200_fft(a.data(), log_n, &omega), "best_fft");
}
inline Fr eval_polynomial(const std::vector<Fr>& poly, const Fr& point) {
    Fr out;
    check(b200_poly_eval(poly.data(), poly.size(), &point, &out), "eval_polynomial");
    return out;
}
inline std::vector<Fr> kate_division(const std::vector<Fr>& a, const Fr& b) {
    if (a.empty()) throw std::runtime_error("kate_division: empty polynomial");
    std::vector<Fr> q(a.size() - 1);
    check(b200_kate_division(a.data(), a.size(), &b, q.data()), "kate_division");
    return q;
}

// multiopen / SHPLONK linear combination q(X) = sum_j scalars[j] * polys[j](X) in one pass
inline std::vector<Fr> poly_lincomb(const std::vector<const Fr*>& polys, const std::vector<Fr>& scalars, size_t n) {
    if (polys.size() != scalars.size()) throw std::runtime_error("poly_lincomb: polys / scalars length mismatch");
    std::vector<Fr> out(n);
    check(b200_poly_lincomb(polys.data(), scalars.data(), polys.size(), n, out.data()), "poly_lincomb");
    return out;
}

// ---- plonk/evaluation.rs ---------------------------------------------------------------------------------------------
// A lowered GraphEvaluator program (see include/ezkl_b200.h for the operand encoding); evaluate_h runs it once per row of the
// extended domain.  columns[c] has 2^extended_k elements; rotations are in rows of the original domain.
struct QuotientProgram {
    std::vector<b200_col_ref> loads;
    std::vector<Fr> constants;
    std::vector<b200_instr> instructions;
    static uint32_t slot(uint32_t i) { return i; }
    static uint32_t constant(uint32_t i) { return (1u << 30) | i; }
    static uint32_t load(uint32_t i) { return (2u << 30) | i; }
    void push(uint32_t op, uint32_t dst_slot, uint32_t a, uint32_t b = 0) { instructions.push_back(b200_instr{op | (dst_slot << 8), a, b}); }
};
inline std::vector<Fr> evaluate_h(const QuotientProgram& prog, const std::vector<const Fr*>& columns, uint32_t k, uint32_t extended_k) {
    std::vector<Fr> out((size_t)1 << extended_k);
    check(b200_quotient_eval(columns.data(), columns.size(), k, extended_k, prog.loads.data(), prog.loads.size(), prog.constants.data(),
                             prog.constants.size(), prog.instructions.data(), prog.instructions.size(), out.data()), "evaluate_h");
    return out;
}

// ---- poly/domain.rs -------------------------------------------------------------------------------------------------
class EvaluationDomain {
public:
    // EvaluationDomain::new(j, k): j = constraint-system degree, n = 2^k
    EvaluationDomain(uint32_t j, uint32_t k) : k_(k), n_((uint64_t)1 << k), quotient_poly_degree_(j - 1) {
        extended_k_ = k;
        while (((uint64_t)1 << extended_k_) < n_ * quotient_poly_degree_) ++extended_k_;
        if (extended_k_ > fr::S) throw std::runtime_error("EvaluationDomain: extended_k exceeds Fr::S");
        Fr root = fr::to_mont(fr::ROOT_OF_UNITY_C);
        extended_omega_ = root;
        for (uint32_t i = extended_k_; i < fr::S; ++i) extended_omega_ = fr::mul(extended_omega_, extended_omega_);
        omega_ = extended_omega_;
        for (uint32_t i = k; i < extended_k_; ++i) omega_ = fr::mul(omega_, omega_);
        omega_inv_ = fr::inv(omega_);
        extended_omega_inv_ = fr::inv(extended_omega_);
        g_coset_ = fr::to_mont(fr::ZETA_C);
        g_coset_inv_ = fr::mul(g_coset_, g_coset_);
        ifft_divisor_ = fr::inv(fr::from_u64(n_));
        extended_ifft_divisor_ = fr::inv(fr::from_u64((uint64_t)1 << extended_k_));
        // t_evaluations[i] = ((zeta * extended_omega^i)^n - 1)^-1, i < 2^(extended_k - k)  (stored inverted, as upstream)
        Fr cur = g_coset_;
        for (uint64_t i = 0; i < ((uint64_t)1 << (extended_k_ - k)); ++i) {
            t_evaluations_.push_back(fr::inv(fr::sub(fr::pow_u64(cur, n_), fr::ONE)));
            cur = fr::mul(cur, extended_omega_);
        }
    }
    uint32_t k() const { return k_; }
    uint32_t extended_k() const { return extended_k_; }
    size_t extended_len() const { return (size_t)1 << extended_k_; }
    const Fr& get_omega() const { return omega_; }
    const Fr& get_extended_omega() const { return extended_omega_; }
    uint64_t get_quotient_poly_degree() const { return quotient_poly_degree_; }

    void lagrange_to_coeff(std::vector<Fr>& a) const {
        if (a.size() != n_) throw std::runtime_error("lagrange_to_coeff: wrong length");
        check(b200_ifft(a.data(), k_, &omega_inv_, &ifft_divisor_), "lagrange_to_coeff");
    }
    void coeff_to_lagrange(std::vector<Fr>& a) const {
        if (a.size() != n_) throw std::runtime_error("coeff_to_lagrange: wrong length");
        check(b200_fft(a.data(), k_, &omega_), "coeff_to_lagrange");
    }
    std::vector<Fr> coeff_to_extended(const std::vector<Fr>& a) const {
        if (a.size() != n_) throw std::runtime_error("coeff_to_extended: wrong length");
        std::vector<Fr> out(extended_len());
        check(b200_coeff_to_extended(a.data(), a.size(), extended_k_, &extended_omega_, &g_coset_, out.data()), "coeff_to_extended");
        return out;
    }
    // returns n * quotient_poly_degree coefficients (upstream truncates identically)
    std::vector<Fr> extended_to_coeff(std::vector<Fr> a) const {
        if (a.size() != extended_len()) throw std::runtime_error("extended_to_coeff: wrong length");
        check(b200_extended_to_coeff(a.data(), extended_k_, &extended_omega_inv_, &extended_ifft_divisor_, &g_coset_), "extended_to_coeff");
        a.resize(n_ * quotient_poly_degree_);
        return a;
    }
    void divide_by_vanishing_poly(std::vector<Fr>& a) const {
        if (a.size() != extended_len()) throw std::runtime_error("divide_by_vanishing_poly: wrong length");
        check(b200_poly_scale_cycle(a.data(), a.size(), t_evaluations_.data(), (uint32_t)t_evaluations_.size()), "divide_by_vanishing_poly");
    }

private:
    uint32_t k_, extended_k_;
    uint64_t n_, quotient_poly_degree_;
    Fr omega_, omega_inv_, extended_omega_, extended_omega_inv_, g_coset_, g_coset_inv_, ifft_divisor_, extended_ifft_divisor_;
    std::vector<Fr> t_evaluations_;
};

// ---- poly/kzg/commitment.rs -------------------------------------------------------------------------------------------
// ParamsKZG<Bn256>: file layout of ParamsKZG::write = u32 LE k | g[n] | g_lagrange[n] | g2 (128 B) | s_g2 (128 B).
class ParamsKZG {
public:
    static ParamsKZG read(const std::string& path) {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("ParamsKZG::read: cannot open " + path);
        ParamsKZG p;
        uint32_t k = 0;
        bool ok = std::fread(&k, 4, 1, f) == 1 && k <= fr::S;
        if (ok) {
            p.k_ = k; p.n_ = (size_t)1 << k;
            p.g_.resize(p.n_); p.g_lagrange_.resize(p.n_); p.tail_.resize(256);
            ok = std::fread(p.g_.data(), 64, p.n_, f) == p.n_ && std::fread(p.g_lagrange_.data(), 64, p.n_, f) == p.n_ &&
                 std::fread(p.tail_.data(), 1, 256, f) == 256;
        }
        std::fclose(f);
        if (!ok) throw std::runtime_error("ParamsKZG::read: truncated or malformed SRS file");
        return p;
    }
    uint32_t k() const { return k_; }
    size_t n() const { return n_; }
    const std::vector<G1Affine>& get_g() const { return g_; }
    const std::vector<G1Affine>& get_g_lagrange() const { return g_lagrange_; }
    // commit(poly, _blind): coefficient form against g[..len]; the blind is ignored for KZG, as upstream
    G1 commit(const std::vector<Fr>& poly) {
        if (!bases_g_) bases_g_.reset(new Bases(g_.data(), n_));
        return best_multiexp(poly, *bases_g_);
    }
    // commit_lagrange(poly, _blind): Lagrange form against g_lagrange; poly.len() must equal n (upstream assert)
    G1 commit_lagrange(const std::vector<Fr>& poly) {
        if (poly.size() != n_) throw std::runtime_error("commit_lagrange: poly.len() != n");
        if (!bases_l_) bases_l_.reset(new Bases(g_lagrange_.data(), n_));
        return best_multiexp(poly, *bases_l_);
    }

private:
    struct Del { void operator()(Bases* b) const { delete b; } };
    uint32_t k_ = 0;
    size_t n_ = 0;
    std::vector<G1Affine> g_, g_lagrange_;
    std::vector<uint8_t> tail_;
    std::unique_ptr<Bases, Del> bases_g_, bases_l_;
};

}  // namespace halo2_b200
