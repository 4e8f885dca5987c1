"""Pure-Python bigint restatement of the BN254 arithmetic on the prover hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ezkl_b200/`` may import this module; it is used by
``tests/`` and ``tests/golden/make_golden.py`` to (a) validate the C oracle (``oracle/bn254_oracle.c``) on
small cases and (b) pin both against the reference's checked-in fixtures
(``/root/reference/tests/assets/{kzg,pk.key}``).

The algorithms live in un-vendored dependencies of the reference:
  * halo2_proofs 0.3.0 @ zkonduit/halo2#01c88842   (arithmetic.rs: best_fft / best_multiexp /
    eval_polynomial / kate_division;  poly/domain.rs: EvaluationDomain)
  * halo2curves 0.7.0 @ privacy-scaling-explorations/halo2curves#b753a832 (bn256 Fr/Fq/G1)
Their in-tree call sites: /root/reference/src/pfsys/mod.rs:390,396,456 ; src/pfsys/srs.rs:15,36,46 ;
src/circuit/modules/polycommit.rs:52,71.
"""
from __future__ import annotations

# ---- constants ------------------------------------------------------------
P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # Fq modulus
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # Fr modulus
MONT = 1 << 256
FR_S = 28
FR_GENERATOR = 7
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R - 1) >> FR_S, R)
FR_ZETA = pow(FR_GENERATOR, 2 * (R - 1) // 3, R)
assert FR_ROOT_OF_UNITY == 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c
assert FR_ZETA == 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
CURVE_B = 3


def to_mont(x: int, m: int) -> int:
    return (x * MONT) % m


def from_mont(x: int, m: int) -> int:
    return (x * pow(MONT, -1, m)) % m


def le32(x: int) -> bytes:
    return x.to_bytes(32, "little")


def fr_from_wire(b: bytes) -> int:
    """32 B little-endian Montgomery limbs (halo2curves SerdeObject raw form) -> canonical int."""
    return from_mont(int.from_bytes(b, "little"), R)


def fr_to_wire(x: int) -> bytes:
    return le32(to_mont(x % R, R))


def fq_from_wire(b: bytes) -> int:
    return from_mont(int.from_bytes(b, "little"), P)


def fq_to_wire(x: int) -> bytes:
    return le32(to_mont(x % P, P))


def g1_from_wire(b: bytes):
    """64 B affine (x‖y Montgomery LE); identity is (0,0) -> None."""
    x, y = fq_from_wire(b[:32]), fq_from_wire(b[32:64])
    if x == 0 and y == 0:
        return None
    return (x, y)


def g1_to_wire(pt) -> bytes:
    if pt is None:
        return bytes(64)
    return fq_to_wire(pt[0]) + fq_to_wire(pt[1])


# ---- G1 affine arithmetic (canonical; y^2 = x^3 + 3) -----------------------------------------------
def g1_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - CURVE_B) % P == 0


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_neg(a):
    return None if a is None else (a[0], (-a[1]) % P)


def g1_mul(a, s: int):
    s %= R
    acc = None
    while s:
        if s & 1:
            acc = g1_add(acc, a)
        a = g1_add(a, a)
        s >>= 1
    return acc


def msm_naive(scalars, bases):
    acc = None
    for s, b in zip(scalars, bases):
        acc = g1_add(acc, g1_mul(b, s))
    return acc


# ---- halo2 EvaluationDomain / best_fft semantics ------------------------
def omega_for(k: int) -> int:
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_S - k), R)


def bitrev(i: int, bits: int) -> int:
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def best_fft(a, omega: int, log_n: int):
    """Natural-order in/out radix-2 DIT; out[j] = sum_i a[i] * omega^(i*j)."""
    n = 1 << log_n
    a = list(a)
    assert len(a) == n
    for i in range(n):
        r = bitrev(i, log_n)
        if i < r:
            a[i], a[r] = a[r], a[i]
    m = 1
    for _ in range(log_n):
        wm = pow(omega, n // (2 * m), R)
        for s in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[s + j + m] * w % R
                u = a[s + j]
                a[s + j] = (u + t) % R
                a[s + j + m] = (u - t) % R
                w = w * wm % R
        m *= 2
    return a


def dft_naive(a, omega: int):
    n = len(a)
    return [sum(a[i] * pow(omega, i * j, R) for i in range(n)) % R for j in range(n)]


def lagrange_to_coeff(vals, k: int):
    w_inv = pow(omega_for(k), -1, R)
    n_inv = pow(1 << k, -1, R)
    return [x * n_inv % R for x in best_fft(vals, w_inv, k)]


def coeff_to_extended(coeffs, k: int, ext_k: int):
    z = [1, FR_ZETA, FR_ZETA * FR_ZETA % R]
    a = [c * z[i % 3] % R for i, c in enumerate(coeffs)]
    a += [0] * ((1 << ext_k) - len(a))
    return best_fft(a, omega_for(ext_k), ext_k)


def extended_to_coeff(ext, ext_k: int):
    w_inv = pow(omega_for(ext_k), -1, R)
    div = pow(1 << ext_k, -1, R)
    a = [x * div % R for x in best_fft(ext, w_inv, ext_k)]
    zi = [1, FR_ZETA * FR_ZETA % R, FR_ZETA]
    return [c * zi[i % 3] % R for i, c in enumerate(a)]


def eval_polynomial(coeffs, x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def kate_division(a, b: int):
    """Quotient of a(X) / (X - b), remainder dropped; len(a)-1 coefficients (halo2 arithmetic.rs)."""
    b = (-b) % R
    q = [0] * (len(a) - 1)
    tmp = 0
    for i in range(len(a) - 1, 0, -1):
        lead = (a[i] - tmp) % R
        q[i - 1] = lead
        tmp = lead * b % R
    return q
