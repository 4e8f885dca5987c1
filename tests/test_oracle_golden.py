"""Pins the CPU oracle (oracle/bn254_oracle.c) against the reference's own fixtures and the bigint restatement.

Mirrors what the reference's prove->verify tests exercise implicitly (tests/integration_tests.rs:1313-1425) at the
arithmetic level: SRS relations from tests/assets/kzg, NTT/coset relations from tests/assets/pk.key.
"""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import pyref
from tests import helpers as H


def test_constants():
    assert H.fr_unwire(orc.omega(28)) == pyref.FR_ROOT_OF_UNITY
    for k in (1, 6, 9, 17, 22):
        assert H.fr_unwire(orc.omega(k)) == pyref.omega_for(k)
    assert H.fr_unwire(orc.zeta()) == pyref.FR_ZETA
    assert H.fr_unwire(orc.fr_one()) == 1


def test_field_ops_vs_bigint():
    rng = random.Random(1)
    for field, mod in (("fr", pyref.R), ("fq", pyref.P)):
        xs = [rng.randrange(mod) for _ in range(200)] + [0, 1, mod - 1, mod - 2]
        ys = [rng.randrange(mod) for _ in range(200)] + [mod - 1, 0, mod - 1, 2]
        a = np.stack([H.int_to_limbs(pyref.to_mont(x, mod)) for x in xs])
        b = np.stack([H.int_to_limbs(pyref.to_mont(y, mod)) for y in ys])
        for op, f in (("add", lambda x, y: x + y), ("sub", lambda x, y: x - y), ("mul", lambda x, y: x * y)):
            got = orc.field_op(field, op, a, b)
            exp = [pyref.to_mont(f(x, y) % mod, mod) for x, y in zip(xs, ys)]
            assert [H.limbs_to_int(r) for r in got] == exp, (field, op)
    xs = [rng.randrange(1, pyref.R) for _ in range(20)]
    inv = orc.fr_inv(H.fr_array(xs))
    assert H.fr_list(inv) == [pow(x, -1, pyref.R) for x in xs]


def test_srs_fixture_msm_known_answers():
    """g_lagrange[j] == MSM(n^-1 omega^-ij, g) for all 64 j: 64 size-64 MSM KATs from the reference's SRS."""
    k, g, gl = H.load_srs_fixture()
    n = 1 << k
    assert all(orc.g1_is_on_curve(p) for p in g) and all(orc.g1_is_on_curve(p) for p in gl)
    w_inv = pow(pyref.omega_for(k), -1, pyref.R)
    n_inv = pow(n, -1, pyref.R)
    for j in range(n):
        sc = H.fr_array([pow(w_inv, i * j, pyref.R) * n_inv % pyref.R for i in range(n)])
        for threads in (1, 3):
            assert np.array_equal(orc.msm(sc, g, threads), gl[j]), (j, threads)
    # sum of the Lagrange basis commitments is the commitment to the constant 1 = g[0]
    ones = np.tile(orc.fr_one(), (n, 1))
    assert np.array_equal(orc.msm(ones, gl, 2), g[0])


def test_pk_fixture_ntt_known_answers():
    pk = H.load_pk_fixture()
    k, ext_k = 6, 9
    cols = [("fixed_values_%d" % c, "fixed_polys_%d" % c, "fixed_cosets_%d" % c) for c in (0, 1, 5, 37)]
    cols.append(("perm_values_0", "perm_polys_0", "perm_cosets_0"))
    for v, p, c in cols:
        for threads in (1, 4):
            assert np.array_equal(orc.coeff_to_lagrange(pk[p], k, threads), pk[v])
            assert np.array_equal(orc.lagrange_to_coeff(pk[v], k, threads), pk[p])
            assert np.array_equal(orc.coeff_to_extended(pk[p], ext_k, threads), pk[c])
            back = orc.extended_to_coeff(pk[c], ext_k, threads)
            assert np.array_equal(back[:64], pk[p]) and not back[64:].any()
    # l0 on the extended coset
    l0 = np.zeros((64, 4), np.uint64)
    l0[0] = orc.fr_one()
    assert np.array_equal(orc.coeff_to_extended(orc.lagrange_to_coeff(l0, k), ext_k), pk["l0"])


@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 8])
def test_fft_vs_bigint(log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    xs = [rng.randrange(pyref.R) for _ in range(n)]
    w = pyref.omega_for(log_n)
    exp = pyref.best_fft(xs, w, log_n)
    if log_n <= 5:
        assert exp == pyref.dft_naive(xs, w)
    for threads in (1, 2, 8, 64):
        assert H.fr_list(orc.best_fft(H.fr_array(xs), log_n, H.fr_wire(w), threads)) == exp


def test_msm_vs_naive_and_bigint():
    rng = random.Random(7)
    n = 70
    bases = orc.gen_bases(n, seed=3, threads=3)
    assert all(orc.g1_is_on_curve(p) for p in bases)
    assert len({bytes(p) for p in bases}) == n
    xs = [rng.randrange(pyref.R) for _ in range(n)]
    xs[3], xs[4], xs[5], xs[6] = 0, 1, pyref.R - 1, 5
    sc = H.fr_array(xs)
    exp = pyref.msm_naive(xs, [H.g1_unwire(p) for p in bases])
    assert H.g1_unwire(orc.msm_naive(sc, bases)) == exp
    for threads in (1, 2, 7, 100):
        assert H.g1_unwire(orc.msm(sc, bases, threads)) == exp
    # degenerate inputs the reference tolerates: all-equal points (doubling inside buckets), identity bases, zeros
    same = np.tile(bases[0], (n, 1))
    assert H.g1_unwire(orc.msm(sc, same, 2)) == pyref.g1_mul(H.g1_unwire(bases[0]), sum(xs))
    withid = bases.copy()
    withid[::3] = 0
    exp = pyref.msm_naive(xs, [H.g1_unwire(p) for p in withid])
    assert H.g1_unwire(orc.msm(sc, withid, 3)) == exp
    assert not orc.msm(np.zeros((n, 4), np.uint64), bases, 2).any()
    small = H.fr_array([rng.randrange(1 << 12) for _ in range(n)])
    assert np.array_equal(orc.msm(small, bases, 3), orc.msm_naive(small, bases))


def test_poly_helpers_vs_bigint():
    rng = random.Random(11)
    n = 33
    xs = [rng.randrange(pyref.R) for _ in range(n)]
    x = rng.randrange(pyref.R)
    a = H.fr_array(xs)
    assert H.fr_unwire(orc.eval_polynomial(a, H.fr_wire(x))) == pyref.eval_polynomial(xs, x)
    assert H.fr_list(orc.kate_division(a, H.fr_wire(x))) == pyref.kate_division(xs, x)
    ys = list(xs)
    ys[2] = ys[9] = 0
    inv = H.fr_list(orc.batch_invert(H.fr_array(ys)))
    assert inv == [pow(y, -1, pyref.R) if y else 0 for y in ys]
    pp = H.fr_list(orc.prefix_scan(a, H.fr_wire(1), True))
    acc, exp = 1, []
    for v in xs:
        exp.append(acc)
        acc = acc * v % pyref.R
    assert pp == exp
    ext = [rng.randrange(pyref.R) for _ in range(1 << 5)]
    got = H.fr_list(orc.divide_by_vanishing(H.fr_array(ext), 3, 5))
    d = 4
    t = [(pow(pyref.FR_ZETA * pow(pyref.omega_for(5), i, pyref.R), 8, pyref.R) - 1) % pyref.R for i in range(d)]
    assert got == [e * pow(t[i % d], -1, pyref.R) % pyref.R for i, e in enumerate(ext)]
    s = rng.randrange(pyref.R)
    b = H.fr_array(ys)
    assert H.fr_list(orc.poly_op("axpy", a, b, H.fr_wire(s), threads=3)) == [(u + s * v) % pyref.R for u, v in zip(xs, ys)]
    assert H.fr_list(orc.poly_op("mul", a, b, threads=2)) == [u * v % pyref.R for u, v in zip(xs, ys)]
