"""The KZG / SHPLONK final check with a REAL pairing on the reference's own SRS fixture (tests/golden/kzg_k6.srs = /root/reference/tests/
assets/kzg, whose trapdoor nobody knows): the G2 half of the file parses as halo2curves lays it out, the fixture satisfies
e(g[i+1], g2) == e(g[i], s_g2), and a proof the create_proof mirror makes ON THAT SRS passes the restated verifier with the pairing the
reference's verifier would run (tests/pairing_bn254.py), while tampered proofs do not.  CPU only (tests/cpu_backend.py)."""
import random

import numpy as np

from ezkl_b200 import prover as pv
from oracle import oracle as orc
from oracle import pyref
from tests import cpu_backend as cb
from tests import helpers as H
from tests import pairing_bn254 as pr
from tests import test_prover_mirror as tpm

R = pyref.R
R_INV_Q = pow(1 << 256, -1, pr.P)
G2_STANDARD = ((10857046999023057135944570762232829481370756359578518086990519993285655852781, 11559732032986387107991004021392285783925812861821192530917403151452391805634),
               (8495653923123431417604973247489272438418190587263600148770280649306958101930, 4082367875863433681332203403145435568316851327593401208105741076214120093531))


def _fq(b: bytes) -> int:
    return int.from_bytes(b, "little") * R_INV_Q % pr.P                 # raw form = Montgomery limbs, little endian


def load_srs():
    """k, g [n,8] wire, g_lagrange [n,8] wire, g2, s_g2 as ((x0, x1), (y0, y1)) python ints (Fq2 = c0 + c1 u, c0 first)."""
    data = open(H.GOLDEN + "/kzg_k6.srs", "rb").read()
    k = int.from_bytes(data[:4], "little")
    n = 1 << k
    g = np.frombuffer(data[4:4 + 64 * n], np.uint64).reshape(n, 8).copy()
    gl = np.frombuffer(data[4 + 64 * n:4 + 128 * n], np.uint64).reshape(n, 8).copy()
    tail = data[4 + 128 * n:]
    assert len(tail) == 256
    pt = lambda b: ((_fq(b[0:32]), _fq(b[32:64])), (_fq(b[64:96]), _fq(b[96:128])))
    return k, g, gl, pt(tail[:128]), pt(tail[128:])


def _g1_xy(wire_row):
    return (_fq(wire_row[:4].tobytes()), _fq(wire_row[4:].tobytes()))


def test_pairing_is_bilinear_and_nondegenerate():
    e = pr.pairing(G2_STANDARD, (1, 2))
    assert e != pr.FQ12.one() and e ** pr.R_ORDER == pr.FQ12.one()
    G = np.array(list(H.fq_wire(1)) + list(H.fq_wire(2)), np.uint64).reshape(1, 8)
    p7 = _g1_xy(orc.g1_scalar_mul(G, H.fr_array([7]))[0])
    assert pr.pairing(G2_STANDARD, p7) == e ** 7


def test_reference_srs_fixture_satisfies_the_kzg_pairing_relation():
    k, g, gl, g2, s_g2 = load_srs()
    assert k == 6 and g2 == G2_STANDARD and pr.g2_on_curve(g2) and pr.g2_on_curve(s_g2)
    for i in (0, 1, 40, 62):
        assert pr.pairing(g2, _g1_xy(g[i + 1])) == pr.pairing(s_g2, _g1_xy(g[i])), i          # g[i+1] = [s] g[i]
    # a Lagrange-basis element against the monomial basis: sum_j g_lagrange[j] = g[0] (the constant polynomial 1)
    acc = gl[0:1]
    for j in range(1, 1 << k):
        acc = orc.g1_add_affine(acc, gl[j:j + 1])
    assert np.array_equal(acc[0], g[0])


class SrsParams:
    """ParamsKZG over the fixture's real bases, commitments by the CPU oracle's MSM."""

    def __init__(self, k, g, g_lagrange):
        self.k, self.n, self.g, self.g_lagrange = k, 1 << k, g, g_lagrange

    def commit(self, poly):
        poly = np.asarray(poly, np.uint64).reshape(-1, 4)
        return cb._jac(orc.msm(poly, self.g[: poly.shape[0]], 2))

    def commit_lagrange(self, col):
        return cb._jac(orc.msm(np.asarray(col, np.uint64).reshape(-1, 4), self.g_lagrange, 2))

    def commit_lagrange_batch(self, cols):
        return np.stack([self.commit_lagrange(c) for c in cols]) if len(cols) else np.zeros((0, 12), np.uint64)

    def commit_batch(self, polys):
        return np.stack([self.commit(p) for p in polys]) if len(polys) else np.zeros((0, 12), np.uint64)


def test_mirror_proof_on_the_reference_srs_passes_the_real_pairing_check(monkeypatch):
    cb.patch_backend(monkeypatch)
    k, g, gl, g2, s_g2 = load_srs()
    _, _, cs, fixed, sigmas, advice = tpm.golden_case()                  # k = 6, the fixture's size
    keys = pv.Keys(SrsParams(k, g, gl), cs, fixed, sigmas, vk_repr=0xE2C1)
    proof = pv.create_proof(keys, advice, rng=pv.ChaCha12Rng(bytes(32)))
    calls = []

    def pairing_check(lhs, pi):
        calls.append((lhs, pi))
        return pr.pairing(g2, lhs) == pr.pairing(s_g2, pi)              # e(L, [1]_2) == e(pi, [s]_2)

    assert pv.verify_proof_with_pairing(keys, proof, pairing_check)
    assert len(calls) == 1 and calls[0][0] is not None and calls[0][1] is not None
    bad = bytearray(proof)
    bad[-100] ^= 1                                                       # inside the evaluations / SHPLONK points
    assert not pv.verify_proof_with_pairing(keys, bytes(bad), pairing_check)
    advice_bad = [list(c) for c in advice]
    advice_bad[0][2] = (advice_bad[0][2] + 1) % R
    assert not pv.verify_proof_with_pairing(keys, pv.create_proof(keys, advice_bad, rng=pv.ChaCha12Rng(bytes(32))), pairing_check)
