"""TEST-ONLY stand-in for `ezkl_b200.halo2` built on the CPU oracle, so that the HOST LOGIC of the Python mirrors (rotation-set
bookkeeping, interpolants, challenge powers, linearisation — code that runs the same whether the polynomial steps execute on the
device or not) can be exercised by the not-gpu test tier.  Tests monkeypatch it into a mirror module (`mo.h2 = cpu_backend`);
the product never imports this file (tests/test_host_logic.py::test_product_does_not_import_oracle keeps checking that)."""
import numpy as np

from oracle import oracle as orc
from oracle import pyref
from tests import helpers as H

R = pyref.R


def eval_polynomial(poly, point) -> np.ndarray:
    return np.asarray(orc.eval_polynomial(poly, point), np.uint64).reshape(4)


def kate_division(a, b) -> np.ndarray:
    return orc.kate_division(a, b)


def poly_op(op: str, a, b=None, s=None) -> np.ndarray:
    return orc.poly_op(op, a, b, s)


def poly_lincomb(polys, scalars) -> np.ndarray:
    acc = orc.poly_op("scale", polys[0], None, scalars[0])
    for p, s in zip(polys[1:], scalars[1:]):
        acc = orc.poly_op("axpy", acc, p, s)
    return acc


class TrapdoorParams:
    """ParamsKZG whose trapdoor s is known: commit(p) = [p(s)] G, returned in the normalised Jacobian wire form (x, y, 1)."""

    def __init__(self, k: int, s: int):
        self.k, self.n, self.s = k, 1 << k, s % R
        self._g = np.array(list(H.fq_wire(1)) + list(H.fq_wire(2)), np.uint64).reshape(1, 8)

    def commit(self, poly) -> np.ndarray:
        v = H.fr_unwire(eval_polynomial(poly, H.fr_wire(self.s)))
        aff = orc.g1_scalar_mul(self._g, H.fr_array([v]))[0]
        return np.concatenate([aff, np.array(list(H.fq_wire(1)), np.uint64)]) if aff.any() else np.array([0] * 4 + list(H.fq_wire(1)) + [0] * 4, np.uint64)


# ---- whole-backend patch: the create_proof mirror (ezkl_b200/prover.py) end to end on the CPU ------------------------------------------
def _jac(aff) -> np.ndarray:
    aff = np.asarray(aff, np.uint64).reshape(8)
    if not aff.any():
        return np.array([0] * 4 + list(H.fq_wire(1)) + [0] * 4, np.uint64)            # identity = (0, 1, 0)
    return np.concatenate([aff, np.array(list(H.fq_wire(1)), np.uint64)])


class Bases:
    def __init__(self, points, window_bits: int = 0):
        self.points = np.ascontiguousarray(np.asarray(points, np.uint64).reshape(-1, 8))

    def release(self):
        pass


def best_multiexp(scalars, bases) -> np.ndarray:
    sc = np.ascontiguousarray(np.asarray(scalars, np.uint64).reshape(-1, 4))
    return _jac(orc.msm(sc, bases.points[: sc.shape[0]], 2))


def batch_invert(a) -> np.ndarray:
    return orc.batch_invert(a)


def prefix_scan(a, init, product: bool) -> np.ndarray:
    return orc.prefix_scan(a, init, product)


class FullTrapdoorParams(TrapdoorParams):
    """commit / commit_lagrange (+ batches) of a ParamsKZG with a known trapdoor: [p(s)] G, Lagrange columns interpolated first."""

    def commit_lagrange(self, col) -> np.ndarray:
        return self.commit(orc.lagrange_to_coeff(col, self.k))

    def commit_lagrange_batch(self, cols) -> np.ndarray:
        return np.stack([self.commit_lagrange(c) for c in cols]) if len(cols) else np.zeros((0, 12), np.uint64)

    def commit_batch(self, polys) -> np.ndarray:
        return np.stack([self.commit(p) for p in polys]) if len(polys) else np.zeros((0, 12), np.uint64)


def patch_backend(monkeypatch):
    """Redirects every device-backed primitive the Python mirrors call to the CPU oracle (host logic under test, arithmetic by the
    checker).  Undone by pytest's monkeypatch at the end of the test."""
    from ezkl_b200 import _native as nat
    from ezkl_b200 import evaluation as ev
    from ezkl_b200 import halo2 as h2
    monkeypatch.setattr(nat, "ensure_init", lambda: None)
    for name, fn in (("eval_polynomial", eval_polynomial), ("kate_division", kate_division), ("poly_op", poly_op), ("poly_lincomb", poly_lincomb),
                     ("batch_invert", batch_invert), ("prefix_scan", prefix_scan), ("best_multiexp", best_multiexp), ("Bases", Bases)):
        monkeypatch.setattr(h2, name, fn)
    D = h2.EvaluationDomain
    monkeypatch.setattr(D, "lagrange_to_coeff", lambda self, a: orc.lagrange_to_coeff(a, self.k))
    monkeypatch.setattr(D, "lagrange_to_coeff_batch", lambda self, cols: [orc.lagrange_to_coeff(c, self.k) for c in cols])
    monkeypatch.setattr(D, "coeff_to_lagrange", lambda self, a: orc.coeff_to_lagrange(a, self.k))
    monkeypatch.setattr(D, "coeff_to_extended", lambda self, a: orc.coeff_to_extended(a, self.extended_k))
    monkeypatch.setattr(D, "coeff_to_extended_batch", lambda self, cols: [orc.coeff_to_extended(c, self.extended_k) for c in cols])
    monkeypatch.setattr(D, "extended_to_coeff", lambda self, a: orc.extended_to_coeff(a, self.extended_k)[: self.n * self.quotient_poly_degree])
    monkeypatch.setattr(D, "divide_by_vanishing_poly", lambda self, a: orc.divide_by_vanishing(a, self.k, self.extended_k))

    def evaluate_h(program, columns, k, ext_k):
        loads, consts, prog = program.arrays()
        return orc.quotient_eval(columns, k, ext_k, loads, consts, prog, 2)

    def evaluate_h_from_polys(program, polys, domain, finish=False):
        N = 1 << domain.extended_k
        cols = [np.asarray(c, np.uint64).reshape(-1, 4) for c in polys]
        cols = [c if c.shape[0] == N else orc.coeff_to_extended(c, domain.extended_k) for c in cols]
        num = evaluate_h(program, cols, domain.k, domain.extended_k)
        return orc.extended_to_coeff(orc.divide_by_vanishing(num, domain.k, domain.extended_k), domain.extended_k) if finish else num

    def lookup_multiplicities(table, inputs, n_rows):
        t = [H.fr_unwire(r) for r in np.asarray(table, np.uint64).reshape(-1, 4)]
        first, m = {}, [0] * len(t)
        for i, v in enumerate(t):
            first.setdefault(v, i)
        for col in inputs:
            for r in np.asarray(col, np.uint64).reshape(-1, 4)[:n_rows]:
                v = H.fr_unwire(r)
                if v not in first:
                    raise nat.B200Error("lookup_multiplicities: an input cell is not in the table")
                m[first[v]] += 1
        return H.fr_array(m)

    monkeypatch.setattr(ev, "evaluate_h", evaluate_h)
    monkeypatch.setattr(ev, "evaluate_h_from_polys", evaluate_h_from_polys)
    monkeypatch.setattr(ev, "lookup_multiplicities", lookup_multiplicities)
