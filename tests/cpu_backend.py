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
