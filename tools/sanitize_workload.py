#!/usr/bin/env python
"""Small end-to-end workload touching every kernel family, for `compute-sanitizer --tool memcheck|racecheck`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from ezkl_b200 import _native as nat  # noqa: E402
from ezkl_b200 import evaluation as ev  # noqa: E402
from ezkl_b200 import fields as F  # noqa: E402
from ezkl_b200 import halo2 as h2  # noqa: E402

nat.init(0)
k = 8
n = 1 << k
rng = np.random.default_rng(1)
params = h2.ParamsKZG.setup(k, 0x123456789)                       # fixed-base mul, scans, batch invert
cols = []
for i in range(4):
    a = rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    cols.append(a)
small = np.zeros((n, 4), np.uint64)
small[:, 0] = 7
small = h2.poly_op("scale", small, s=F.fr_to_limbs(pow(2, 256, F.FR_MODULUS)))      # heavy bucket column
params.commit_lagrange_batch(cols + [small])                         # MSM pipeline incl. k_combine_heavy
dom = h2.EvaluationDomain(5, k)
coeffs = dom.lagrange_to_coeff_batch(cols)                           # NTT v2 / v1
exts = dom.coeff_to_extended_batch(coeffs)
prog = ev.QuotientProgram((ev.Query(0) * ev.Query(1, 1) + ev.Query(2, -1)) * ev.Constant(5) - ev.Query(3))
h = ev.evaluate_h(prog, exts, k, dom.extended_k)                     # quotient interpreter
dom.extended_to_coeff(dom.divide_by_vanishing_poly(h))
h2.poly_lincomb(coeffs, np.stack([F.fr_to_limbs(i + 2) for i in range(4)]))
h2.kate_division(coeffs[0], F.fr_to_limbs(9))
h2.eval_polynomial_batch(coeffs, np.stack([F.fr_to_limbs(3)] * 4))
h2.prefix_scan(cols[0], F.fr_to_limbs(1), True)
h2.g_to_lagrange(params.g[:16], 4)                                   # group FFT
print("sanitize workload done")
