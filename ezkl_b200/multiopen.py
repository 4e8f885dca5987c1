"""Host mirror of halo2's SHPLONK multi-opening prover (UPSTREAM poly/kzg/multiopen/shplonk/prover.rs — the type ezkl selects at
/root/reference/src/execute.rs:1604, `ProverSHPLONK<_>`; stage 9 of create_proof) on top of the C ABI:
every polynomial-sized step (the y / v linear combinations, the divisions by (X - point), both commitments, the sanity
evaluation) runs on the device; only rotation-set bookkeeping and the tiny low-degree interpolants stay on the host.

The transcript (challenges y, v, u and the order points are written in) belongs to the Rust side; here the challenges are
arguments, and `create_proof` returns the two commitments plus the intermediate polynomials so tests can check the algebra.
"""
from __future__ import annotations

import numpy as np

from . import fields as F
from . import halo2 as h2

R = F.FR_MODULUS


def _poly_from_ints(coeffs, n):
    out = np.zeros((n, 4), np.uint64)
    for i, c in enumerate(coeffs):
        out[i] = F.fr_to_limbs(c)
    return out


def lagrange_interpolate(points, evals):
    """Coefficients (python ints, low degree first) of the unique polynomial of degree < len(points) through (points, evals)."""
    m = len(points)
    coeffs = [0] * m
    for i in range(m):
        num = [1]                                           # prod_{j != i} (X - x_j)
        den = 1
        for j in range(m):
            if j == i:
                continue
            num = [((num[t - 1] if t > 0 else 0) - points[j] * (num[t] if t < len(num) else 0)) % R for t in range(len(num) + 1)]
            den = den * (points[i] - points[j]) % R
        scale = evals[i] * pow(den, -1, R) % R
        for t in range(len(num)):
            coeffs[t] = (coeffs[t] + num[t] * scale) % R
    return coeffs


def evaluate_vanishing_polynomial(roots, z):
    acc = 1
    for r in roots:
        acc = acc * (z - r) % R
    return acc


def _powers(base: int, count: int):
    """[1, base, base^2, ...]: SHPLONK combines with `.zip(powers(y))` / `.zip(powers(v))` (UPSTREAM shplonk/prover.rs and
    shplonk/verifier.rs, PSE halo2 v0.3 lineage) — ASCENDING powers, the first polynomial / rotation set gets the exponent 0.
    (The descending Horner fold `acc * base + p` is the GWC multiopen's rule, not SHPLONK's.)"""
    return [pow(base, j, R) for j in range(count)]


class ProverQuery:
    """One opening claim: polynomial `poly` ([n,4] wire coefficients) at `point` (python int)."""

    def __init__(self, point: int, poly: np.ndarray):
        self.point, self.poly = point % R, poly


def construct_rotation_sets(queries):
    """Group polynomials by the SET of points they are opened at (halo2's construct_intermediate_sets): returns
    [(points tuple, [poly, ...])] in first-appearance order, and the ordered union of all points."""
    poly_points, order = {}, []
    for q in queries:
        key = q.poly.ctypes.data
        if key not in poly_points:
            poly_points[key] = (q.poly, [])
            order.append(key)
        if q.point not in poly_points[key][1]:
            poly_points[key][1].append(q.point)
    sets, set_order = {}, []
    for key in order:
        poly, pts = poly_points[key]
        sk = tuple(sorted(pts))
        if sk not in sets:
            sets[sk] = []
            set_order.append(sk)
        sets[sk].append(poly)
    super_points = []
    for sk in set_order:
        for p in sk:
            if p not in super_points:
                super_points.append(p)
    return [(sk, sets[sk]) for sk in set_order], super_points


def create_proof(params: h2.ParamsKZG, queries, y: int = None, v: int = None, u: int = None, transcript=None):
    """ProverSHPLONK::create_proof.  With `transcript` (ezkl_b200.transcript.EvmTranscriptWrite) the challenges are squeezed and the
    two commitments written exactly where upstream does (y, v, write h1, u, write h2); without it the three challenges are
    arguments.  Returns a dict with the commitments `h1`, `h2` (normalised Jacobian wire) and the polynomials behind them."""
    n = params.n
    if transcript is not None:
        y, v = transcript.squeeze_challenge(), transcript.squeeze_challenge()
    sets, super_points = construct_rotation_sets(queries)
    one = F.fr_to_limbs(1)
    quotient_polys, set_numerators, set_r = [], [], []
    for points, polys in sets:
        # N_i(X) = sum_j y^j (P_ij(X) - R_ij(X)),  R_ij = low-degree interpolant of P_ij on the set's points
        r_polys = [lagrange_interpolate(list(points), [F.fr_from_limbs(h2.eval_polynomial(p, F.fr_to_limbs(x))) for x in points]) for p in polys]
        ys = _powers(y, len(polys))
        n_x = h2.poly_lincomb(polys, np.stack([F.fr_to_limbs(s) for s in ys]))
        r_comb = [sum(ys[j] * r_polys[j][t] for j in range(len(polys))) % R for t in range(len(points))]
        n_x = h2.poly_op("sub", n_x, _poly_from_ints(r_comb, n))
        q_x = n_x
        for x in points:                                     # div_by_vanishing: one kate_division per point of the set
            q_x = np.concatenate([h2.kate_division(q_x, F.fr_to_limbs(x)), np.zeros((1, 4), np.uint64)])
        quotient_polys.append(q_x)
        set_numerators.append(n_x)
        set_r.append(r_polys)
    vs = _powers(v, len(sets))
    vs_w = np.stack([F.fr_to_limbs(s) for s in vs])
    h_x = h2.poly_lincomb(quotient_polys, vs_w)
    h1 = params.commit(h_x)
    if transcript is not None:
        transcript.write_ec_point(h1)
        u = transcript.squeeze_challenge()
    # linearisation at u
    l_parts, z_diffs = [], []
    for (points, polys), r_polys in zip(sets, set_r):
        diffs = [p for p in super_points if p not in points]
        z_i = evaluate_vanishing_polynomial(diffs, u)
        ys = _powers(y, len(polys))
        l_x = h2.poly_lincomb(polys, np.stack([F.fr_to_limbs(s) for s in ys]))
        r_at_u = sum(ys[j] * sum(c * pow(u, t, R) for t, c in enumerate(r_polys[j])) for j in range(len(polys))) % R
        l_x = h2.poly_op("sub", l_x, _poly_from_ints([r_at_u], n))
        l_parts.append(h2.poly_op("scale", l_x, s=F.fr_to_limbs(z_i)))
        z_diffs.append(z_i)
    l_x = h2.poly_lincomb(l_parts, vs_w)
    zt_eval = evaluate_vanishing_polynomial(super_points, u)
    l_x = h2.poly_op("axpy", l_x, h_x, F.fr_to_limbs((-zt_eval) % R))
    l_x = h2.poly_op("scale", l_x, s=F.fr_to_limbs(pow(z_diffs[0], -1, R)))
    must_be_zero = F.fr_from_limbs(h2.eval_polynomial(l_x, F.fr_to_limbs(u)))
    h2_x = np.concatenate([h2.kate_division(l_x, F.fr_to_limbs(u)), np.zeros((1, 4), np.uint64)])
    h2c = params.commit(h2_x)
    if transcript is not None:
        transcript.write_ec_point(h2c)
    return {"h1": h1, "h2": h2c, "h_x": h_x, "l_x": l_x, "h2_x": h2_x, "must_be_zero": must_be_zero, "sets": sets,
            "super_points": super_points, "numerators": set_numerators, "r_polys": set_r, "z_diffs": z_diffs, "one": one}
