"""evaluate_h: the expression compiler (host logic, CPU) and the device interpreter against the oracle's interpreter (GPU)."""
import random

import numpy as np
import pytest

from ezkl_b200 import evaluation as ev
from ezkl_b200 import fields as F
from oracle import oracle as orc
from oracle import pyref
from tests import helpers as H

R = pyref.R


def random_expr(rng, ncols, depth):
    if depth == 0 or rng.random() < 0.15:
        if rng.random() < 0.3:
            return ev.Constant(rng.randrange(R))
        return ev.Query(rng.randrange(ncols), rng.choice([0, 0, 1, -1, 2, -3]))
    c = rng.random()
    a = random_expr(rng, ncols, depth - 1)
    if c < 0.15:
        return -a
    if c < 0.25:
        return a * a
    b = random_expr(rng, ncols, depth - 1)
    return a + b if c < 0.5 else (a - b if c < 0.65 else a * b)


def eval_tree(e, cols, idx, n, scale):
    if e.kind == "constant":
        return e.args[0]
    if e.kind == "query":
        return cols[e.args[0]][(idx + e.args[1] * scale) % n]
    v = [eval_tree(a, cols, idx, n, scale) for a in e.args]
    return {"sum": lambda: v[0] + v[1], "sub": lambda: v[0] - v[1], "product": lambda: v[0] * v[1], "negated": lambda: -v[0]}[e.kind]() % R


def test_compiler_matches_tree_semantics_and_oracle():
    rng = random.Random(2)
    k, ext_k, ncols = 3, 5, 4
    N, scale = 1 << ext_k, 1 << (ext_k - k)
    cols_int = [[rng.randrange(R) for _ in range(N)] for _ in range(ncols)]
    cols = [H.fr_array(c) for c in cols_int]
    for trial in range(25):
        e = random_expr(rng, ncols, 5)
        prog = ev.QuotientProgram(e)
        assert all((dst & 0xFFFF) < ev.MAX_SLOTS for _, dst, *_ in prog.instrs)
        loads, consts, instrs = prog.arrays()
        got = H.fr_list(orc.quotient_eval(cols, k, ext_k, loads, consts, instrs, threads=2))
        for idx in (0, 1, N - 1, rng.randrange(N)):
            exp = eval_tree(e, cols_int, idx, N, scale)
            assert prog.evaluate_ints(cols_int, idx, N, scale) == exp
            assert got[idx] == exp
    # bare leaves and shared sub-expressions
    q = ev.Query(1, 1)
    for e in (q, ev.Constant(5), (q + 1) * (q + 1) - (q + 1)):
        prog = ev.QuotientProgram(e)
        loads, consts, instrs = prog.arrays()
        got = H.fr_list(orc.quotient_eval(cols, k, ext_k, loads, consts, instrs))
        assert got[3] == eval_tree(e, cols_int, 3, N, scale)
    assert len(ev.QuotientProgram((q + 1) * (q + 1) - (q + 1)).instrs) == 3      # (q+1) computed once, squared, subtracted
    # a Horner chain (GraphEvaluator's Horner calculation) lowers to one multiply-add per part, none of them stored
    y, h = ev.Constant(77), ev.Query(0)
    for t in range(40):
        h = h * y + ev.Query(t % ncols, t % 3 - 1)
    prog = ev.QuotientProgram(h)
    assert len(prog.instrs) == 40 and all(op == ev.OP_MULADD and dst & ev.NOSTORE for op, dst, *_ in prog.instrs)
    loads, consts, instrs = prog.arrays()
    assert H.fr_list(orc.quotient_eval(cols, k, ext_k, loads, consts, instrs))[5] == eval_tree(h, cols_int, 5, N, scale)
    # shared DAGs cost O(nodes): 200 repeated squarings are 200 instructions (2^200 leaves as a tree)
    e = ev.Query(2)
    for _ in range(200):
        e = e * e
    assert len(ev.QuotientProgram(e).instrs) == 200
    # more than 32 live intermediates: 100 independent products kept alive until a final sum tree
    parts = [ev.Query(i % ncols, i % 5 - 2) * ev.Query((i + 1) % ncols, i % 3) - ev.Constant(i) for i in range(100)]
    while len(parts) > 1:
        parts = [parts[i] * parts[i + 1] if i + 1 < len(parts) else parts[i] for i in range(0, len(parts), 2)]
    prog = ev.QuotientProgram(parts[0])
    loads, consts, instrs = prog.arrays()
    assert H.fr_list(orc.quotient_eval(cols, k, ext_k, loads, consts, instrs))[7] == eval_tree(parts[0], cols_int, 7, N, scale)


@pytest.mark.gpu
def test_device_interpreter_vs_oracle():
    from ezkl_b200 import _native as nat
    nat.init(-1)
    rng = random.Random(4)
    for k, ext_k, ncols in ((4, 6, 3), (10, 13, 8), (13, 15, 12)):
        N = 1 << ext_k
        cols = [orc.gen_scalars(N, seed=100 * k + i) for i in range(ncols)]
        for trial in range(4):
            prog = ev.QuotientProgram(random_expr(rng, ncols, 6 + trial))
            loads, consts, instrs = prog.arrays()
            got = ev.evaluate_h(prog, cols, k, ext_k)
            assert np.array_equal(got, orc.quotient_eval(cols, k, ext_k, loads, consts, instrs, threads=orc.host_threads())), (k, trial)


@pytest.mark.gpu
def test_fused_evaluate_h_entry_equals_the_composition():
    """b200_evaluate_h (coefficient columns in, cosets built on the device, optional divide + extended_to_coeff) against the same steps
    done call by call: mixed column kinds (coefficient columns of two lengths, extended columns), ragged grouping, both finish modes."""
    from ezkl_b200 import _native as nat
    from ezkl_b200 import halo2 as h2
    nat.init(-1)
    rng = random.Random(31)
    k = 9
    n = 1 << k
    dom = h2.EvaluationDomain(5, k)
    N = 1 << dom.extended_k
    coeffs = [orc.gen_scalars(n, seed=700 + i) for i in range(5)]
    short = orc.gen_scalars(n // 2, seed=710)                           # a lower-degree polynomial: shorter coefficient vector
    ext_cols = [orc.gen_scalars(N, seed=720 + i) for i in range(2)]
    polys = [coeffs[0], ext_cols[0], coeffs[1], coeffs[2], short, ext_cols[1], coeffs[3], coeffs[4]]
    short_padded = np.concatenate([short, np.zeros((n - n // 2, 4), np.uint64)])
    cosets = [dom.coeff_to_extended(p) if p.shape[0] == n else (dom.coeff_to_extended(short_padded) if p.shape[0] == n // 2 else p) for p in polys]
    prog = ev.QuotientProgram(random_expr(rng, len(polys), 7) + ev.Query(4, 1) * ev.Query(5, -2))
    num = ev.evaluate_h(prog, cosets, k, dom.extended_k)
    assert np.array_equal(ev.evaluate_h_from_polys(prog, polys, dom), num)
    want = dom.extended_to_coeff(dom.divide_by_vanishing_poly(num))
    got = ev.evaluate_h_from_polys(prog, polys, dom, finish=True)
    assert np.array_equal(got[: want.shape[0]], want)


@pytest.mark.gpu
def test_quotient_pipeline_on_a_satisfied_gate():
    """End-to-end property on a tiny PLONK-style circuit (q_m*a*b + q_c - c = 0 on every row, plus a rotation term
    a(wX) - d = 0): the GPU pipeline iNTT -> coset NTT -> evaluate_h -> divide_by_vanishing -> extended_to_coeff must give
    h with  numerator(x) == h(x) * (x^n - 1)  at a random point, and h of degree < n*(deg-1)."""
    from ezkl_b200 import _native as nat
    from ezkl_b200 import halo2 as h2
    nat.init(-1)
    rng = random.Random(9)
    k = 8
    n = 1 << k
    dom = h2.EvaluationDomain(4, k)          # degree-3 gate -> quotient_poly_degree 3 -> extended_k = k + 2
    ext_k = dom.extended_k
    a = [rng.randrange(R) for _ in range(n)]
    b = [rng.randrange(R) for _ in range(n)]
    qm = [rng.randrange(R) for _ in range(n)]
    qc = [rng.randrange(R) for _ in range(n)]
    c = [(qm[i] * a[i] * b[i] + qc[i]) % R for i in range(n)]
    d = [a[(i + 1) % n] for i in range(n)]
    y = rng.randrange(R)
    lag = [H.fr_array(v) for v in (a, b, c, d, qm, qc)]
    coeffs = dom.lagrange_to_coeff_batch(lag)
    cosets = dom.coeff_to_extended_batch(coeffs)
    A, B, Cc, D, QM, QC = (ev.Query(i) for i in range(6))
    gate1 = QM * A * B + QC - Cc
    gate2 = ev.Query(0, 1) - D
    expr = gate1 * ev.Constant(y) + gate2               # fold with y like evaluate_h does
    prog = ev.QuotientProgram(expr)
    num_ext = ev.evaluate_h(prog, cosets, k, ext_k)
    h_ext = dom.divide_by_vanishing_poly(num_ext)
    h = dom.extended_to_coeff(h_ext)
    assert h.shape[0] == n * 3
    assert not h[2 * n:].any()                          # deg(numerator) <= 3(n-1)  =>  deg(h) < 2n
    x = rng.randrange(R)
    xv = H.fr_wire(x)
    ev_at = [H.fr_unwire(h2.eval_polynomial(p, xv)) for p in coeffs]
    a_rot = H.fr_unwire(h2.eval_polynomial(coeffs[0], H.fr_wire(x * pyref.omega_for(k) % R)))
    numerator = ((ev_at[4] * ev_at[0] * ev_at[1] + ev_at[5] - ev_at[2]) * y + (a_rot - ev_at[3])) % R
    hx = H.fr_unwire(h2.eval_polynomial(h, xv))
    assert numerator == hx * (pow(x, n, R) - 1) % R


@pytest.mark.gpu
def test_permutation_product_and_lookup_sum():
    """Grand product z(X) of a real permutation (copy constraints satisfied) closes to 1; logUp grand sum closes to 0;
    both match a python-int restatement row by row."""
    from ezkl_b200 import _native as nat
    nat.init(-1)
    rng = random.Random(12)
    k, m = 7, 3
    n = 1 << k
    w = pyref.omega_for(k)
    # a random permutation over the m*n cells, values constant on its cycles => the argument is satisfied
    cells = [(j, i) for j in range(m) for i in range(n)]
    perm = cells[:]
    rng.shuffle(perm)
    sigma_of = dict(zip(cells, perm))
    val, seen = {}, set()
    for c in cells:
        if c in seen:
            continue
        v, cur = rng.randrange(R), c
        while cur not in seen:
            seen.add(cur)
            val[cur] = v
            cur = sigma_of[cur]
    label = lambda j, i: pow(ev.DELTA, j, R) * pow(w, i, R) % R
    values = [[val[(j, i)] for i in range(n)] for j in range(m)]
    sigmas = [[label(*sigma_of[(j, i)]) for i in range(n)] for j in range(m)]
    beta, gamma = rng.randrange(R), rng.randrange(R)
    z_w, last_z = ev.permutation_product([H.fr_array(v) for v in values], [H.fr_array(s) for s in sigmas], k, beta, gamma)
    z = H.fr_list(z_w)
    assert last_z == z[n - 1]
    exp, acc = [], 1
    for i in range(n):
        exp.append(acc)
        num = den = 1
        for j in range(m):
            num = num * (values[j][i] + beta * label(j, i) + gamma) % R
            den = den * (values[j][i] + beta * sigmas[j][i] + gamma) % R
        acc = acc * num * pow(den, -1, R) % R
    assert z == exp
    assert acc == 1                      # full product over the domain is 1 for a satisfied permutation
    # logUp: every input value appears in the table; multiplicities count occurrences
    table = [rng.randrange(R) for _ in range(n)]
    ins = [[table[rng.randrange(n)] for _ in range(n)] for _ in range(2)]
    mult = [0] * n
    pos = {}
    for i, t in enumerate(table):
        pos.setdefault(t, i)
    for col in ins:
        for v in col:
            mult[pos[v]] += 1
    phi = H.fr_list(ev.lookup_grand_sum([H.fr_array(c) for c in ins], H.fr_array(table), H.fr_array(mult), k, beta))
    exp, acc = [], 0
    for i in range(n):
        exp.append(acc)
        acc = (acc + sum(pow(c[i] + beta, -1, R) for c in ins) - mult[i] * pow(table[i] + beta, -1, R)) % R
    assert phi == exp and acc == 0


def test_shplonk_host_logic_on_the_cpu_backend(monkeypatch):
    """The SHPLONK mirror's host logic (rotation sets, interpolants, ASCENDING powers of y and v, linearisation at u) with the
    polynomial steps executed by the CPU oracle (tests/cpu_backend.py): every rotation set's quotient divides exactly, L(u) == 0, and
    VerifierSHPLONK's equation holds in scalar form at the trapdoor.  With DESCENDING powers (the GWC rule) the same equation fails —
    the ordering is what this pins."""
    from ezkl_b200 import multiopen as mo
    from tests import cpu_backend as cb
    monkeypatch.setattr(mo, "h2", cb)
    rng = random.Random(43)
    k = 6
    n = 1 << k
    s = rng.randrange(2, R)
    params = cb.TrapdoorParams(k, s)
    w = pyref.omega_for(k)
    x = rng.randrange(R)
    polys = [orc.gen_scalars(n, seed=700 + i) for i in range(5)]
    pts_a, pts_b, pts_c = [x, x * w % R], [x], [x, x * w % R, x * pow(w, -1, R) % R]
    queries = []
    for p, pts in ((polys[0], pts_a), (polys[1], pts_b), (polys[2], pts_a), (polys[3], pts_c), (polys[4], pts_b)):
        queries += [mo.ProverQuery(pt, p) for pt in pts]
    y, v, u = (rng.randrange(R) for _ in range(3))
    prf = mo.create_proof(params, queries, y, v, u)
    assert prf["must_be_zero"] == 0
    assert [len(pl) for _, pl in prf["sets"]] == [2, 2, 1] and len(prf["super_points"]) == 3
    sv = H.fr_wire(s)
    evs = lambda poly: H.fr_unwire(cb.eval_polynomial(poly, sv))
    # the set quotients divide exactly: q_i(z) * Z_i(z) == N_i(z) at a random z, and h = sum v^i q_i
    z = rng.randrange(R)
    ez = lambda poly: H.fr_unwire(cb.eval_polynomial(poly, H.fr_wire(z)))
    acc = 0
    for i, ((points, _), num) in enumerate(zip(prf["sets"], prf["numerators"])):
        acc = (acc + pow(v, i, R) * ez(num) * pow(mo.evaluate_vanishing_polynomial(points, z), -1, R)) % R
    assert ez(prf["h_x"]) == acc and ez(prf["h2_x"]) * (z - u) % R == ez(prf["l_x"])

    def verifier_rhs(y_pows, v_pows):
        zt = mo.evaluate_vanishing_polynomial(prf["super_points"], u)
        z0_inv = pow(prf["z_diffs"][0], -1, R)
        outer = 0
        for i, ((points, polys_i), r_polys, zd) in enumerate(zip(prf["sets"], prf["r_polys"], prf["z_diffs"])):
            inner = 0
            for j, (pl, rp) in enumerate(zip(polys_i, r_polys)):
                r_u = sum(c * pow(u, t, R) for t, c in enumerate(rp)) % R
                inner = (inner + y_pows(j, len(polys_i)) * (evs(pl) - r_u)) % R
            outer = (outer + v_pows(i, len(prf["sets"])) * zd % R * z0_inv % R * inner) % R
        return (outer - zt * z0_inv % R * evs(prf["h_x"])) % R

    lhs = evs(prf["h2_x"]) * (s - u) % R
    assert lhs == verifier_rhs(lambda j, m: pow(y, j, R), lambda i, m: pow(v, i, R))                      # ascending: SHPLONK
    assert lhs != verifier_rhs(lambda j, m: pow(y, m - 1 - j, R), lambda i, m: pow(v, m - 1 - i, R))      # descending: GWC's fold
    # the commitments are the trapdoor evaluations of the polynomials behind them
    G = np.array(list(H.fq_wire(1)) + list(H.fq_wire(2)), np.uint64).reshape(1, 8)
    assert np.array_equal(prf["h1"][:8], orc.g1_scalar_mul(G, H.fr_array([evs(prf["h_x"])]))[0])
    assert np.array_equal(prf["h2"][:8], orc.g1_scalar_mul(G, H.fr_array([evs(prf["h2_x"])]))[0])


@pytest.mark.gpu
def test_shplonk_multiopen_flow_with_known_trapdoor():
    """ProverSHPLONK mirror (ezkl_b200/multiopen.py) composed from the device primitives, on an SRS whose trapdoor s is known:
    the quotient of every rotation set divides exactly, L(u) == 0, and both KZG relations hold in the exponent
    (commit(h2) * (s - u) == commit(L); h(z) == sum_i v^i N_i(z) / Z_i(z) at a random z) — no pairing needed."""
    from ezkl_b200 import _native as nat
    from ezkl_b200 import halo2 as h2
    from ezkl_b200 import multiopen as mo
    nat.init(-1)
    rng = random.Random(41)
    k = 8
    n = 1 << k
    s = rng.randrange(2, R)
    params = h2.ParamsKZG.setup(k, s)
    w = pyref.omega_for(k)
    x = rng.randrange(R)
    polys = [orc.gen_scalars(n, seed=500 + i) for i in range(5)]
    pts_a, pts_b, pts_c = [x, x * w % R], [x], [x, x * w % R, x * pow(w, -1, R) % R]
    queries = []
    for p, pts in ((polys[0], pts_a), (polys[1], pts_b), (polys[2], pts_a), (polys[3], pts_c), (polys[4], pts_b)):
        queries += [mo.ProverQuery(pt, p) for pt in pts]
    y, v, u = (rng.randrange(R) for _ in range(3))
    prf = mo.create_proof(params, queries, y, v, u)
    assert prf["must_be_zero"] == 0
    assert len(prf["sets"]) == 3 and len(prf["super_points"]) == 3
    z = rng.randrange(R)
    zv = H.fr_wire(z)
    ev = lambda poly: H.fr_unwire(h2.eval_polynomial(poly, zv))
    vs = [pow(v, i, R) for i in range(len(prf["sets"]))]                      # ascending: .zip(powers(v)), as SHPLONK combines
    rhs = 0
    for (points, _), num, vi in zip(prf["sets"], prf["numerators"], vs):
        rhs = (rhs + vi * ev(num) * pow(mo.evaluate_vanishing_polynomial(points, z), -1, R)) % R
    assert ev(prf["h_x"]) == rhs
    assert ev(prf["h2_x"]) * (z - u) % R == ev(prf["l_x"])
    # KZG in the exponent with the trapdoor
    G = np.array(list(H.fq_wire(1)) + list(H.fq_wire(2)), np.uint64).reshape(1, 8)
    sv = H.fr_wire(s)
    assert np.array_equal(prf["h1"][:8], orc.g1_scalar_mul(G, H.fr_array([H.fr_unwire(h2.eval_polynomial(prf["h_x"], sv))]))[0])
    lhs = orc.g1_scalar_mul(prf["h2"][:8].reshape(1, 8), H.fr_array([(s - u) % R]))
    assert np.array_equal(lhs[0], params.commit(prf["l_x"])[:8])
    # VerifierSHPLONK's equation restated in the exponent with the trapdoor (no pairing): with z_i = Z_{T \ S_i}(u) and r_ij the
    # interpolants, the verifier forms  sum_i v^i (z_i / z_0) sum_j y^j (C_ij - [r_ij(u)] G) - (Z_T(u) / z_0) h1 + u h2  and pairs it
    # against h2 * [s]:  h2(s) * (s - u) == that combination evaluated at s.  The powers of y and v ascend on both sides.
    zt = mo.evaluate_vanishing_polynomial(prf["super_points"], u)
    z0_inv = pow(prf["z_diffs"][0], -1, R)
    evs = lambda poly: H.fr_unwire(h2.eval_polynomial(poly, sv))
    outer = 0
    for i, ((points, polys_i), r_polys, zd) in enumerate(zip(prf["sets"], prf["r_polys"], prf["z_diffs"])):
        inner = 0
        for j, (pl, rp) in enumerate(zip(polys_i, r_polys)):
            r_u = sum(c * pow(u, t, R) for t, c in enumerate(rp)) % R
            inner = (inner + pow(y, j, R) * (evs(pl) - r_u)) % R
        outer = (outer + pow(v, i, R) * zd % R * z0_inv % R * inner) % R
    assert evs(prf["h2_x"]) * (s - u) % R == (outer - zt * z0_inv % R * evs(prf["h_x"])) % R
