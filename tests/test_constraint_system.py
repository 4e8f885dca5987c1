"""evaluate_h on an ezkl-shaped constraint system: BaseConfig custom gates (/root/reference/src/circuit/ops/chip.rs:362-424,
base.rs:28-66), a two-chunk permutation argument with blinded grand products chained through last_z, and one mv-lookup whose
multiplicities come from the device histogram.  A satisfying witness is built row by row; the folded numerator is evaluated on the
device from the compiled program and checked against (a) the oracle's interpreter, (b) the quotient identity
numerator(x) == h(x) * (x^n - 1) at a random point, and (c) the same identity FAILING once a single witness cell is tampered with."""
import random

import numpy as np
import pytest

from ezkl_b200 import evaluation as ev
from ezkl_b200 import fields as F
from oracle import oracle as orc
from oracle import pyref
from tests import helpers as H

R = pyref.R
# flat column indices
A0, A1, B0, B1, OUT = 0, 1, 2, 3, 4
SEL = {"ADD": 5, "MULT": 6, "DOTINIT": 7, "DOT": 8, "SUM": 9}
TABLE, SEL_L = 10, 11
SIG = [12, 13, 14, 15, 16]
Z = [17, 18]
M, PHI = 19, 20
L0, LLAST, LACT, XCOL = 21, 22, 23, 24
NCOLS = 25
BLIND = 5
CHUNK = 3


def build_witness(rng, k):
    """Rows 0 .. u-1 are active (u = n - BLIND - 1); the rest hold blinding values with every selector off."""
    n = 1 << k
    u = n - BLIND - 1
    col = [[0] * n for _ in range(NCOLS)]
    distinct = [rng.randrange(1 << 20) for _ in range(40)]
    for i in range(u):
        col[TABLE][i] = distinct[min(i, 39)]                      # ezkl pads a table with its last entry: duplicates from row 39 on
    perm_cols = [A0, A1, B0, B1, OUT]
    copies = []                                                   # ((column, row), (column, row)) pairs forced equal
    for i in range(n):
        for c in (A0, A1, B0, B1):
            col[c][i] = rng.randrange(R)
        if i >= u:
            col[OUT][i] = rng.randrange(R)
            continue
        if rng.random() < 0.5:                                    # lookup row: a1 must be a table value
            col[SEL_L][i] = 1
            col[A1][i] = col[TABLE][rng.randrange(u)]
        if i > 2 and rng.random() < 0.4:                          # copy constraint: an earlier output feeds this row
            j = rng.randrange(i)
            col[A0][i] = col[OUT][j]
            copies.append(((OUT, j), (A0, i)))
        if i > 2 and rng.random() < 0.3:
            j = rng.randrange(i)
            col[B0][i] = col[B1][j]
            copies.append(((B1, j), (B0, i)))
        op = rng.choice(["ADD", "MULT", "DOTINIT"] + (["DOT", "SUM"] if i > 0 else []))
        col[SEL[op]][i] = 1
        a0, a1, b0, b1 = col[A0][i], col[A1][i], col[B0][i], col[B1][i]
        prev = col[OUT][i - 1]
        col[OUT][i] = {"ADD": a0 + b0, "MULT": a0 * b0, "DOTINIT": a0 * b0 + a1 * b1, "DOT": prev + a0 * b0 + a1 * b1, "SUM": prev + b0 + b1}[op] % R
    # permutation: union the copy pairs into cycles (a cell joins at most one pair here, so cycles are 2-cycles or chains)
    w = pyref.omega_for(k)
    label = lambda ci, i: pow(ev.DELTA, ci, R) * pow(w, i, R) % R
    nxt = {}
    for x, y_ in copies:
        cx = (perm_cols.index(x[0]), x[1])
        cy = (perm_cols.index(y_[0]), y_[1])
        # splice two cycles: swap successors
        sx, sy = nxt.get(cx, cx), nxt.get(cy, cy)
        nxt[cx], nxt[cy] = sy, sx
    for ci in range(5):
        for i in range(n):
            t = nxt.get((ci, i), (ci, i))
            col[SIG[ci]][i] = label(*t)
    return col, u


def test_term_builders_shape():
    terms = ev.base_op_gates(SEL, [A0, A1], [B0, B1], OUT) + ev.permutation_terms([A0, A1, B0, B1, OUT], SIG, Z, L0, LLAST, LACT, XCOL, 3, 5, CHUNK, BLIND) + \
        ev.mv_lookup_terms([ev.Query(SEL_L) * ev.Query(A1) + (ev.Constant(1) - ev.Query(SEL_L)) * ev.Constant(7)], ev.Query(TABLE), M, PHI, L0, LLAST, LACT, 11)
    assert len(terms) == 5 + (2 + 1 + 2) + 3
    prog = ev.QuotientProgram(ev.fold_y(terms, 99))
    assert 60 <= len(prog.instrs) <= 200 and prog.n_slots <= 32
    rots = {r for _, r in prog.loads}
    assert rots == {0, 1, -1, -(BLIND + 1)}


def eval_expr(e, value_of):
    if e.kind == "constant":
        return e.args[0]
    if e.kind == "query":
        return value_of(e.args[0], e.args[1])
    v = [eval_expr(a, value_of) for a in e.args]
    return {"sum": lambda: v[0] + v[1], "sub": lambda: v[0] - v[1], "product": lambda: v[0] * v[1], "negated": lambda: -v[0]}[e.kind]() % R


def int_columns_of_a_satisfying_system(rng, k, beta, gamma):
    """Every column of the ezkl-shaped system as python ints on the base domain: the witness of build_witness, then m, Phi and the two
    chained, blinded grand products restated directly from their definitions (no library call)."""
    n = 1 << k
    col, u = build_witness(rng, k)
    t_default = col[TABLE][0]
    f = [(col[SEL_L][i] * col[A1][i] + (1 - col[SEL_L][i]) * t_default) % R for i in range(n)]
    first_row = {}
    for i in range(u):
        first_row.setdefault(col[TABLE][i], i)
    for v in f[:u]:
        col[M][first_row[v]] += 1
    inv = lambda v: pow(v % R, -1, R)
    phi = [0]
    for i in range(u):
        phi.append((phi[-1] + inv(f[i] + beta) - col[M][i] * inv(col[TABLE][i] + beta)) % R)
    assert phi[u] == 0                                           # the logUp argument closes over the active rows
    col[PHI] = phi + [rng.randrange(R) for _ in range(n - u - 1)]
    for i in range(u, n):
        col[M][i] = rng.randrange(R)
    w = pyref.omega_for(k)
    perm_cols = [A0, A1, B0, B1, OUT]
    last_z = 1
    for ci, zc in enumerate(Z):
        cs = perm_cols[ci * CHUNK:(ci + 1) * CHUNK]
        ss = SIG[ci * CHUNK:(ci + 1) * CHUNK]
        z = [last_z]
        for i in range(u):
            num = den = 1
            for j, (c, s_) in enumerate(zip(cs, ss)):
                num = num * (col[c][i] + beta * pow(ev.DELTA, ci * CHUNK + j, R) * pow(w, i, R) + gamma) % R
                den = den * (col[c][i] + beta * col[s_][i] + gamma) % R
            z.append(z[-1] * num % R * inv(den) % R)
        last_z = z[u]
        col[zc] = z + [rng.randrange(R) for _ in range(n - u - 1)]
    assert last_z == 1                                           # the permutation argument closes
    # l0, l_last, l_active and the identity polynomial X on the base domain
    for i in range(n):
        col[L0][i] = 1 if i == 0 else 0
        col[LLAST][i] = 1 if i == u else 0
        col[LACT][i] = 1 if i < u else 0
        col[XCOL][i] = pow(w, i, R)
    return col, u, t_default


def test_ezkl_shaped_system_vanishes_on_the_domain_and_detects_tampering():
    """CPU-only pin of the term builders and the compiler on the whole ezkl-shaped system (BaseConfig gates, two-chunk blinded
    permutation chained through last_z, one mv-lookup): with the grand products, the grand sum and the multiplicities restated in python
    ints, the folded numerator is 0 on EVERY row of the base domain (tree semantics, the compiled program's integer evaluator, and the
    oracle's interpreter agree), and a single tampered cell makes it non-zero on the rows that cell touches."""
    rng = random.Random(77)
    k = 5
    n = 1 << k
    beta, gamma, y = (rng.randrange(R) for _ in range(3))
    col, u, t_default = int_columns_of_a_satisfying_system(rng, k, beta, gamma)
    perm_cols = [A0, A1, B0, B1, OUT]
    lookup_in = ev.Query(SEL_L) * ev.Query(A1) + (ev.Constant(1) - ev.Query(SEL_L)) * ev.Constant(t_default)
    terms = ev.base_op_gates(SEL, [A0, A1], [B0, B1], OUT) + ev.permutation_terms(perm_cols, SIG, Z, L0, LLAST, LACT, XCOL, beta, gamma, CHUNK, BLIND) + \
        ev.mv_lookup_terms([lookup_in], ev.Query(TABLE), M, PHI, L0, LLAST, LACT, beta)
    expr = ev.fold_y(terms, y)
    prog = ev.QuotientProgram(expr)
    loads, consts, instrs = prog.arrays()
    # every single term vanishes on every row (so the fold does, whatever y is)
    for ti, t in enumerate(terms):
        for i in range(n):
            assert eval_expr(t, lambda c, rot: col[c][(i + rot) % n]) == 0, (ti, i)
    got = H.fr_list(orc.quotient_eval([H.fr_array(c) for c in col], k, k, loads, consts, instrs, threads=2))     # ext_k = k: rotations step by one row
    for i in range(n):
        assert prog.evaluate_ints(col, i, n, 1) == 0 and got[i] == 0, i
    # tamper with one active output cell: its own gate row, the row that accumulates from it, and the permutation rows break
    bad = [list(c) for c in col]
    bad[OUT][3] = (bad[OUT][3] + 1) % R
    nz = [i for i in range(n) if prog.evaluate_ints(bad, i, n, 1) != 0]
    assert 3 in nz and len(nz) >= 1
    assert H.fr_list(orc.quotient_eval([H.fr_array(c) for c in bad], k, k, loads, consts, instrs, threads=2))[3] == prog.evaluate_ints(bad, 3, n, 1) != 0
    # a wrong multiplicity breaks only the lookup's running-sum term
    bad = [list(c) for c in col]
    bad[M][0] = (bad[M][0] + 1) % R
    broken = [ti for ti, t in enumerate(terms) if any(eval_expr(t, lambda c, rot: bad[c][(i + rot) % n]) != 0 for i in range(n))]
    assert broken == [len(terms) - 1]


@pytest.mark.gpu
def test_ezkl_shaped_constraint_system_through_evaluate_h():
    from ezkl_b200 import _native as nat
    from ezkl_b200 import halo2 as h2
    nat.init(-1)
    rng = random.Random(2024)
    k = 7
    n = 1 << k
    col, u = build_witness(rng, k)
    beta, gamma, y = (rng.randrange(R) for _ in range(3))
    t_default = col[TABLE][0]
    f = [(col[SEL_L][i] * col[A1][i] + (1 - col[SEL_L][i]) * t_default) % R for i in range(n)]
    # stage 2: multiplicities on the device, against a dictionary restatement of the CPU prover's map
    m_dev = H.fr_list(ev.lookup_multiplicities(H.fr_array(col[TABLE][:u]), [H.fr_array(f[:u])], u))
    first_row, m_exp = {}, [0] * u
    for i in range(u):
        first_row.setdefault(col[TABLE][i], i)
    for v in f[:u]:
        m_exp[first_row[v]] += 1
    assert m_dev == m_exp and sum(m_exp) == u
    with pytest.raises(nat.B200Error):
        ev.lookup_multiplicities(H.fr_array(col[TABLE][:u]), [H.fr_array([R - 1])], 1)
    col[M][:u] = m_exp
    # stage 3: grand sum and the two chained grand products, blinded rows drawn here
    phi = H.fr_list(ev.lookup_grand_sum([H.fr_array(f)], H.fr_array(col[TABLE]), H.fr_array(col[M]), k, beta))
    assert phi[0] == 0 and phi[u] == 0                           # the argument closes on the active rows
    col[PHI] = phi[: u + 1] + [rng.randrange(R) for _ in range(n - u - 1)]
    for i in range(u, n):
        col[M][i] = rng.randrange(R)
    blinds = [[rng.randrange(R) for _ in range(BLIND)] for _ in range(2)]
    perm_cols = [A0, A1, B0, B1, OUT]
    zs = ev.permutation_products([H.fr_array(col[c]) for c in perm_cols], [H.fr_array(col[s_]) for s_ in SIG], k, beta, gamma, CHUNK, BLIND, blinds)
    assert len(zs) == 2
    for zi, zc in zip(Z, zs):
        col[zi] = H.fr_list(zc)
    assert col[Z[0]][0] == 1 and col[Z[1]][0] == col[Z[0]][u] and col[Z[1]][u] == 1     # chunk 1 starts at chunk 0's last_z; the product closes
    assert col[Z[0]][n - BLIND:] == blinds[0]
    # cosets of everything
    dom = h2.EvaluationDomain(5, k)
    ext_k = dom.extended_k
    assert ext_k == k + 2
    l0, l_last, l_active = dom.keygen_l_polys(BLIND)
    x_coeff = np.zeros((n, 4), np.uint64)
    x_coeff[1] = F.fr_to_limbs(1)
    witness_cols = [c for c in range(NCOLS) if c not in (L0, LLAST, LACT, XCOL)]
    coeffs = dict(zip(witness_cols, dom.lagrange_to_coeff_batch([H.fr_array(col[c]) for c in witness_cols])))
    coeffs[XCOL] = x_coeff
    cos = dict(zip(witness_cols + [XCOL], dom.coeff_to_extended_batch([coeffs[c] for c in witness_cols + [XCOL]])))
    cos[L0], cos[LLAST], cos[LACT] = l0, l_last, l_active
    cosets = [cos[c] for c in range(NCOLS)]
    # the folded numerator
    lookup_in = ev.Query(SEL_L) * ev.Query(A1) + (ev.Constant(1) - ev.Query(SEL_L)) * ev.Constant(t_default)
    terms = ev.base_op_gates(SEL, [A0, A1], [B0, B1], OUT) + ev.permutation_terms(perm_cols, SIG, Z, L0, LLAST, LACT, XCOL, beta, gamma, CHUNK, BLIND) + \
        ev.mv_lookup_terms([lookup_in], ev.Query(TABLE), M, PHI, L0, LLAST, LACT, beta)
    expr = ev.fold_y(terms, y)
    prog = ev.QuotientProgram(expr)
    loads, consts, instrs = prog.arrays()
    num = ev.evaluate_h(prog, cosets, k, ext_k)
    assert np.array_equal(num, orc.quotient_eval(cosets, k, ext_k, loads, consts, instrs, threads=orc.host_threads()))
    h = dom.extended_to_coeff(dom.divide_by_vanishing_poly(num))

    # quotient identity at a random point, the l-polynomials and X evaluated from their definitions
    def identity_holds(h_coeffs, coeff_map):
        x = rng.randrange(R)
        w = pyref.omega_for(k)
        lag = lambda row, pt: (pow(pt, n, R) - 1) * pow(w, row, R) % R * pow(n * (pt - pow(w, row, R)) % R, -1, R) % R
        cache = {}

        def value_of(c, rot):
            if (c, rot) not in cache:
                pt = x * pow(w, rot, R) % R
                if c == XCOL:
                    v = pt
                elif c == L0:
                    v = lag(0, pt)
                elif c == LLAST:
                    v = lag(u, pt)
                elif c == LACT:
                    v = (1 - lag(u, pt) - sum(lag(i, pt) for i in range(u + 1, n))) % R
                else:
                    v = H.fr_unwire(h2.eval_polynomial(coeff_map[c], H.fr_wire(pt)))
                cache[(c, rot)] = v
            return cache[(c, rot)]

        numerator = eval_expr(expr, value_of)
        hx = H.fr_unwire(h2.eval_polynomial(h_coeffs, H.fr_wire(x)))
        return numerator == hx * (pow(x, n, R) - 1) % R

    assert identity_holds(h, coeffs)
    # tamper with one active output cell: the gate on that row (and its copies) breaks, the numerator stops being divisible
    bad = list(col[OUT])
    bad[3] = (bad[3] + 1) % R
    bad_coeff = dom.lagrange_to_coeff(H.fr_array(bad))
    cosets_bad = list(cosets)
    cosets_bad[OUT] = dom.coeff_to_extended(bad_coeff)
    h_bad = dom.extended_to_coeff(dom.divide_by_vanishing_poly(ev.evaluate_h(prog, cosets_bad, k, ext_k)))
    coeffs_bad = dict(coeffs)
    coeffs_bad[OUT] = bad_coeff
    assert not identity_holds(h_bad, coeffs_bad)


@pytest.mark.gpu
def test_multiplicities_device_resident_and_duplicates():
    """b200_lookup_multiplicities_dev on resident columns: several input columns, heavy repetition, duplicate table rows."""
    import ctypes as C
    import torch
    from ezkl_b200 import _native as nat
    from ezkl_b200 import device as dev
    nat.init(-1)
    rng = random.Random(5)
    n, n_in = 1 << 12, 3
    vals = [rng.randrange(1 << 30) for _ in range(1000)]
    table = [vals[min(i, 999)] for i in range(n)]                 # rows 999.. repeat the last value
    ins = [[table[min(int(rng.expovariate(1 / 50.0)), n - 1)] for _ in range(n)] for _ in range(n_in)]      # skewed towards the first rows
    d_t = dev.from_host(H.fr_array(table))
    d_in = [dev.from_host(H.fr_array(c)) for c in ins]
    d_m = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    ptrs = (C.c_void_p * n_in)(*[t.data_ptr() for t in d_in])
    missing = C.c_uint64(123)
    nat.check(nat.lib().b200_lookup_multiplicities_dev(nat.dev(d_t.data_ptr()), C.c_size_t(n), ptrs, C.c_size_t(n_in), C.c_size_t(n), nat.dev(d_m.data_ptr()),
                                                        C.byref(missing), dev._stream()))
    first, exp = {}, [0] * n
    for i, v in enumerate(table):
        first.setdefault(v, i)
    for c in ins:
        for v in c:
            exp[first[v]] += 1
    assert missing.value == 0 and H.fr_list(dev.to_host(d_m)) == exp
