"""TEST-ONLY: the optimal ate pairing on BN254 (alt_bn128) in plain Python integers, so that the not-gpu tier can run the KZG / SHPLONK
final check the reference's verifier performs — e(L, [1]_2) == e(pi, [s]_2) — on the reference's OWN SRS fixture (whose trapdoor nobody
knows), instead of only on trapdoor SRSs.  Published algorithm, restated for the tests: Fp12 = Fp[w] / (w^12 - 18 w^6 + 82) (from u^2 = -1,
w^6 = 9 + u), the D-type twist (x, y) -> (x w^2, y w^3), Miller loop over 6x + 2 with the two Frobenius corrections, final exponentiation
by (p^12 - 1) / r.  Pinned by bilinearity and by the fixture relation e(g[i+1], g2) == e(g[i], s_g2) (tests/test_pairing_srs.py).
Slow by design (about a second per pairing); nothing in the product imports it."""

P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R_ORDER = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
ATE_LOOP_COUNT = 29793968203157093288          # 6x + 2, x = 4965661367192848881
LOG_ATE_LOOP_COUNT = 63
FQ12_MODULUS = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]


class FQ12:
    """Element of Fp[w] / (w^12 - 18 w^6 + 82): 12 coefficients, low degree first."""
    __slots__ = ("c",)

    def __init__(self, coeffs):
        self.c = [int(x) % P for x in coeffs]

    @classmethod
    def one(cls):
        return cls([1] + [0] * 11)

    @classmethod
    def zero(cls):
        return cls([0] * 12)

    def __add__(self, o):
        return FQ12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return FQ12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return FQ12([-a for a in self.c])

    def __eq__(self, o):
        return self.c == o.c

    def scale(self, k: int):
        return FQ12([a * k for a in self.c])

    def __mul__(self, o):
        t = [0] * 23
        for i, a in enumerate(self.c):
            if a:
                for j, b in enumerate(o.c):
                    t[i + j] += a * b
        for d in range(22, 11, -1):                # w^12 = 18 w^6 - 82
            top = t[d]
            if top:
                t[d - 6] += 18 * top
                t[d - 12] -= 82 * top
        return FQ12(t[:12])

    def inv(self):
        """Extended Euclid on polynomials over Fp against the field modulus."""
        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = self.c + [0], [x % P for x in FQ12_MODULUS] + [1]
        deg = lambda p: max([i for i, x in enumerate(p) if x] + [0])
        while deg(low):
            r = _poly_rounded_div(high, low)
            r += [0] * (13 - len(r))
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] -= lm[i] * r[j]
                    new[i + j] -= low[i] * r[j]
            nm = [x % P for x in nm]
            new = [x % P for x in new]
            lm, low, hm, high = nm, new, lm, low
        k = pow(low[0], -1, P)
        return FQ12([x * k for x in lm[:12]])

    def __truediv__(self, o):
        return self * o.inv()

    def __pow__(self, e: int):
        result, base = FQ12.one(), self
        while e:
            if e & 1:
                result = result * base
            base = base * base
            e >>= 1
        return result


def _poly_rounded_div(a, b):
    dega = max([i for i, x in enumerate(a) if x] + [0])
    degb = max([i for i, x in enumerate(b) if x] + [0])
    temp = list(a)
    o = [0] * len(a)
    inv_lead = pow(b[degb], -1, P)
    for i in range(dega - degb, -1, -1):
        q = temp[degb + i] * inv_lead % P
        o[i] = (o[i] + q) % P
        for c in range(degb + 1):
            temp[c + i] = (temp[c + i] - q * b[c]) % P
    return [x % P for x in o[: max([i for i, x in enumerate(o) if x] + [0]) + 1]]


W = FQ12([0, 1] + [0] * 10)
W2, W3 = W * W, W * W * W


def cast_g1(pt):
    """(x, y) in Fp^2 -> point with Fp12 coordinates."""
    x, y = pt
    return (FQ12([x] + [0] * 11), FQ12([y] + [0] * 11))


def twist(pt):
    """G2 point ((x0, x1), (y0, y1)) on y^2 = x^3 + 3 / (9 + u) -> point on y^2 = x^3 + 3 over Fp12 (u -> w^6 - 9)."""
    (x0, x1), (y0, y1) = pt
    nx = FQ12([x0 - 9 * x1] + [0] * 5 + [x1] + [0] * 5)
    ny = FQ12([y0 - 9 * y1] + [0] * 5 + [y1] + [0] * 5)
    return (nx * W2, ny * W3)


def _double(pt):
    x, y = pt
    m = (x * x).scale(3) / y.scale(2)
    nx = m * m - x.scale(2)
    return (nx, m * (x - nx) - y)


def _add(p1, p2):
    if p1 is None or p2 is None:
        return p1 if p2 is None else p2
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        return _double(p1) if y1 == y2 else None
    m = (y2 - y1) / (x2 - x1)
    nx = m * m - x1 - x2
    return (nx, m * (x1 - nx) - y1)


def _line(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if not x1 == x2:
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1).scale(3) / y1.scale(2)
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q, p):
    if q is None or p is None:
        return FQ12.one()
    r, f = q, FQ12.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f * f * _line(r, r, p)
        r = _double(r)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _line(r, q, p)
            r = _add(r, q)
    q1 = (q[0] ** P, q[1] ** P)
    nq2 = (q1[0] ** P, -(q1[1] ** P))
    f = f * _line(r, q1, p)
    r = _add(r, q1)
    f = f * _line(r, nq2, p)
    return f ** ((P ** 12 - 1) // R_ORDER)


def pairing(q_g2, p_g1):
    """e(P, Q) for P = (x, y) in G1 (ints) and Q = ((x0, x1), (y0, y1)) in G2; None = the identity."""
    if q_g2 is None or p_g1 is None:
        return FQ12.one()
    return miller_loop(twist(q_g2), cast_g1(p_g1))


def g2_on_curve(pt) -> bool:
    x, y = twist(pt)
    return y * y - x * x * x == FQ12([3] + [0] * 11)
