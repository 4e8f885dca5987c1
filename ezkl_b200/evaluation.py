"""Host mirror of halo2's plonk/evaluation.rs: Expression -> straight-line program -> b200_quotient_eval (evaluate_h).

halo2's `Expression<F>` (Constant / Fixed / Advice / Instance query at a Rotation / Negated / Sum / Product / Scaled) is
lowered by its `GraphEvaluator` into calculations over value sources; this module does the same lowering into the
instruction format of include/ezkl_b200.h (b200_instr / b200_col_ref): common sub-expressions are shared, results live in
at most 32 slots (linear-scan allocation by last use), constants and column loads are referenced in place.
All queries index one flat list of extended-coset columns; rotations are in rows of the ORIGINAL domain (Rotation(r)),
the kernel scales them by 2^(ext_k - k).  Values are python ints mod r on the host side, wire limbs on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat
from . import fields as F

OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_DOUBLE, OP_SQUARE, OP_MOV = range(7)
MAX_SLOTS = 32
_SLOT, _CONST, _LOAD = 0, 1, 2


class Expression:
    """Immutable expression node; build with Constant / Query and python operators."""
    __slots__ = ("kind", "args")

    def __init__(self, kind, *args):
        self.kind, self.args = kind, args

    def __add__(self, o):
        return Expression("sum", self, _wrap(o))

    def __radd__(self, o):
        return Expression("sum", _wrap(o), self)

    def __sub__(self, o):
        return Expression("sub", self, _wrap(o))

    def __rsub__(self, o):
        return Expression("sub", _wrap(o), self)

    def __mul__(self, o):
        return Expression("product", self, _wrap(o))

    def __rmul__(self, o):
        return Expression("product", _wrap(o), self)

    def __neg__(self):
        return Expression("negated", self)

    def key(self):
        return (self.kind,) + tuple(a.key() if isinstance(a, Expression) else a for a in self.args)


def Constant(v: int) -> Expression:
    return Expression("constant", v % F.FR_MODULUS)


def Query(column: int, rotation: int = 0) -> Expression:
    """Value of extended-coset column `column` at Rotation(rotation) (Fixed / Advice / Instance query)."""
    return Expression("query", int(column), int(rotation))


def _wrap(o):
    return o if isinstance(o, Expression) else Constant(int(o))


class QuotientProgram:
    """Compiled program: loads [(column, rotation)], constants [int], instructions [(op, dst_slot, a, b)]."""

    def __init__(self, expr: Expression):
        self.loads, self.consts, self.instrs = [], [], []
        self._load_ix, self._const_ix = {}, {}
        nodes, order = {}, []            # key -> (kind, operand keys) in post-order, shared sub-expressions once

        def visit(e):
            k = e.key()
            if k in nodes:
                return k
            if e.kind in ("constant", "query"):
                nodes[k] = (e.kind, e.args)
            else:
                nodes[k] = (e.kind, tuple(visit(a) for a in e.args))
                order.append(k)
            return k

        root = visit(expr)
        if nodes[root][0] in ("constant", "query"):      # a bare leaf still needs one instruction to produce a row value
            order.append(("mov", root))
            nodes[("mov", root)] = ("mov", (root,))
            root = ("mov", root)
        last_use = {}
        for i, k in enumerate(order):
            for a in nodes[k][1]:
                last_use[a] = i
        last_use[root] = len(order)
        free, slot_of = list(range(MAX_SLOTS - 1, -1, -1)), {}

        def operand(k):
            kind, args = nodes[k]
            if kind == "constant":
                if args[0] not in self._const_ix:
                    self._const_ix[args[0]] = len(self.consts)
                    self.consts.append(args[0])
                return (_CONST << 30) | self._const_ix[args[0]]
            if kind == "query":
                if args not in self._load_ix:
                    self._load_ix[args] = len(self.loads)
                    self.loads.append(args)
                return (_LOAD << 30) | self._load_ix[args]
            return (_SLOT << 30) | slot_of[k]

        opmap = {"sum": OP_ADD, "sub": OP_SUB, "product": OP_MUL, "negated": OP_NEG, "mov": OP_MOV}
        for i, k in enumerate(order):
            kind, args = nodes[k]
            a = operand(args[0])
            b = operand(args[1]) if len(args) > 1 else 0
            op = opmap[kind]
            if kind == "product" and args[0] == args[1]:
                op, b = OP_SQUARE, 0
            for x in set(args):                        # operands whose last use is here free their slot before dst is chosen
                if x in slot_of and last_use.get(x) == i:
                    free.append(slot_of[x])
            if not free:
                raise nat.B200Error("QuotientProgram: more than %d live intermediates" % MAX_SLOTS)
            dst = free.pop()
            slot_of[k] = dst
            self.instrs.append((op, dst, a, b))

    def arrays(self):
        loads = np.array(self.loads, dtype=np.int64).reshape(-1, 2).astype(np.int32)
        consts = np.stack([F.fr_to_limbs(c) for c in self.consts]) if self.consts else np.zeros((0, 4), np.uint64)
        prog = np.array([[op | (dst << 8), a, b] for op, dst, a, b in self.instrs], dtype=np.uint32).reshape(-1, 3)
        return np.ascontiguousarray(loads), np.ascontiguousarray(consts), np.ascontiguousarray(prog)

    def evaluate_ints(self, column_values, idx: int, n_rows: int, rot_scale: int) -> int:
        """Reference semantics on python ints (used by the tests to cross-check the compiler itself)."""
        r = F.FR_MODULUS
        slots = [0] * MAX_SLOTS

        def src(s):
            kind, i = s >> 30, s & 0x3FFFFFFF
            if kind == _SLOT:
                return slots[i]
            if kind == _CONST:
                return self.consts[i]
            col, rot = self.loads[i]
            return column_values[col][(idx + rot * rot_scale) % n_rows]

        last = 0
        for op, dst, a, b in self.instrs:
            x = src(a)
            if op == OP_ADD:
                v = x + src(b)
            elif op == OP_SUB:
                v = x - src(b)
            elif op == OP_MUL:
                v = x * src(b)
            elif op == OP_NEG:
                v = -x
            elif op == OP_DOUBLE:
                v = 2 * x
            elif op == OP_SQUARE:
                v = x * x
            else:
                v = x
            slots[dst] = v % r
            last = dst
        return slots[last]


def evaluate_h(program: QuotientProgram, columns, k: int, ext_k: int) -> np.ndarray:
    """Host-buffer path: columns = list of [2^ext_k, 4] wire arrays -> [2^ext_k, 4]."""
    nat.ensure_init()
    cols = [nat.as_u64(c, 4) for c in columns]
    N = 1 << ext_k
    assert all(c.shape[0] == N for c in cols)
    loads, consts, prog = program.arrays()
    out = np.zeros((N, 4), np.uint64)
    nat.check(nat.lib().b200_quotient_eval(nat.ptr_array(cols) if cols else None, C.c_size_t(len(cols)), C.c_uint32(k), C.c_uint32(ext_k),
                                           loads.ctypes.data_as(C.c_void_p), C.c_size_t(loads.shape[0]), nat.ptr(consts) if consts.size else None,
                                           C.c_size_t(consts.shape[0]), prog.ctypes.data_as(C.c_void_p), C.c_size_t(prog.shape[0]), nat.ptr(out)))
    return out


def evaluate_h_device(program: QuotientProgram, columns, k: int, ext_k: int, out=None):
    """Device path: columns = list of torch int64 CUDA tensors [2^ext_k, 4]; enqueued on torch's current stream."""
    import torch
    from .device import _stream
    N = 1 << ext_k
    for c in columns:
        assert c.is_cuda and c.dtype == torch.int64 and c.is_contiguous() and c.shape == (N, 4)
    if out is None:
        out = torch.empty((N, 4), dtype=torch.int64, device="cuda")
    loads, consts, prog = program.arrays()
    ptrs = (C.c_void_p * max(1, len(columns)))(*[c.data_ptr() for c in columns])
    nat.check(nat.lib().b200_quotient_eval_dev(ptrs, C.c_size_t(len(columns)), C.c_uint32(k), C.c_uint32(ext_k), loads.ctypes.data_as(C.c_void_p),
                                               C.c_size_t(loads.shape[0]), nat.ptr(consts) if consts.size else None, C.c_size_t(consts.shape[0]),
                                               prog.ctypes.data_as(C.c_void_p), C.c_size_t(prog.shape[0]), nat.dev(out.data_ptr()), _stream()))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Permutation grand product and mv-lookup grand sum (halo2 plonk/permutation/prover.rs, plonk/mv_lookup/prover.rs; stage 3
# of create_proof): row-wise numerator / denominator programs on the Lagrange domain (the same
# interpreter with k == ext_k), one batch inversion, one running product / sum.  All on the device.
DELTA = pow(7, 1 << 28, F.FR_MODULUS)          # Fr::DELTA = GENERATOR^(2^S): coset separator of the permutation argument


def _omega_powers_column(k: int) -> np.ndarray:
    """Lagrange-domain column of omega^i (the identity polynomial's values), built on the host once per k."""
    w = pow(F.FR_ROOT_OF_UNITY, 1 << (F.FR_S - k), F.FR_MODULUS)
    out, cur = [], 1
    for _ in range(1 << k):
        out.append(F.fr_to_limbs(cur))
        cur = cur * w % F.FR_MODULUS
    return np.stack(out)


def permutation_product(values, sigmas, k: int, beta: int, gamma: int, delta_start: int = 1) -> np.ndarray:
    """z(X) in Lagrange form for one permutation chunk:  z[0] = 1,
        z[i+1] = z[i] * prod_j (v_j[i] + beta * delta_start * DELTA^j * omega^i + gamma) / (v_j[i] + beta * sigma_j[i] + gamma).
    values / sigmas: lists of [n,4] Lagrange columns (wire form)."""
    from . import halo2 as h2
    m = len(values)
    assert m == len(sigmas) and m > 0
    r = F.FR_MODULUS
    cols = list(values) + list(sigmas) + [_omega_powers_column(k)]
    X = Query(2 * m)
    num = den = None
    for j in range(m):
        d = delta_start * pow(DELTA, j, r) % r
        tn = Query(j) + X * Constant(beta * d % r) + Constant(gamma)
        td = Query(j) + Query(m + j) * Constant(beta) + Constant(gamma)
        num = tn if num is None else num * tn
        den = td if den is None else den * td
    numer = evaluate_h(QuotientProgram(num), cols, k, k)
    denom = h2.batch_invert(evaluate_h(QuotientProgram(den), cols, k, k))
    ratio = h2.poly_op("mul", numer, denom)
    return h2.prefix_scan(ratio, F.fr_to_limbs(1), True)


def lookup_grand_sum(inputs, table, multiplicities, k: int, beta: int) -> np.ndarray:
    """phi(X) in Lagrange form for a logUp / mv-lookup argument:  phi[0] = 0,
        phi[i+1] = phi[i] + sum_j 1 / (f_j[i] + beta) - m[i] / (t[i] + beta).
    inputs: list of compressed input-expression columns f_j; table: compressed table column t; multiplicities: m."""
    from . import halo2 as h2
    nin = len(inputs)
    cols = list(inputs) + [table]
    one = F.fr_to_limbs(1)
    dens = [h2.batch_invert(evaluate_h(QuotientProgram(Query(j) + Constant(beta)), cols, k, k)) for j in range(nin + 1)]
    acc = dens[0]
    for j in range(1, nin):
        acc = h2.poly_op("add", acc, dens[j])
    acc = h2.poly_op("sub", acc, h2.poly_op("mul", multiplicities, dens[nin]))
    return h2.prefix_scan(acc, np.zeros(4, np.uint64), False)
