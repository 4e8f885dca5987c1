"""ctypes front end of oracle/liboracle.so (the CPU restatement in bn254_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs;
never from ezkl_b200/.  Arrays are numpy uint64 in wire format: Fr -> [n,4], G1Affine -> [n,8].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "bn254_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _fr(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == 4
    return a


def host_threads() -> int:
    return len(os.sched_getaffinity(0))


# ---- constants ------------------------------------------------------------------------------------------
def omega(k: int) -> np.ndarray:
    out = np.zeros(4, np.uint64)
    lib().orc_fr_omega(C.c_uint32(k), _p(out))
    return out


def zeta() -> np.ndarray:
    out = np.zeros(4, np.uint64)
    lib().orc_fr_zeta(_p(out))
    return out


def fr_one() -> np.ndarray:
    return np.array([0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f], np.uint64)


# ---- field ----------------------------------------------------------------------------------------------
def field_op(field: str, op: str, a, b) -> np.ndarray:
    a, b = _fr(a), _fr(b)
    out = np.empty_like(a)
    lib().orc_field_op(C.c_int({"fr": 0, "fq": 1}[field]), C.c_int({"add": 0, "sub": 1, "mul": 2}[op]), _p(a), _p(b), _p(out),
                       C.c_size_t(a.size // 4))
    return out


def fr_inv(a) -> np.ndarray:
    a = _fr(a)
    out = np.empty_like(a)
    lib().orc_fr_inv(_p(a), _p(out), C.c_size_t(a.size // 4))
    return out


def fq_inv(a) -> np.ndarray:
    a = _fr(a)
    out = np.empty_like(a)
    lib().orc_fq_inv(_p(a), _p(out), C.c_size_t(a.size // 4))
    return out


def fr_from_mont(a) -> np.ndarray:
    a = _fr(a)
    out = np.empty_like(a)
    lib().orc_fr_from_mont(_p(a), _p(out), C.c_size_t(a.size // 4))
    return out


def fr_to_mont(a) -> np.ndarray:
    a = _fr(a)
    out = np.empty_like(a)
    lib().orc_fr_to_mont(_p(a), _p(out), C.c_size_t(a.size // 4))
    return out


def fq_to_mont(a) -> np.ndarray:
    a = _fr(a)
    out = np.empty_like(a)
    lib().orc_fq_to_mont(_p(a), _p(out), C.c_size_t(a.size // 4))
    return out


def fr_pow(a, e: int) -> np.ndarray:
    out = np.zeros(4, np.uint64)
    lib().orc_fr_pow(_p(_fr(a)), C.c_uint64(e), _p(out))
    return out


def poly_op(op: str, a, b=None, s=None, threads: int = 1) -> np.ndarray:
    a = _fr(a)
    out = np.empty_like(a)
    code = {"add": 0, "sub": 1, "mul": 2, "scale": 3, "axpy": 4}[op]
    bp = _p(_fr(b)) if b is not None else None
    sp = _p(_fr(s)) if s is not None else None
    lib().orc_poly_op(C.c_int(code), _p(a), bp, sp, _p(out), C.c_size_t(a.size // 4), C.c_int(threads))
    return out


# ---- G1 -------------------------------------------------------------------------------------------------
def g1_is_on_curve(p) -> bool:
    p = np.ascontiguousarray(p, np.uint64)
    return bool(lib().orc_g1_is_on_curve(_p(p)))


def g1_add_affine(a, b) -> np.ndarray:
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    out = np.empty_like(a)
    lib().orc_g1_add_affine(_p(a), _p(b), _p(out), C.c_size_t(a.size // 8))
    return out


def g1_jac_to_affine(p) -> np.ndarray:
    p = np.ascontiguousarray(p, np.uint64).reshape(-1, 12)
    out = np.empty((p.shape[0], 8), np.uint64)
    lib().orc_g1_jac_to_affine(_p(p), _p(out), C.c_size_t(p.shape[0]))
    return out


def g1_scalar_mul(bases, scalars) -> np.ndarray:
    bases = np.ascontiguousarray(bases, np.uint64)
    scalars = _fr(scalars)
    out = np.empty_like(bases)
    lib().orc_g1_scalar_mul(_p(bases), _p(scalars), _p(out), C.c_size_t(bases.size // 8))
    return out


def msm_naive(scalars, bases) -> np.ndarray:
    scalars, bases = _fr(scalars), np.ascontiguousarray(bases, np.uint64)
    out = np.zeros(8, np.uint64)
    lib().orc_msm_naive(_p(scalars), _p(bases), C.c_size_t(scalars.size // 4), _p(out))
    return out


def msm(scalars, bases, threads: int = 1) -> np.ndarray:
    """best_multiexp + to_affine -> one 64-byte affine point (uint64[8])."""
    scalars, bases = _fr(scalars), np.ascontiguousarray(bases, np.uint64)
    assert scalars.size // 4 == bases.size // 8
    out = np.zeros(8, np.uint64)
    lib().orc_msm(_p(scalars), _p(bases), C.c_size_t(scalars.size // 4), C.c_int(threads), _p(out))
    return out


def gen_bases(n: int, seed: int = 0xE2C1B200, threads: int | None = None) -> np.ndarray:
    out = np.zeros((n, 8), np.uint64)
    lib().orc_gen_bases(_p(out), C.c_size_t(n), C.c_uint64(seed), C.c_int(threads or host_threads()))
    return out


def gen_scalars(n: int, seed: int = 0xE2C1B200) -> np.ndarray:
    out = np.zeros((n, 4), np.uint64)
    lib().orc_gen_scalars(_p(out), C.c_size_t(n), C.c_uint64(seed))
    return out


# ---- NTT / domain ---------------------------------------------------------------------------------------
def best_fft(a, log_n: int, omega_, threads: int = 1) -> np.ndarray:
    a = _fr(a).copy()
    assert a.shape[0] == 1 << log_n
    lib().orc_best_fft(_p(a), C.c_uint32(log_n), _p(_fr(omega_)), C.c_int(threads))
    return a


def lagrange_to_coeff(a, k: int, threads: int = 1) -> np.ndarray:
    a = _fr(a).copy()
    lib().orc_lagrange_to_coeff(_p(a), C.c_uint32(k), C.c_int(threads))
    return a


def coeff_to_lagrange(a, k: int, threads: int = 1) -> np.ndarray:
    a = _fr(a).copy()
    lib().orc_coeff_to_lagrange(_p(a), C.c_uint32(k), C.c_int(threads))
    return a


def coeff_to_extended(coeffs, ext_k: int, threads: int = 1) -> np.ndarray:
    coeffs = _fr(coeffs)
    out = np.zeros((1 << ext_k, 4), np.uint64)
    lib().orc_coeff_to_extended(_p(coeffs), C.c_size_t(coeffs.shape[0]), C.c_uint32(ext_k), _p(out), C.c_int(threads))
    return out


def extended_to_coeff(a, ext_k: int, threads: int = 1) -> np.ndarray:
    a = _fr(a).copy()
    lib().orc_extended_to_coeff(_p(a), C.c_uint32(ext_k), C.c_int(threads))
    return a


def divide_by_vanishing(a, k: int, ext_k: int) -> np.ndarray:
    a = _fr(a).copy()
    lib().orc_divide_by_vanishing(_p(a), C.c_uint32(k), C.c_uint32(ext_k))
    return a


def eval_polynomial(coeffs, x) -> np.ndarray:
    coeffs = _fr(coeffs)
    out = np.zeros(4, np.uint64)
    lib().orc_eval_polynomial(_p(coeffs), C.c_size_t(coeffs.shape[0]), _p(_fr(x)), _p(out))
    return out


def kate_division(a, b) -> np.ndarray:
    a = _fr(a)
    q = np.zeros((a.shape[0] - 1, 4), np.uint64)
    lib().orc_kate_division(_p(a), C.c_size_t(a.shape[0]), _p(_fr(b)), _p(q))
    return q


def batch_invert(a) -> np.ndarray:
    a = _fr(a).copy()
    lib().orc_batch_invert(_p(a), C.c_size_t(a.shape[0]))
    return a


def prefix_scan(a, init, product: bool) -> np.ndarray:
    a = _fr(a)
    out = np.empty_like(a)
    lib().orc_prefix_scan(C.c_int(1 if product else 0), _p(a), C.c_size_t(a.shape[0]), _p(_fr(init)), _p(out))
    return out


def quotient_eval(columns, k: int, ext_k: int, loads, consts, prog, threads: int = 1) -> np.ndarray:
    """columns: list of [2^ext_k,4] arrays; loads int32/uint32 [n,2] (column, rotation); consts [m,4]; prog uint32 [p,4] (op_dst, a, b, c)."""
    cols = [_fr(c) for c in columns]
    N = 1 << ext_k
    assert all(c.shape[0] == N for c in cols)
    arr = (C.c_void_p * max(1, len(cols)))(*[c.ctypes.data for c in cols])
    loads = np.ascontiguousarray(np.asarray(loads, dtype=np.int64).astype(np.int32).view(np.uint32).reshape(-1, 2))
    consts = np.ascontiguousarray(np.asarray(consts, dtype=np.uint64).reshape(-1, 4))
    prog = np.ascontiguousarray(np.asarray(prog, dtype=np.uint32).reshape(-1, 4))
    out = np.zeros((N, 4), np.uint64)
    lib().orc_quotient_eval(arr, C.c_uint32(k), C.c_uint32(ext_k), loads.ctypes.data_as(C.c_void_p), consts.ctypes.data_as(C.c_void_p),
                            prog.ctypes.data_as(C.c_void_p), C.c_size_t(prog.shape[0]), _p(out), C.c_int(threads))
    return out
