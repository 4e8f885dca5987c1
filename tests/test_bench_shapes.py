"""The exact shapes bench.py times, against the CPU oracle (GPU box, `-m gpu`): the 60-column commit batch at k = 17 with the
registered 16-bit window, a k = 20 commit batch, the 32-column coset NTT 2^17 -> 2^20, the 33-column evaluate_h group at 2^20 with
bench.py's own gate program, the evaluation batch, and the three-pass transforms at ext_k = 23 (inverse with the coset post-scale
and zero-padded coset forward).  BASELINE.md §4's parity gate on the timed inputs; bench.py repeats the same comparison on its own
buffers before timing."""
import ctypes as C

import numpy as np
import pytest

import bench
from ezkl_b200 import _native as nat
from ezkl_b200 import evaluation as ev
from ezkl_b200 import halo2 as h2
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
THREADS = orc.host_threads()


@pytest.fixture(scope="module", autouse=True)
def _init():
    nat.init(-1)
    yield


def test_commit_batch_60_columns_k17_window16():
    n = 1 << 17
    bases_np = orc.gen_bases(n, seed=1701, threads=THREADS)
    bases = h2.Bases(bases_np)
    info = bases.info()
    assert info["window_bits"] == 16 and info["windows"] == 16
    cols = [orc.gen_scalars(n, seed=1710 + i) for i in range(60)]
    cols[3][:] = 0                                                   # an all-zero column commits to the identity
    cols[5][1::2] = 0                                                # half zeros
    got = h2.best_multiexp_batch(cols, bases)                        # b200_msm_batch, host pointers, batch = 60
    for i, c in enumerate(cols):
        exp = orc.msm(c, bases_np, THREADS)
        if not exp.any():
            assert not got[i, 8:].any(), i
        else:
            assert np.array_equal(got[i, :8], exp), i
    bases.release()


def test_commit_batch_k20():
    n = 1 << 20
    bases_np = orc.gen_bases(n, seed=2001, threads=THREADS)
    bases = h2.Bases(bases_np)
    cols = [orc.gen_scalars(n, seed=2010 + i) for i in range(4)]
    got = h2.best_multiexp_batch(cols, bases)
    for i, c in enumerate(cols):
        assert np.array_equal(got[i, :8], orc.msm(c, bases_np, THREADS)), i
    bases.release()


def test_coset_ntt_32_columns_and_quotient_group_2p20():
    k, tr = 17, bench.TRACES["conv2d_mnist"]
    n = 1 << k
    dom = h2.EvaluationDomain((1 << tr["ext_bits"]) + 1, k)
    ext_k = dom.extended_k
    assert ext_k == 20
    m = bench.QUOTIENT_GROUP
    coeffs = [orc.gen_scalars(n, seed=3000 + i) for i in range(m)]
    exts = dom.coeff_to_extended_batch(coeffs)                       # 32 x (2^17 -> 2^20), zero-padded coset transform
    for i in (0, 1, 7, 31):
        assert np.array_equal(exts[i], orc.coeff_to_extended(coeffs[i], ext_k, THREADS)), i
    prog = bench.gate_program(m)                                     # the program bench.py times: 32 columns + the running sum
    hq = orc.gen_scalars(1 << ext_k, seed=3100)
    loads, consts, instrs = prog.arrays()
    got = ev.evaluate_h(prog, exts + [hq], k, ext_k)
    assert np.array_equal(got, orc.quotient_eval(exts + [hq], k, ext_k, loads, consts, instrs, threads=THREADS))
    xs = orc.gen_scalars(16, seed=3200)
    evs = h2.eval_polynomial_batch(coeffs[:16], xs)
    for i in range(16):
        assert np.array_equal(evs[i], orc.eval_polynomial(coeffs[i], xs[i])), i


def test_three_pass_transforms_ext_k23():
    k, ext_k = 20, 23
    dom = h2.EvaluationDomain(9, k)
    assert dom.extended_k == ext_k
    coeff = orc.gen_scalars(1 << k, seed=4000)
    ext = dom.coeff_to_extended(coeff)                               # zero-padded coset forward, 3 passes
    assert np.array_equal(ext, orc.coeff_to_extended(coeff, ext_k, THREADS))
    full = orc.gen_scalars(1 << ext_k, seed=4001)
    back = dom.extended_to_coeff(full)                               # inverse with the zeta^-i / 2^-ext_k post-scale, 3 passes
    assert np.array_equal(back, orc.extended_to_coeff(full, ext_k, THREADS)[: back.shape[0]])


def test_batched_prefix_scan_matches_per_column_oracle():
    """b200_prefix_scan_batch_dev: the independent grand sums / products of a proof in one call (ragged size, distinct initial values)."""
    from ezkl_b200 import device as dev
    n, batch = 5000, 7
    cols = np.stack([orc.gen_scalars(n, seed=5000 + i) for i in range(batch)])
    inits = orc.gen_scalars(batch, seed=5100)
    d = dev.from_host(cols)
    for product in (False, True):
        got = dev.to_host(dev.prefix_scan_batch(d, inits, product))
        for i in range(batch):
            assert np.array_equal(got[i], orc.prefix_scan(cols[i], inits[i], product)), (product, i)
