"""Multi-device check, run as its own process on a box with >= 2 GPUs (tests/test_multi_device.py launches it; it can also be
run by hand: `python tests/multi_device_check.py 2`).  ONE process owns N devices through b200_init_multi and every result is
compared with the CPU oracle: host-pointer MSM (columns dealt / base-split), host-pointer NTT (batch dealt / one transform
sharded with the exchanges fused into the passes), and the device-resident sharded entry points."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200_SHARD_MIN_LOGN", "14")          # shard small transforms too, so the oracle finishes in seconds

from ezkl_b200 import _native as nat  # noqa: E402
from ezkl_b200 import halo2 as h2  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    nd = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    L = nat.lib()
    nat.check(L.b200_init_multi(C.c_int(nd)))
    nat._inited = True
    assert L.b200_device_count() == nd
    th = orc.host_threads()
    # ---- MSM through the host-pointer ABI
    n = 1 << 15
    bases_np = orc.gen_bases(n, seed=11, threads=th)
    bases = h2.Bases(bases_np)
    cols = [orc.gen_scalars(n, seed=20 + i) for i in range(2 * nd + 1)]
    cols[1][: n // 2] = 0                                              # half zeros
    exp = [orc.msm(c, bases_np, th) for c in cols]
    got = h2.best_multiexp_batch(cols, bases)                          # batch >= devices: columns dealt
    for i, e in enumerate(exp):
        assert np.array_equal(got[i, :8], e), "dealt MSM column %d" % i
    got1 = h2.best_multiexp(cols[0], bases)                            # one column: (scalar, base) pairs split by range
    assert np.array_equal(got1[:8], exp[0]), "base-split MSM"
    if nd > 2:
        got2 = h2.best_multiexp_batch(cols[:2], bases)                 # fewer columns than devices: split path with a batch
        assert np.array_equal(got2[0, :8], exp[0]) and np.array_equal(got2[1, :8], exp[1]), "base-split MSM, batch 2"
    m = n - 37                                                          # ragged: fewer scalars than registered bases, uneven ranges
    got3 = h2.best_multiexp(cols[2][:m], bases)
    assert np.array_equal(got3[:8], orc.msm(cols[2][:m], bases_np[:m], th)), "ragged base-split MSM"
    print("msm host paths ok")
    # ---- NTT through the host-pointer ABI
    k = 16
    dom = h2.EvaluationDomain(5, k)
    polys = [orc.gen_scalars(1 << k, seed=40 + i) for i in range(nd + 1)]
    got = dom.lagrange_to_coeff_batch(polys)                           # batch dealt over the devices
    for i, p in enumerate(polys):
        assert np.array_equal(got[i], orc.lagrange_to_coeff(p, k, th)), "dealt iNTT %d" % i
    one = dom.lagrange_to_coeff(polys[0])                              # one transform: sharded (B200_SHARD_MIN_LOGN=14)
    assert np.array_equal(one, orc.lagrange_to_coeff(polys[0], k, th)), "sharded iNTT"
    ext = dom.coeff_to_extended(one)                                    # zero-padded coset transform, 2^18, sharded, 2 passes
    assert np.array_equal(ext, orc.coeff_to_extended(one, dom.extended_k, th)), "sharded coset NTT"
    back = dom.extended_to_coeff(ext)
    assert np.array_equal(back[: 1 << k], one) and not back[1 << k:].any(), "sharded extended iNTT"
    big = orc.gen_scalars(1 << 21, seed=77)                            # 3-pass plan, sharded
    w = orc.omega(21)
    assert np.array_equal(h2.best_fft(big, w, 21), orc.best_fft(big, 21, w, th)), "sharded 3-pass NTT"
    print("ntt host paths ok")
    # ---- device-resident sharded entry points (slices allocated with the library's own allocator)
    def dev_slices(arr_slices):
        ptrs = []
        for s, a in enumerate(arr_slices):
            p = C.c_void_p(0)
            nat.check(L.b200_dev_alloc_on(C.c_int(s), C.byref(p), C.c_size_t(max(a.nbytes, 32))))
            nat.check(L.b200_dev_upload(p, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes)))
            ptrs.append(p)
        return ptrs

    N = 1 << 21
    sl = N // nd
    srcs = dev_slices([np.ascontiguousarray(big[s * sl:(s + 1) * sl]) for s in range(nd)])
    tmps = dev_slices([np.zeros((sl, 4), np.uint64) for _ in range(nd)])
    arr = (C.c_void_p * nd)
    nat.check(L.b200_ntt_sharded_dev(arr(*srcs), arr(*tmps), arr(*srcs), C.c_uint32(21), C.c_size_t(N), nat.ptr(np.ascontiguousarray(w)), C.c_int(0), None, C.c_int(0), None))
    nat.check(L.b200_sync_all())
    out = np.zeros((N, 4), np.uint64)
    for s in range(nd):
        nat.check(L.b200_dev_download(out[s * sl:(s + 1) * sl].ctypes.data_as(C.c_void_p), srcs[s], C.c_size_t(sl * 32)))
    assert np.array_equal(out, orc.best_fft(big, 21, w, th)), "b200_ntt_sharded_dev"
    # base-split MSM on resident slices, batch 2
    bounds = []
    base, rem = divmod(n, nd)
    lo = 0
    for s in range(nd):
        hi = lo + base + (1 if s < rem else 0)
        bounds.append((lo, hi))
        lo = hi
    sc = dev_slices([np.ascontiguousarray(np.concatenate([cols[0][a:b], cols[1][a:b]])) for a, b in bounds])
    outj = np.zeros((2, 12), np.uint64)
    nat.check(L.b200_msm_sharded_dev(C.c_uint64(bases.handle), arr(*sc), C.c_size_t(n), C.c_size_t(2), nat.ptr(outj)))
    assert np.array_equal(outj[0, :8], exp[0]) and np.array_equal(outj[1, :8], exp[1]), "b200_msm_sharded_dev"
    print("device-resident sharded entry points ok")
    bases.release()
    nat.shutdown()
    print("multi-device check passed on %d devices" % nd)


if __name__ == "__main__":
    main()
