#!/usr/bin/env python
"""Sweep of the bucket-reduction geometry (buckets per thread x CTA size) per batch size, k = 17, c = 16: the library's own
per-class CUDA-event times (recode / accumulate / tail) of a batched MSM call.  Re-initialises the library per setting
(B200_MSM_REDUCE_M / B200_MSM_REDUCE_THREADS are read in b200_init)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ezkl_b200 import _native as nat  # noqa: E402
from ezkl_b200 import device as dev  # noqa: E402


def measure(bases, sc, reps=5):
    L = nat.lib()
    dev.msm_batch(bases, sc)
    torch.cuda.synchronize()
    nat.check(L.b200_profile_enable(1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dev.msm_batch(bases, sc)
    e1.record()
    torch.cuda.synchronize()
    out = {}
    for cls, name in ((4, "recode"), (0, "acc"), (5, "tail"), (1, "total")):
        ms, cnt = C.c_double(0), C.c_uint64(0)
        nat.check(L.b200_profile_read(cls, C.byref(ms), C.byref(cnt)))
        out[name] = ms.value / max(cnt.value, 1)
    nat.check(L.b200_profile_enable(0))
    out["wall"] = e0.elapsed_time(e1) / reps
    return out


def main():
    k = int(os.environ.get("SWEEP_K", "17"))
    n = 1 << k
    batches = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,2,4,7,16,26").split(",")]
    settings = [(0, 0)] + [(m, t) for t in (64, 128, 256) for m in (2, 4, 8, 16, 32)]
    nat.init(0)
    pts = dev.generate_bases(n, seed=3)
    scs = {b: dev.random_scalars(n, batch=b, seed=5) for b in batches}
    torch.cuda.synchronize()
    for m, t in settings:
        nat.shutdown()
        for name, v in (("B200_MSM_REDUCE_M", m), ("B200_MSM_REDUCE_THREADS", t)):
            if v:
                os.environ[name] = str(v)
            else:
                os.environ.pop(name, None)
        nat.init(0)
        bases = dev.DeviceBases(pts, window_bits=16 if k == 17 else 0)
        for b in batches:
            r = measure(bases, scs[b])
            print("k=%d batch=%3d  M=%-4s threads=%-4s  tail %7.3f ms  recode %7.3f  acc %7.3f  total %7.3f" % (k, b, m or "auto", t or "auto", r["tail"], r["recode"], r["acc"], r["total"]), flush=True)
        bases.release()


if __name__ == "__main__":
    main()
