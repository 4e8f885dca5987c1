#!/usr/bin/env python
"""CPU rows of the standalone sweeps (BASELINE.json configs[3]): the restated halo2 algorithms (oracle/, "port")
on all host cores of the box the GPU numbers come from.  MSM n = 2^16 .. 2^24 (uniform scalars; stops at 2^24: one 2^26 point
takes minutes), best_fft log n = 17 .. 25.  Best of `reps` after one warm-up; prints one JSON document."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402


def best(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    th = orc.host_threads()
    out = {"cores": th, "kind": "port (restated halo2 best_multiexp / best_fft, oracle/bn254_oracle.c; not the Rust binary)", "msm": [], "ntt": []}
    for k in (16, 18, 20, 22, 24):
        n = 1 << k
        bases = orc.gen_bases(n, seed=3, threads=th)
        sc = orc.gen_scalars(n, seed=5)
        s = best(lambda: orc.msm(sc, bases, th), 2 if k <= 20 else 1)
        out["msm"].append({"k": k, "s": round(s, 4), "pairs_per_s": round(n / s, 1)})
        print(out["msm"][-1], flush=True, file=sys.stderr)
        del bases, sc
    for k in (17, 19, 20, 22, 23, 25):
        a = orc.gen_scalars(1 << k, seed=7)
        w = orc.omega(k)
        s = best(lambda: orc.best_fft(a, k, w, th), 2 if k <= 22 else 1)
        out["ntt"].append({"log_n": k, "s": round(s, 4), "elts_per_s": round((1 << k) / s, 1)})
        print(out["ntt"][-1], flush=True, file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
