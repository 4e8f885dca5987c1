#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/multi_gpu_check.py : parity and timing of the intra-op sharding layer on N GPUs.
ShardedMsm (base split + all-gather of XYZZ partials + local add) and ShardedNtt (six-step, one all-to-all) must
reproduce the single-GPU results bit-exactly; prints device-timed throughput (max over ranks)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ezkl_b200 import _native as nat  # noqa: E402
from ezkl_b200 import device as dev  # noqa: E402
from ezkl_b200 import fields as F  # noqa: E402
from ezkl_b200 import parallel as par  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = par.init_distributed("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    torch.cuda.set_device(local)
    nat.init(local)
    res = {"world": world}
    for k in [int(x) for x in os.environ.get("CHECK_KS", "17,20,22").split(",")]:
        n = 1 << k
        bases = dev.generate_bases(n, seed=7)                   # same seed on every rank -> same full vector
        sc = dev.random_scalars(n, batch=2, seed=11)
        sm = par.ShardedMsm(bases, n)
        got = sm(sc)
        if world > 1 or k <= 20:
            full = dev.DeviceBases(bases)
            ref = dev.normalize(dev.msm_batch(full, sc))
            assert np.array_equal(got, ref), "sharded MSM != single-GPU MSM at k=%d" % k
            full.release()
        lo, hi = sm.lo, sm.hi
        loc = sc[:, lo:hi].contiguous()
        ms = timed(lambda: sm.combine(sm.partial(loc)))
        res["msm_k%d" % k] = {"ms": round(ms, 3), "pairs_per_s": round(2 * n / ms * 1e3, 1)}
        sm.bases.release()
        del bases, sc
        # NTT
        w = pow(F.FR_ROOT_OF_UNITY, 1 << (F.FR_S - k), F.FR_MODULUS)
        a = dev.random_scalars(n, seed=13)
        s = par.ShardedNtt(k, w)
        loc_in = s.scatter(a)
        out = s.gather(s.forward(loc_in))
        ref = dev.ntt(a, k, F.fr_to_limbs(w))[0]
        assert torch.equal(out, ref), "sharded NTT != single-GPU NTT at k=%d" % k
        ms = timed(lambda: s.forward(loc_in))
        ms1 = timed(lambda: dev.ntt(a, k, F.fr_to_limbs(w)))
        res["ntt_k%d" % k] = {"ms_sharded": round(ms, 3), "ms_single_gpu": round(ms1, 3), "elts_per_s": round(n / ms * 1e3, 1)}
        del a, s, loc_in, out, ref
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
