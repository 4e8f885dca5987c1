#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/msm_sweep_multi.py : BASELINE configs[3] — standalone MSM 2^16..2^26 split across N GPUs
(ShardedMsm: contiguous base slices, local Pippenger, one all-gather of XYZZ partials, local add).  Device-timed, max over ranks.
At world 1 it doubles as the single-GPU sweep; at k <= 20 the sharded result is checked against a single-GPU MSM."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ezkl_b200 import _native as nat  # noqa: E402
from ezkl_b200 import device as dev  # noqa: E402
from ezkl_b200 import parallel as par  # noqa: E402


def main():
    rank, world, local = par.init_distributed("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    torch.cuda.set_device(local)
    nat.init(local)
    out = []
    for k in range(16, 27, 2):
        n = 1 << k
        bases = dev.generate_bases(n, seed=3)
        sm = par.ShardedMsm(bases, n)
        sc = dev.random_scalars(n, batch=1, seed=5)
        if k <= 20 and world > 1:
            full = dev.DeviceBases(bases)
            assert np.array_equal(sm(sc), dev.normalize(dev.msm_batch(full, sc)))
            full.release()
        del bases
        loc = sc[:, sm.lo:sm.hi].contiguous()
        del sc
        torch.cuda.empty_cache()
        sm.combine(sm.partial(loc))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            sm.combine(sm.partial(loc))
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out.append({"k": k, "n_gpus": world, "ms": round(ms, 3), "pairs_per_s": round(n / ms * 1e3, 1)})
        sm.bases.release()
        del loc
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
