#!/usr/bin/env python
"""bench.py — prove-trace replay of the Halo2/KZG prover hot path on B200 (BASELINE.json metric: prove time (s) at k;
MSM G1 pairs/s and NTT Fr elts/s vs the HBM roofline).

A "step" is ONE proof's worth of hot-path work (SURVEY.md §3.1 stages 1-9 minus synthesize / transcript / evaluate_h,
which are outside this round's kernels — see DESIGN.md): the MSMs, (i)NTTs, coset NTTs and column-polynomial passes one
`create_proof` issues for a circuit of the shape named in `config.workload`, on synthetic seeded columns.
  * `value`   : seconds per proof with all columns resident in HBM (device entry points), CUDA-event timed.
  * `e2e`     : the same trace through the host-buffer C ABI (what the halo2 shim calls): pinned host columns in, results
                back in host memory, H2D/D2H inside the timed region.
  * `roofline`: the dominant kernel (MSM bucket accumulation) against the measured HBM peak, timed with CUDA events
                inside the library on the launching stream.
  * `cpu_baseline` / `--impl reference`: the CPU restatement of halo2's Rayon algorithms (oracle/, "port") on the box's
                host cores, on a bounded sample of the same trace.
N > 1 (torchrun): independent columns are dealt round-robin to ranks (strong scaling, no data-path collective inside an
op; one small all-gather of the commitments per step), timed as max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Circuit-shaped op traces (counts per proof).  Shape parameters come from the reference's own k=6 fixture proof
# (114 commitments + 2 SHPLONK points, 231 evaluations, 38 fixed / 32 permutation columns, extended domain 8n:
# SURVEY.md Appendix B/D4) held fixed while k grows; see DESIGN.md §measurement.
TRACES = {
    "conv2d_mnist": dict(advice=60, lookups=20, perm_cols=32, perm_z=6, instance=1, quotient_pieces=7, fixed=38, evals=231, ext_bits=3,
                         shplonk_sets=4),
    "accum_einsum_matmul": dict(advice=12, lookups=2, perm_cols=8, perm_z=2, instance=1, quotient_pieces=5, fixed=6, evals=60, ext_bits=3,
                                shplonk_sets=3),
}
CONFIG_FOR_K = {17: "conv2d_mnist", 20: "accum_einsum_matmul", 22: "conv2d_mnist", 9: "accum_einsum_matmul"}


def trace_ops(tr):
    """Expands a trace into op groups: (kind, count) in create_proof order."""
    A, L, Z, I, Q = tr["advice"], tr["lookups"], tr["perm_z"], tr["instance"], tr["quotient_pieces"]
    ncoset = A + I + Z + 2 * L                      # columns that go coeff -> extended coset for the quotient
    npolys = A + tr["fixed"] + tr["perm_cols"] + Z + 2 * L + 1 + Q
    return [
        ("msm_lagrange", A),                        # stage 1: advice commitments
        ("msm_lagrange", L),                        # stage 2: lookup multiplicities m(X)
        ("batch_invert", tr["perm_cols"] + L),      # stage 3: denominators of z(X) and phi(X)
        ("prefix_product", Z),
        ("prefix_sum", L),
        ("msm_lagrange", Z + L),                    #          commitments to z's and phi's
        ("msm_coeff", 1),                           # stage 4: vanishing random polynomial
        ("intt", ncoset),                           # stage 5: Lagrange -> coefficients
        ("coset_ntt", ncoset),                      # stage 6: coefficients -> extended coset (evaluate_h itself: not replayed)
        ("divide_vanishing", 1),                    # stage 7
        ("ext_intt", 1),
        ("msm_coeff", Q),
        ("eval", tr["evals"]),                      # stage 8
        ("axpy", npolys),                           # stage 9: SHPLONK linear combinations
        ("kate_division", tr["shplonk_sets"]),
        ("msm_coeff", 2),
    ]


def count_units(ops, n, ext_bits):
    pairs = sum(c for k, c in ops if k.startswith("msm")) * n
    ntt_elts = sum(c * (n if k == "intt" else (n << ext_bits)) for k, c in ops if k in ("intt", "coset_ntt", "ext_intt"))
    return pairs, ntt_elts


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[1]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][2]), "reasons": sorted(reasons), "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from ezkl_b200 import _native as nat
    from ezkl_b200 import device as dev
    from ezkl_b200 import fields as F
    from ezkl_b200 import halo2 as h2
    from ezkl_b200 import parallel as par

    rank, world, local = par.init_distributed("nccl" if args.gpus > 1 else None)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local)
    nat.init(local)
    k = args.k
    n = 1 << k
    tname = args.trace or CONFIG_FOR_K.get(k, "conv2d_mnist")
    tr = TRACES[tname]
    ops = trace_ops(tr)
    ext_k = k + tr["ext_bits"]
    dom = h2.EvaluationDomain((1 << tr["ext_bits"]) + 1, k)
    assert dom.extended_k == ext_k

    # ---- synthetic inputs (seeded), resident on the device and mirrored in pinned host memory for the e2e leg
    g_lag = dev.DeviceBases(dev.generate_bases(n, seed=0xE2C1B200))
    g_coef = dev.DeviceBases(dev.generate_bases(n, seed=0xE2C1B201))
    ncols = max(c for _, c in ops if True)
    ncols = min(ncols, 64)                                      # column pool; ops cycle through it
    cols = dev.random_scalars(n, batch=ncols, seed=1234 + rank)
    ext_chunk = max(1, min(16, (4 << 30) // (32 << ext_k)))     # extended columns processed per call
    ext_buf = torch.empty((ext_chunk, 1 << ext_k, 4), dtype=torch.int64, device="cuda")
    ext_tmp = torch.empty_like(ext_buf)
    tmp_n = torch.empty((ncols, n, 4), dtype=torch.int64, device="cuda")
    out_n = torch.empty((ncols, n, 4), dtype=torch.int64, device="cuda")
    xs = dev.to_host(dev.random_scalars(ncols, seed=99))
    one = F.fr_to_limbs(1)
    zeta, zeta2 = F.fr_to_limbs(F.FR_ZETA), F.fr_to_limbs(F.FR_ZETA * F.FR_ZETA % F.FR_MODULUS)
    d = F.fr_from_limbs(dom.extended_ifft_divisor)
    post_ext = [F.fr_to_limbs(d), F.fr_to_limbs(d * F.FR_ZETA * F.FR_ZETA % F.FR_MODULUS), F.fr_to_limbs(d * F.FR_ZETA % F.FR_MODULUS)]
    host_cols = torch.empty((ncols, n, 4), dtype=torch.int64).pin_memory()
    host_cols.copy_(cols.cpu())
    host_ext = torch.empty((1 << ext_k, 4), dtype=torch.int64).pin_memory()
    host_out = torch.empty((ncols, n, 4), dtype=torch.int64).pin_memory()

    def mine(count, base):
        """op instances of one group owned by this rank (round-robin over a running global index)."""
        return [i for i in range(count) if par.column_owner(base + i, world) == rank]

    commits = []

    def step_device():
        """One proof's trace with device-resident columns."""
        commits.clear()
        gidx = 0
        for kind, count in ops:
            m = len(mine(count, gidx))
            gidx += count
            done = 0
            while done < m:
                b = min(m - done, ncols)
                v = cols[:b]
                if kind == "msm_lagrange":
                    commits.append(dev.msm_batch(g_lag, v))
                elif kind == "msm_coeff":
                    commits.append(dev.msm_batch(g_coef, v))
                elif kind == "batch_invert":
                    out_n[:b].copy_(v)
                    dev.batch_invert(out_n[:b])
                elif kind == "prefix_product":
                    for i in range(b):
                        dev.prefix_scan(v[i], one, True, out=out_n[i])
                elif kind == "prefix_sum":
                    for i in range(b):
                        dev.prefix_scan(v[i], one, False, out=out_n[i])
                elif kind == "intt":
                    dev.ntt(v, k, dom.omega_inv, post=[dom.ifft_divisor], out=out_n[:b], tmp=tmp_n[:b])
                elif kind == "coset_ntt":
                    for c0 in range(0, b, ext_chunk):
                        cb = min(ext_chunk, b - c0)
                        dev.ntt(v[c0:c0 + cb], ext_k, dom.extended_omega, n_in=n, pre=[one, zeta, zeta2], out=ext_buf[:cb], tmp=ext_tmp[:cb])
                elif kind == "divide_vanishing":
                    for _ in range(b):
                        dev.scale_cycle(ext_buf[0], dom.t_evaluations)
                elif kind == "ext_intt":
                    for _ in range(b):
                        dev.ntt(ext_buf[:1], ext_k, dom.extended_omega_inv, post=post_ext, out=ext_buf[:1], tmp=ext_tmp[:1])
                elif kind == "eval":
                    dev.eval_batch(v, xs[:b])
                elif kind == "axpy":
                    for i in range(b):
                        dev.poly_op("axpy", out_n[0], v[i], s=xs[i], out=out_n[0])
                elif kind == "kate_division":
                    for i in range(b):
                        dev.kate_division(v[i], xs[i], out=out_n[i][: n - 1])
                done += b
        pts = torch.cat(commits) if commits else torch.zeros((0, 16), dtype=torch.int64, device="cuda")
        if world > 1:
            cnt = torch.tensor([pts.shape[0]], device="cuda")
            cnts = [torch.zeros_like(cnt) for _ in range(world)]
            dist.all_gather(cnts, cnt)
            par.allgather_columns(pts, [int(c.item()) for c in cnts])
        return pts

    L = nat.lib()
    import ctypes as C

    def step_host():
        """The same trace through the host-buffer C ABI (pinned host memory in and out)."""
        hc = host_cols.numpy().view(np.uint64)
        ho = host_out.numpy().view(np.uint64)
        he = host_ext.numpy().view(np.uint64)
        h2d = d2h = 0
        gidx = 0
        res = []
        for kind, count in ops:
            m = len(mine(count, gidx))
            gidx += count
            done = 0
            while done < m:
                b = min(m - done, ncols)
                v = [hc[i] for i in range(b)]
                if kind in ("msm_lagrange", "msm_coeff"):
                    res.append(h2.best_multiexp_batch(v, _HostBases(g_lag if kind == "msm_lagrange" else g_coef)))
                    h2d += b * n * 32; d2h += b * 128
                elif kind == "batch_invert":
                    for i in range(b):
                        nat.check(L.b200_batch_invert(nat.ptr(hc[i]), C.c_size_t(n)))
                    h2d += b * n * 32; d2h += b * n * 32
                elif kind in ("prefix_product", "prefix_sum"):
                    for i in range(b):
                        nat.check(L.b200_prefix_scan(C.c_int(1 if kind == "prefix_product" else 0), nat.ptr(hc[i]), C.c_size_t(n), nat.ptr(one), nat.ptr(ho[i])))
                    h2d += b * n * 32; d2h += b * n * 32
                elif kind == "intt":
                    nat.check(L.b200_ifft_batch(nat.ptr_array(v), C.c_size_t(b), C.c_uint32(k), nat.ptr(dom.omega_inv), nat.ptr(dom.ifft_divisor)))
                    h2d += b * n * 32; d2h += b * n * 32
                elif kind == "coset_ntt":
                    for i in range(b):
                        nat.check(L.b200_coeff_to_extended(nat.ptr(hc[i]), C.c_size_t(n), C.c_uint32(ext_k), nat.ptr(dom.extended_omega), nat.ptr(dom.g_coset), nat.ptr(he)))
                    h2d += b * n * 32; d2h += b * (32 << ext_k)
                elif kind == "divide_vanishing":
                    for _ in range(b):
                        nat.check(L.b200_poly_scale_cycle(nat.ptr(he), C.c_size_t(1 << ext_k), nat.ptr(dom.t_evaluations), C.c_uint32(dom.t_evaluations.shape[0])))
                    h2d += b * (32 << ext_k); d2h += b * (32 << ext_k)
                elif kind == "ext_intt":
                    for _ in range(b):
                        nat.check(L.b200_extended_to_coeff(nat.ptr(he), C.c_uint32(ext_k), nat.ptr(dom.extended_omega_inv), nat.ptr(dom.extended_ifft_divisor), nat.ptr(dom.g_coset)))
                    h2d += b * (32 << ext_k); d2h += b * (32 << ext_k)
                elif kind == "eval":
                    res.append(h2.eval_polynomial_batch(v, xs[:b]))
                    h2d += b * n * 32; d2h += b * 32
                elif kind == "axpy":
                    for i in range(b):
                        nat.check(L.b200_poly_op(C.c_int(4), nat.ptr(ho[0]), nat.ptr(hc[i]), nat.ptr(xs[i]), nat.ptr(ho[0]), C.c_size_t(n)))
                    h2d += 2 * b * n * 32; d2h += b * n * 32
                elif kind == "kate_division":
                    for i in range(b):
                        nat.check(L.b200_kate_division(nat.ptr(hc[i]), C.c_size_t(n), nat.ptr(xs[i]), nat.ptr(ho[i][: n - 1])))
                    h2d += b * n * 32; d2h += b * (n - 1) * 32
                done += b
        return h2d, d2h

    class _HostBases:      # adapter: halo2.best_multiexp_batch wants .handle / .n
        def __init__(self, db):
            self.handle, self.n = db.handle, db.n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    nat.check(L.b200_profile_enable(1))
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    launches = nat.launch_count() - l0
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    prof = {}
    for cls, name in ((0, "msm_accumulate"), (1, "msm_total"), (2, "ntt")):
        ms, cnt = C.c_double(0), C.c_uint64(0)
        nat.check(L.b200_profile_read(cls, C.byref(ms), C.byref(cnt)))
        prof[name] = (ms.value, cnt.value)
    nat.check(L.b200_profile_enable(0))

    # ---- end-to-end (host buffers through the C ABI)
    e2e_steps = max(1, min(args.steps, 3))
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h2d, d2h = step_host()
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    clocks = sampler.summary()

    pairs, ntt_elts = count_units(ops, n, tr["ext_bits"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    # dominant kernel: MSM bucket accumulation.  Algorithmic bytes = 32 B scalar + 64/b B shared base per pair
    # (SURVEY.md §8d), per launch = pairs in that launch; summed over the step and divided by the summed kernel time.
    my_msm_cols = 0
    gidx = 0
    for kind, count in ops:
        if kind.startswith("msm"):
            my_msm_cols += len(mine(count, gidx))
        gidx += count
    acc_ms, acc_cnt = prof["msm_accumulate"]
    launch_cols = my_msm_cols * args.steps / max(acc_cnt, 1)
    alg_bytes_per_launch = launch_cols * n * (32.0 + 64.0 / max(launch_cols, 1.0))
    achieved = alg_bytes_per_launch / ((acc_ms / max(acc_cnt, 1)) * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    msm_ms = prof["msm_total"][0] / args.steps
    ntt_ms = prof["ntt"][0] / args.steps
    line = {
        "metric": "prove_time_s", "value": round(ms_dev / 1e3, 6), "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev, 3), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32 limbs (254-bit Montgomery integers mod BN254 r/p)", "data": "synthetic",
        "config": {"workload": "prove-trace replay, %s-shaped circuit at k=%d (MSM+NTT+poly stages of create_proof; synthesize, transcript and "
                               "evaluate_h not replayed)" % (tname, k), "k": k, "trace": tr, "msm_pairs_per_step": pairs, "ntt_elts_per_step": ntt_elts,
                   "parallelism": "columns round-robin over %d GPU(s)" % world,
                   "l2": "inputs larger than L2: %d MB of columns + %d MB tables per step" % (ncols * n * 32 >> 20, (2 * n * 64 * 17) >> 20)},
        "e2e": {"value": round(ms_e2e / 1e3, 6), "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "k_accumulate (MSM bucket accumulation)", "bound": "hbm", "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                     "frac": round(achieved / hbm_peak, 5), "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(alg_bytes_per_launch), "avg_launch_ms": round(acc_ms / max(acc_cnt, 1), 4),
                     "note": "integer-issue bound (254-bit modular arithmetic), not HBM bound: see DESIGN.md"},
        "msm_pairs_per_s": round(pairs / world / (msm_ms * 1e-3), 1) * world if msm_ms > 0 else None,
        "ntt_elts_per_s": round(ntt_elts / world / (ntt_ms * 1e-3), 1) * world if ntt_ms > 0 else None,
        "msm_ms_per_step": round(msm_ms, 3), "ntt_ms_per_step": round(ntt_ms, 3),
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_trace(k, tname, budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle ("port" of halo2's Rayon algorithms) on the host cores, bounded sample of the same trace
def cpu_trace(k, tname, budget_s=20.0, threads=None):
    from oracle import oracle as orc
    threads = threads or orc.host_threads()
    n = 1 << k
    tr = TRACES[tname]
    ops = trace_ops(tr)
    ext_k = k + tr["ext_bits"]
    bases = orc.gen_bases(n, seed=5, threads=threads)
    col = orc.gen_scalars(n, seed=6)
    x = orc.gen_scalars(1, seed=7)[0]
    one = orc.fr_one()
    ext = None

    def run_one(kind):
        nonlocal ext
        if kind.startswith("msm"):
            orc.msm(col, bases, threads)
        elif kind == "batch_invert":
            orc.batch_invert(col)
        elif kind == "prefix_product":
            orc.prefix_scan(col, one, True)
        elif kind == "prefix_sum":
            orc.prefix_scan(col, one, False)
        elif kind == "intt":
            orc.lagrange_to_coeff(col, k, threads)
        elif kind == "coset_ntt":
            ext = orc.coeff_to_extended(col, ext_k, threads)
        elif kind == "divide_vanishing":
            orc.divide_by_vanishing(ext if ext is not None else orc.coeff_to_extended(col, ext_k, threads), k, ext_k)
        elif kind == "ext_intt":
            orc.extended_to_coeff(ext if ext is not None else orc.coeff_to_extended(col, ext_k, threads), ext_k, threads)
        elif kind == "eval":
            orc.eval_polynomial(col, x)
        elif kind == "axpy":
            orc.poly_op("axpy", col, col, x, threads=threads)
        elif kind == "kate_division":
            orc.kate_division(col, x)

    # time one instance of every op kind, then extrapolate by count; repeat kinds until the budget is used
    kinds = []
    for kind, _ in ops:
        if kind not in kinds:
            kinds.append(kind)
    per = {}
    t_start = time.perf_counter()
    reps = 0
    while True:
        for kind in kinds:
            t0 = time.perf_counter()
            run_one(kind)
            per.setdefault(kind, []).append(time.perf_counter() - t0)
        reps += 1
        if time.perf_counter() - t_start > budget_s or reps >= 5:
            break
    total = sum(min(per[kind]) * count for kind, count in ops)
    return {"value": round(total, 4), "unit": "s", "cores": threads, "kind": "port",
            "sample": "each of the %d op kinds of the trace timed %d time(s) on the host (best-of), multiplied by its per-proof count; "
                      "restated halo2 algorithms (oracle/bn254_oracle.c), not the Rust binary" % (len(kinds), reps),
            "per_op_s": {kk: round(min(v), 5) for kk, v in per.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k = args.k
    tname = args.trace or CONFIG_FOR_K.get(k, "conv2d_mnist")
    tr = TRACES[tname]
    ops = trace_ops(tr)
    pairs, ntt_elts = count_units(ops, 1 << k, tr["ext_bits"])
    vals = []
    base = None
    for _ in range(args.warmup + args.steps):
        base = cpu_trace(k, tname, budget_s=args.cpu_budget / max(1, args.steps))
        vals.append(base["value"])
    v = sum(vals[args.warmup:]) / max(1, len(vals[args.warmup:]))
    base["value"] = round(v, 4)
    line = {"impl": "reference", "metric": "prove_time_s", "value": round(v, 4), "unit": "s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(v * 1e3, 1), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64 limbs (254-bit Montgomery integers)", "data": "synthetic",
            "config": {"workload": "prove-trace replay, %s-shaped circuit at k=%d (MSM+NTT+poly stages of create_proof; synthesize, transcript and "
                                   "evaluate_h not replayed)" % (tname, k), "k": k, "trace": tr, "msm_pairs_per_step": pairs, "ntt_elts_per_step": ntt_elts},
            "cpu_baseline": base,
            "e2e": {"value": round(v, 4), "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=17)
    ap.add_argument("--trace", default=None, choices=[None] + list(TRACES))
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
