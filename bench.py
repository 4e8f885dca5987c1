#!/usr/bin/env python
"""bench.py — prove-trace replay of the Halo2/KZG prover hot path on B200 (BASELINE.json metric: prove time (s) at k;
MSM G1 pairs/s and NTT Fr elts/s vs the HBM roofline).

A "step" is ONE proof's worth of hot-path work (SURVEY.md §3.1 stages 1-9 minus synthesize / transcript, which stay on the
CPU in Rust — see DESIGN.md): the MSMs, (i)NTTs, coset NTTs, the quotient-numerator evaluation (evaluate_h, a synthetic
gate program touching every coset column at three rotations) and the column-polynomial passes one `create_proof` issues
for a circuit of the shape named in `config.workload`, on synthetic seeded columns.
  * `value`   : seconds per proof with all columns resident in HBM (device entry points), CUDA-event timed.
  * `e2e`     : the same trace through the C ABI starting from HOST buffers: each witness-derived column is uploaded once from
                pinned host memory (b200_dev_upload), later stages use the device-pointer entry points (the resident-column
                shim of INTEGRATION.md §2b), commitments are normalised on the host and evaluations read back; H2D/D2H and
                the host tail are inside the timed region.
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
    return [
        ("msm_lagrange", A),                        # stage 1: advice commitments
        ("msm_lagrange", L),                        # stage 2: lookup multiplicities m(X)
        ("batch_invert", tr["perm_cols"] + L),      # stage 3: denominators of z(X) and phi(X)
        ("prefix_product", Z),
        ("prefix_sum", L),
        ("msm_lagrange", Z + L),                    #          commitments to z's and phi's
        ("msm_coeff", 1),                           # stage 4: vanishing random polynomial
        ("intt", ncoset),                           # stage 5: Lagrange -> coefficients
        ("quotient", 1),                            # stages 6-7: ncoset coset NTTs -> evaluate_h -> divide by vanishing -> extended iNTT
        ("msm_coeff", Q),
        ("eval", tr["evals"]),                      # stage 8
        ("lincomb", tr["shplonk_sets"]),            # stage 9: SHPLONK linear combinations (npolys polynomials over the rotation sets)
        ("kate_division", tr["shplonk_sets"]),
        ("msm_coeff", 2),
    ]


def n_coset_columns(tr):
    return tr["advice"] + tr["instance"] + tr["perm_z"] + 2 * tr["lookups"]


def count_units(ops, n, tr):
    pairs = sum(c for k, c in ops if k.startswith("msm")) * n
    ntt_elts = sum(c * n for k, c in ops if k == "intt") + (n_coset_columns(tr) + 1) * (n << tr["ext_bits"])
    return pairs, ntt_elts


QUOTIENT_GROUP = 32          # coset columns evaluated per evaluate_h call (partial sums carried in h)
GATE_Y = 0x1234567890ABCDEF1234567890ABCDEF


def gate_program(m):
    """Synthetic gate set over m coset columns + the running sum h (column m): for every column t one degree-2 term reading
    three columns at rotations 0 / +1 / -1, folded with y exactly like evaluate_h folds gates:  h <- h*y + (a*b + c*a - b)."""
    from ezkl_b200 import evaluation as ev
    value = ev.Query(m)
    y = ev.Constant(GATE_Y)
    for t in range(m):
        a, b, c = ev.Query(t), ev.Query((t + 1) % m, 1), ev.Query((t + 2) % m, -1)
        value = value * y + (a * b + c * a - b)
    return ev.QuotientProgram(value)


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed regions run: ONE long-lived `nvidia-smi -lms 100` process (started
    before the warm-up so its start-up cost is outside the timed region); rows are stamped on arrival and only those that fall
    inside a marked region are summarised."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.regions, self.proc = device, [], [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 8:
                    self.rows.append((time.time(), parts))
        except Exception:
            pass

    def mark(self, t0, t1):
        self.regions.append((t0, t1))

    def summary(self):
        time.sleep(0.15)
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=3)
        inside = [r for (ts, r) in self.rows if any(t0 <= ts <= t1 + 0.1 for t0, t1 in self.regions)] or [r for _, r in self.rows]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[1]) for r in inside)
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(inside[0][2]), "power_w_max": max(float(r[3]) for r in inside),
                "reasons": sorted(reasons), "samples": len(inside)}


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
    tmp_n = torch.empty((ncols, n, 4), dtype=torch.int64, device="cuda")
    out_n = torch.empty((ncols, n, 4), dtype=torch.int64, device="cuda")
    xs = dev.to_host(dev.random_scalars(ncols, seed=99))
    one = F.fr_to_limbs(1)
    zeta, zeta2 = F.fr_to_limbs(F.FR_ZETA), F.fr_to_limbs(F.FR_ZETA * F.FR_ZETA % F.FR_MODULUS)
    d = F.fr_from_limbs(dom.extended_ifft_divisor)
    post_ext = [F.fr_to_limbs(d), F.fr_to_limbs(d * F.FR_ZETA * F.FR_ZETA % F.FR_MODULUS), F.fr_to_limbs(d * F.FR_ZETA % F.FR_MODULUS)]
    host_cols = torch.empty((ncols, n, 4), dtype=torch.int64).pin_memory()
    host_cols.copy_(cols.cpu())
    host_evals = torch.empty((max(ncols, 256), 4), dtype=torch.int64).pin_memory()

    def mine(count, base):
        """op instances of one group owned by this rank (round-robin over a running global index)."""
        return [i for i in range(count) if par.column_owner(base + i, world) == rank]

    from ezkl_b200 import evaluation as ev
    ncoset = n_coset_columns(tr)
    log_g = world.bit_length() - 1
    assert world == 1 << log_g and world <= (1 << tr["ext_bits"]), "quotient stage: world must be a power of two <= 2^ext_bits"
    N_ext = 1 << ext_k
    slab = N_ext // world
    group = QUOTIENT_GROUP if ext_k <= 23 else 16
    programs = {}
    d_ext = 1 << tr["ext_bits"]
    tinv_local = np.ascontiguousarray(np.stack([dom.t_evaluations[(rank + world * t) % d_ext] for t in range(max(1, d_ext // world))]))

    def quotient_stage(get_col, put_h):
        """Stages 6-7.  Coset NTTs are dealt by column; evaluate_h runs row-cyclic (row idx on rank idx mod world): since
        world divides 2^ext_bits every Rotation(r) = r * 2^ext_bits rows stays on its rank, so the only exchange is one
        all-to-all per column group; h slabs are all-gathered once for the single extended iNTT on rank 0."""
        h = torch.zeros((slab, 4), dtype=torch.int64, device="cuda")
        for g0 in range(0, ncoset, group):
            gcols = list(range(g0, min(ncoset, g0 + group)))
            my = [j for j in gcols if par.column_owner(j, world) == rank]
            if my:
                src = torch.stack([get_col(j) for j in my])
                ext_my = dev.ntt(src, ext_k, dom.extended_omega, n_in=n, pre=[one, zeta, zeta2])
            else:
                ext_my = torch.empty((0, N_ext, 4), dtype=torch.int64, device="cuda")
            if world > 1:
                view = ext_my.view(len(my), slab, world, 4)
                send = [view[:, :, s_].contiguous() for s_ in range(world)]
                counts = [len([j for j in gcols if par.column_owner(j, world) == q]) for q in range(world)]
                recv = [torch.empty((counts[q], slab, 4), dtype=torch.int64, device="cuda") for q in range(world)]
                dist.all_to_all(recv, send)
                by_col = {}
                for q in range(world):
                    for i_, j in enumerate([j for j in gcols if par.column_owner(j, world) == q]):
                        by_col[j] = recv[q][i_]
                slabs = [by_col[j] for j in gcols]
            else:
                slabs = [ext_my[i_] for i_ in range(len(my))]
            m = len(gcols)
            if m not in programs:
                programs[m] = gate_program(m)
            h = ev.evaluate_h_device(programs[m], slabs + [h], k, ext_k - log_g)
        dev.scale_cycle(h, tinv_local)
        if world > 1:
            parts = [torch.empty_like(h) for _ in range(world)]
            dist.all_gather(parts, h)
            full = torch.stack(parts, dim=1).reshape(1, N_ext, 4).contiguous()      # idx = t*world + rank
        else:
            full = h.view(1, N_ext, 4)
        if rank == 0:
            coeff = dev.ntt(full, ext_k, dom.extended_omega_inv, post=post_ext)
            if put_h is not None:
                put_h(coeff[0, : tr["quotient_pieces"] * n])

    commits = []
    commit_counts = [0] * world
    _g = 0
    for _kind, _count in ops:
        if _kind.startswith("msm"):
            for _i in range(_count):
                commit_counts[par.column_owner(_g + _i, world)] += 1
        _g += _count

    evals = []
    npolys_total = tr["advice"] + tr["fixed"] + tr["perm_cols"] + tr["perm_z"] + 2 * tr["lookups"] + 1 + tr["quotient_pieces"]
    lin_scalars = np.ascontiguousarray(np.tile(xs, (npolys_total // ncols + 1, 1))[:npolys_total])

    def step_device(pool=None):
        """One proof's trace with device-resident columns (pool defaults to the resident synthetic columns)."""
        cols_ = cols if pool is None else pool
        commits.clear()
        evals.clear()
        gidx = 0
        for kind, count in ops:
            m = len(mine(count, gidx)) if kind != "quotient" else 1     # the quotient stage is cooperative: every rank takes part
            gidx += count
            done = 0
            while done < m:
                b = min(m - done, ncols)
                v = cols_[:b]
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
                elif kind == "quotient":
                    quotient_stage(lambda j: cols_[j % ncols], None)
                elif kind == "eval":
                    evals.append(dev.eval_batch(v, xs[:b]))
                elif kind == "lincomb":
                    per_set = max(1, npolys_total // tr["shplonk_sets"])
                    for i in range(b):
                        dev.lincomb([cols_[j % ncols] for j in range(per_set)], lin_scalars[:per_set], out=out_n[i])
                elif kind == "kate_division":
                    for i in range(b):
                        dev.kate_division(v[i], xs[i], out=out_n[i][: n - 1])
                done += b
        pts = torch.cat(commits) if commits else torch.zeros((0, 16), dtype=torch.int64, device="cuda")
        if world > 1:
            par.allgather_columns(pts, commit_counts)      # per-rank counts follow from the deal: no size exchange, no host sync
        return pts

    L = nat.lib()
    import ctypes as C
    n_inputs = tr["advice"] + tr["instance"] + tr["lookups"] + 1      # witness-derived columns that exist only on the host before a proof
    e2e_pool = torch.empty((max(ncols, n_inputs), n, 4), dtype=torch.int64, device="cuda")
    for _s in range(0, e2e_pool.shape[0], ncols):      # valid uniform scalars everywhere: slots a rank does not upload into must not be zeros
        e2e_pool[_s:_s + ncols].copy_(cols[: min(ncols, e2e_pool.shape[0] - _s)])
    upload_stream = torch.cuda.Stream()
    upload_first = torch.cuda.Event()
    upload_done = torch.cuda.Event()

    def step_host():
        """End to end through the C ABI from HOST buffers: every witness-derived column crosses PCIe once (b200_dev_upload from
        pinned memory into a device-resident column), all later stages use the device-pointer entry points, and the step's
        results come back to the host: commitments (XYZZ -> b200_g1_normalize on the host) and the evaluations."""
        my_inputs = [i for i in range(n_inputs) if par.column_owner(i, world) == rank] if world > 1 else list(range(n_inputs))
        h2d = d2h = 0
        # uploads are enqueued on a side stream in column order; the compute stream waits only for the columns the trace reads
        # (the first `ncols` slots), so the tail of the witness upload overlaps the first commit batch
        with torch.cuda.stream(upload_stream):
            for slot, i in enumerate(my_inputs):
                nat.check(L.b200_dev_upload_async(nat.dev(e2e_pool[slot].data_ptr()), C.c_void_p(host_cols[i % ncols].data_ptr()), C.c_size_t(n * 32),
                                                  C.c_void_p(upload_stream.cuda_stream)))
                h2d += n * 32
                if slot == min(ncols, len(my_inputs)) - 1:
                    upload_first.record(upload_stream)
            upload_done.record(upload_stream)
        torch.cuda.current_stream().wait_event(upload_first)
        pts = step_device(e2e_pool)
        torch.cuda.current_stream().wait_event(upload_done)
        jac = dev.normalize(pts)                                  # D2H of the XYZZ partials + host normalisation
        d2h += pts.numel() * 8
        for e in evals:
            host_evals[: e.shape[0]].copy_(e)
            d2h += e.numel() * 8
        torch.cuda.synchronize()
        return h2d, d2h, jac

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

    if args.profile_one_step:          # for `ncu`: setup + exactly one device-resident step, nothing else
        step_device()
        torch.cuda.synchronize()
        print("one step done")
        return

    # ---- device-resident timing (library-side event profiling OFF: nothing but the kernels in the timed region)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    t_reg0 = time.time()
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    sampler.mark(t_reg0, time.time())
    launches = nat.launch_count() - l0
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    # ---- same steps again with per-kernel-class CUDA events (roofline leg; not part of `value`)
    nat.check(L.b200_profile_enable(1))
    for _ in range(args.steps):
        step_device()
    barrier()
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
    t_reg0 = time.time()
    for _ in range(e2e_steps):
        h2d, d2h, _ = step_host()
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    sampler.mark(t_reg0, time.time())
    clocks = sampler.summary()

    pairs, ntt_elts = count_units(ops, n, tr)
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
    _c, _w = C.c_int(0), C.c_int(0)
    nat.check(L.b200_bases_info(C.c_uint64(g_lag.handle), None, C.byref(_c), C.byref(_w)))
    win = _w.value                                  # windows per scalar = bucket additions per (scalar, base) pair
    launch_cols = my_msm_cols * args.steps / max(acc_cnt, 1)
    alg_bytes_per_launch = launch_cols * n * (32.0 + 64.0 / max(launch_cols, 1.0))
    achieved = alg_bytes_per_launch / ((acc_ms / max(acc_cnt, 1)) * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of this same command (k = 17 only)
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_bench_step_k17.json")))["k_accumulate"]
        if k == 17 and tname == "conv2d_mnist" and world == 1:
            rd, wr = cap["dram__bytes_read.sum"]["per_launch"], cap["dram__bytes_write.sum"]["per_launch"]
            traffic = int(sum(rd + wr) * 1e9 / len(rd))
    except Exception:
        pass
    msm_ms = prof["msm_total"][0] / args.steps
    ntt_ms = prof["ntt"][0] / args.steps
    line = {
        "metric": "prove_time_s", "value": round(ms_dev / 1e3, 6), "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev, 3), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32 limbs (254-bit Montgomery integers mod BN254 r/p)", "data": "synthetic",
        "config": {"workload": "prove-trace replay, %s-shaped circuit at k=%d (MSM, NTT, evaluate_h with a synthetic gate program and poly stages of create_proof; "
                               "synthesize and transcript stay on the CPU and are not replayed)" % (tname, k), "k": k, "trace": tr, "msm_pairs_per_step": pairs, "ntt_elts_per_step": ntt_elts,
                   "parallelism": "columns round-robin over %d GPU(s)" % world,
                   "l2": "inputs larger than L2: %d MB of columns + %d MB tables per step" % (ncols * n * 32 >> 20, (2 * n * 64 * 17) >> 20)},
        "e2e": {"value": round(ms_e2e / 1e3, 6), "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "k_accumulate (MSM bucket accumulation)", "bound": "hbm", "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                     "frac": round(achieved / hbm_peak, 5), "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(alg_bytes_per_launch), "avg_launch_ms": round(acc_ms / max(acc_cnt, 1), 4),
                     "kernel_share_of_step": round(acc_ms / args.steps / ms_dev, 4) if ms_dev > 0 else None,
                     "issue_bound": {"what": "bucket additions (XYZZ += affine: 6 multiplies + 2 dedicated squarings (0.78 each) + one two-product multiply (1.5) + 7 add/sub = 9.46 multiply-equivalents in IMAD.WIDE work) per second against the measured "
                                             "254-bit multiply ceiling of 67.5 G mulmod/s (profiles/r01_microbench_mulmod.txt)",
                                     "adds_per_s": round(my_msm_cols * n * win / (acc_ms / args.steps * 1e-3), 1) if acc_ms > 0 else None,
                                     "frac": round(my_msm_cols * n * win * 9.46 / (acc_ms / args.steps * 1e-3) / 67.5e9, 4) if acc_ms > 0 else None},
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
        elif kind == "quotient":
            # stages 6-7 on the CPU: one coset NTT and one evaluate_h group are timed and scaled by their counts
            ncoset = n_coset_columns(tr)
            m = min(ncoset, QUOTIENT_GROUP if ext_k <= 23 else 16)
            t0 = time.perf_counter()
            ext = orc.coeff_to_extended(col, ext_k, threads)
            t_coset = time.perf_counter() - t0
            loads, consts, prog = gate_program(m).arrays()
            t0 = time.perf_counter()
            hq = orc.quotient_eval([ext] * (m + 1), k, ext_k, loads, consts, prog, threads)
            t_eval = time.perf_counter() - t0
            t0 = time.perf_counter()
            orc.extended_to_coeff(orc.divide_by_vanishing(hq, k, ext_k), ext_k, threads)
            t_tail = time.perf_counter() - t0
            return t_coset * ncoset + t_eval * (ncoset / m) + t_tail
        elif kind == "eval":
            orc.eval_polynomial(col, x)
        elif kind == "lincomb":
            per_set = max(1, (tr["advice"] + tr["fixed"] + tr["perm_cols"] + tr["perm_z"] + 2 * tr["lookups"] + 1 + tr["quotient_pieces"]) // tr["shplonk_sets"])
            t0 = time.perf_counter()
            orc.poly_op("axpy", col, col, x, threads=threads)
            return (time.perf_counter() - t0) * per_set
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
            modelled = run_one(kind)
            per.setdefault(kind, []).append(modelled if modelled is not None else time.perf_counter() - t0)
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
    pairs, ntt_elts = count_units(ops, 1 << k, tr)
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
            "config": {"workload": "prove-trace replay, %s-shaped circuit at k=%d (MSM, NTT, evaluate_h with a synthetic gate program and poly stages of create_proof; "
                                   "synthesize and transcript stay on the CPU and are not replayed)" % (tname, k), "k": k, "trace": tr, "msm_pairs_per_step": pairs, "ntt_elts_per_step": ntt_elts},
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
    ap.add_argument("--profile-one-step", action="store_true", help="setup + one device step only (for ncu launch lists)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
