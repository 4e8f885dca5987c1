#!/bin/bash
# A/B of the two-stream schedule (1 GPU): full step and simulated rank-of-8 step, with and without overlap
mkdir -p gpurun_out
for mode in "" "--no-overlap"; do
  tag=$([ -z "$mode" ] && echo two_stream || echo one_stream)
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e $mode > gpurun_out/r02_bench_k17_$tag.json 2> gpurun_out/r02_bench_k17_$tag.err
  tail -c 400 gpurun_out/r02_bench_k17_$tag.err
  python bench.py --simulate-rank-of 8 --steps 10 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e --no-parity-gate $mode > gpurun_out/r02_sim8_$tag.json 2> gpurun_out/r02_sim8_$tag.err
  tail -c 400 gpurun_out/r02_sim8_$tag.err
  python - <<PY
import json
for f in ("gpurun_out/r02_bench_k17_$tag.json", "gpurun_out/r02_sim8_$tag.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("$tag", f, "value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity_checked"], "issue", d["host_issue_ms_per_step"], "classes", d["kernel_class_ms_per_step"])
    except Exception as e:
        print("$tag", f, "FAILED", e)
PY
done
