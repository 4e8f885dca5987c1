#!/bin/bash
mkdir -p gpurun_out
for pr in "0,0" "0,-1" "-2,0"; do
  BENCH_STREAM_PRIORITIES=$pr python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e --no-parity-gate > gpurun_out/r02_prio.json 2> gpurun_out/r02_prio.err
  tail -c 300 gpurun_out/r02_prio.err
  BENCH_STREAM_PRIORITIES=$pr python bench.py --simulate-rank-of 8 --steps 10 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e --no-parity-gate > gpurun_out/r02_prio8.json 2> gpurun_out/r02_prio8.err
  python - <<PY
import json
for f in ("gpurun_out/r02_prio.json", "gpurun_out/r02_prio8.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("prio $pr", f, "value", d["value"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print("prio $pr", f, "FAILED", e)
PY
done
