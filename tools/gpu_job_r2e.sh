#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bench_shapes.py -m gpu -x -q -k "msm or srs or kzg or commit or bench or shapes or splitting or empty or reentrancy or group_law" 2>&1 | tail -5 | tee gpurun_out/r02_gputests_msm_tail.txt
python tools/bench_msm.py one 1,2,4,7,8,16,26,60 > gpurun_out/r02_msm_small_batches_after.txt 2>&1
cat gpurun_out/r02_msm_small_batches_after.txt
python bench.py --simulate-rank-of 8 --steps 10 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e --no-parity-gate > gpurun_out/r02_sim_rank_of_8.json 2> gpurun_out/r02_sim_rank_of_8.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_sim_rank_of_8.json").read().strip().splitlines()[-1])
print("sim8 value", d["value"], "issue", d["host_issue_ms_per_step"], "launches", d["gpu_launches"], "classes", d["kernel_class_ms_per_step"])
PY
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e > gpurun_out/r02_bench_k17_quick.json 2> gpurun_out/r02_bench_k17_quick.err
tail -c 300 gpurun_out/r02_bench_k17_quick.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_k17_quick.json").read().strip().splitlines()[-1])
print("N=1 value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity_checked"], "classes", d["kernel_class_ms_per_step"], d["roofline"]["issue_bound"]["frac"])
PY
