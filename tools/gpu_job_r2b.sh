#!/bin/bash
# round-2 session-2 job A (1 GPU): carry-cost probes, small-batch MSM launch lists, simulated rank-of-8 step
mkdir -p gpurun_out
tools/experiments/pipe_probe2 > gpurun_out/r02_pipe_probe2.txt 2>&1
cat gpurun_out/r02_pipe_probe2.txt
python tools/bench_msm.py one 1,2,4,7,8 > gpurun_out/r02_msm_small_batches.txt 2>&1
cat gpurun_out/r02_msm_small_batches.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_msm_b1_b7.csv python tools/bench_msm.py one 1,7 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_msm_b1_b7.csv > gpurun_out/r02_launches_msm_b1_b7.txt 2>&1 || true
python bench.py --simulate-rank-of 8 --steps 10 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e --no-parity-gate > gpurun_out/r02_sim_rank_of_8.json 2> gpurun_out/r02_sim_rank_of_8.err
tail -c 300 gpurun_out/r02_sim_rank_of_8.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_sim_rank_of_8.json").read().strip().splitlines()[-1])
print("sim8 value", d["value"], "issue", d["host_issue_ms_per_step"], "launches", d["gpu_launches"], "classes", d["kernel_class_ms_per_step"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_sim_rank_of_8.csv python bench.py --simulate-rank-of 8 --profile-one-step > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_sim_rank_of_8.csv > gpurun_out/r02_launches_sim_rank_of_8.txt 2>&1 || true
head -40 gpurun_out/r02_launches_sim_rank_of_8.txt
