#!/bin/bash
# 1-GPU validation of the round: all GPU tests, smoke, default bench line, k = 20 and k = 22 lines; with NCU=1 also the launch list and the
# ncu --set full captures of one single-stream bench step (reports are reduced to text / JSON on the box: gpurun returns at most 64 MiB)
mkdir -p gpurun_out
S=gpurun_out/r02_final_summary.txt
: > $S
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gpu_tests_1gpu.txt >> $S
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.txt >> $S
python bench.py > gpurun_out/r02_bench_k17.json 2> gpurun_out/r02_bench_k17.err
tail -c 300 gpurun_out/r02_bench_k17.err >> $S
python bench.py --k 20 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_k20.json 2> gpurun_out/r02_bench_k20.err
tail -c 300 gpurun_out/r02_bench_k20.err >> $S
timeout 600 python bench.py --k 22 --steps 3 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e > gpurun_out/r02_bench_k22.json 2> gpurun_out/r02_bench_k22.err
tail -c 300 gpurun_out/r02_bench_k22.err >> $S
python - >> $S <<PY
import json
for f in ["gpurun_out/r02_bench_k17.json", "gpurun_out/r02_bench_k20.json", "gpurun_out/r02_bench_k22.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "hp", (d.get("e2e_host_pointer") or {}).get("value"), "cold", d["cold_start"]["total_s"], "parity", d["parity_checked"],
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d["clocks"], "roofline", d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["issue_bound"]["frac"], "classes", d["kernel_class_ms_per_step"], d["schedule"][:12])
    except Exception as e:
        print(f, "FAILED", e)
PY
if [ -n "$NCU" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_one_step_k17.csv python bench.py --profile-one-step --no-overlap > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r02_launches_one_step_k17.csv k_table_next_level k_g1_generate k_powers k_stage_twiddles > gpurun_out/r02_launches_one_step_k17.txt 2>&1
  ncu --set full --clock-control none -k regex:"k_accumulate" -c 6 -f -o /tmp/ncu_acc python bench.py --profile-one-step --no-overlap > gpurun_out/r02_ncu_full.log 2>&1
  ncu --set full --clock-control none -k regex:"k_ntt_pass2|k_quotient_eval|k_reduce|k_digits|k_scan_buckets" -c 14 -f -o /tmp/ncu_other python bench.py --profile-one-step --no-overlap >> gpurun_out/r02_ncu_full.log 2>&1
  python tools/ncu_to_json.py /tmp/ncu_acc.ncu-rep /tmp/ncu_other.ncu-rep > gpurun_out/r02_ncu_full_bench_step_k17.json 2>> gpurun_out/r02_ncu_full.log
  python tools/ncu_summary.py /tmp/ncu_acc.ncu-rep /tmp/ncu_other.ncu-rep > gpurun_out/r02_ncu_full_bench_step_k17.txt 2>> gpurun_out/r02_ncu_full.log
fi
cat $S
