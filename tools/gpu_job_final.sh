#!/bin/bash
# final 1-GPU validation: all GPU tests, smoke, the default bench line, k = 20 line, launch list and ncu --set full captures of a bench step
# (reports are reduced to text / JSON on the box: gpurun brings back at most 64 MiB)
mkdir -p gpurun_out
S=gpurun_out/r02_final_summary.txt
: > $S
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gpu_tests_1gpu.txt >> $S
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.txt >> $S
python bench.py > gpurun_out/r02_bench_k17.json 2> gpurun_out/r02_bench_k17.err
tail -c 300 gpurun_out/r02_bench_k17.err >> $S
if [ -n "$EXTRA_AB" ]; then
  for v in 4 2; do
    B200_MSM_DIGIT_CTAS_PER_SM=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e --no-parity-gate > gpurun_out/r02_bench_digitcap$v.json 2> gpurun_out/r02_bench_digitcap$v.err
    B200_MSM_DIGIT_CTAS_PER_SM=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-host-pointer-e2e --no-parity-gate --no-overlap > gpurun_out/r02_bench_digitcap${v}_serial.json 2>> gpurun_out/r02_bench_digitcap$v.err
  done
fi
python bench.py --k 20 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_k20.json 2> gpurun_out/r02_bench_k20.err
tail -c 300 gpurun_out/r02_bench_k20.err >> $S
python - >> $S <<PY
import json, glob
for f in ["gpurun_out/r02_bench_k17.json", "gpurun_out/r02_bench_k20.json"] + sorted(glob.glob("gpurun_out/r02_bench_digitcap*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "hp", (d.get("e2e_host_pointer") or {}).get("value"), "cold", d["cold_start"]["total_s"], "parity", d["parity_checked"],
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d["clocks"], "roofline", d["roofline"]["frac"], d["roofline"]["issue_bound"]["frac"], "classes", d["kernel_class_ms_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_one_step_k17.csv python bench.py --profile-one-step --no-overlap > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_one_step_k17.csv k_table_next_level k_g1_generate k_powers k_stage_twiddles > gpurun_out/r02_launches_one_step_k17.txt 2>&1
head -8 gpurun_out/r02_launches_one_step_k17.txt >> $S
ncu --set full --clock-control none -k regex:"k_accumulate" -c 6 -f -o /tmp/ncu_acc python bench.py --profile-one-step --no-overlap > gpurun_out/r02_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"k_ntt_pass2|k_quotient_eval|k_reduce|k_digits|k_scan_buckets" -c 14 -f -o /tmp/ncu_other python bench.py --profile-one-step --no-overlap >> gpurun_out/r02_ncu_full.log 2>&1
python tools/ncu_to_json.py /tmp/ncu_acc.ncu-rep /tmp/ncu_other.ncu-rep > gpurun_out/r02_ncu_full_bench_step_k17.json 2>> gpurun_out/r02_ncu_full.log
python tools/ncu_summary.py /tmp/ncu_acc.ncu-rep /tmp/ncu_other.ncu-rep > gpurun_out/r02_ncu_full_bench_step_k17.txt 2>> gpurun_out/r02_ncu_full.log
ls -la /tmp/*.ncu-rep >> $S
sz=$(stat -c %s /tmp/ncu_acc.ncu-rep 2>/dev/null || echo 999999999)
if [ "$sz" -lt 30000000 ]; then cp /tmp/ncu_acc.ncu-rep gpurun_out/r02_ncu_full_k_accumulate.ncu-rep; fi
du -sh gpurun_out >> $S
cat $S
