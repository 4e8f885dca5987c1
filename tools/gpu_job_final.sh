#!/bin/bash
# final 1-GPU validation: all GPU tests, smoke, the default bench line, k = 20 line, launch list and one ncu --set full --clock-control none --import-source on -k regex:"k_accumulate" -c 6 -f -o gpurun_out/r02_ncu_full_k_accumulate python bench.py --profile-one-step --no-overlap > gpurun_out/r02_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ntt_pass2|k_quotient_eval|k_reduce|k_digits|k_scan_buckets" -c 22 -f -o gpurun_out/r02_ncu_full_other python bench.py --profile-one-step --no-overlap >> gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
