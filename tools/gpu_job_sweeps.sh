python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_gputests_c.txt
python tools/cpu_sweep.py > gpurun_out/r02_cpu_sweep.json 2> gpurun_out/r02_cpu_sweep.err
python tools/bench_msm.py sweep > gpurun_out/r02_msm_sweep.txt 2>&1
python tools/bench_ntt.py > gpurun_out/r02_ntt_sweep.txt 2>&1
ncu --set full --clock-control none -k regex:k_quotient_eval -s 2 -c 1 -f -o gpurun_out/r02_quotient_436 python tools/bench_quotient.py --reps 1 > gpurun_out/r02_quotient_436.txt 2>&1
python bench.py --k 20 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_k20.json 2> gpurun_out/r02_bench_k20.err
tail -c 400 gpurun_out/r02_bench_k20.err
