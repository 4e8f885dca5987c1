#!/bin/bash
# N-GPU job: one process owning N devices (oracle-checked, Python and C++ callers), then the process-per-GPU bench at N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python tests/multi_device_check.py $N > gpurun_out/r02_multi_device_check_n$N.txt 2>&1
echo "multi_device_check rc=$?"; tail -4 gpurun_out/r02_multi_device_check_n$N.txt
B200_SHARD_MIN_LOGN=14 timeout 300 tests/cpp/test_multi $N >> gpurun_out/r02_multi_device_check_n$N.txt 2>&1
echo "test_multi rc=$?"; tail -2 gpurun_out/r02_multi_device_check_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/r02_bench_k17_n$N.json 2> gpurun_out/r02_bench_k17_n$N.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_k17_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_k17_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", d["value"], "e2e", d["e2e"]["value"], "in_process", d.get("in_process"), "classes", d.get("kernel_class_ms_per_step"), "collectives", d.get("collectives_ms_per_step"), "issue", d.get("host_issue_ms_per_step"))
PY
