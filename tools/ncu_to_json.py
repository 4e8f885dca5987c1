#!/usr/bin/env python
"""`ncu --set full` report -> the per-kernel JSON bench.py reads `roofline.traffic` from (read here, no GPU needed).
usage: python tools/ncu_to_json.py gpurun_out/x.ncu-rep > profiles/rNN_ncu_full_bench_step_k17.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size"]
SCALE = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}      # bytes are stored in GB, like the round-1 file


def main():
    out = {}
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        for r in data:
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "").split("<")[0]
            k = out.setdefault(name, {})
            for key in KEYS:
                if key not in col:
                    continue
                v = float(r[col[key]].replace(",", "")) if r[col[key]] not in ("", "n/a") else None
                u = units[col[key]]
                if v is not None and u in SCALE:
                    v, u = v * SCALE[u], "Gbyte"
                e = k.setdefault(key, {"unit": u, "per_launch": []})
                e["per_launch"].append(v)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
