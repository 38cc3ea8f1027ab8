#!/bin/bash
# ncu evidence for one window-group forward (run on the GPU box through gpurun):
#  1. light metrics pass over every launch of the group  -> gpurun_out/plan_metrics_raw.csv
#  2. --set full (+ source) capture of the longest conv_tc_kernel launch -> gpurun_out/top_full_raw.csv, top_source.csv
# The .ncu-rep files stay in /tmp (gpurun_out is limited to 64 MiB).
set -x
export PROFILE_GROUP=${PROFILE_GROUP:-8}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__issue_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/plan_metrics python tools/ncu_plan.py gpurun_out/plan_steps.json > gpurun_out/ncu_plan.log 2>&1
ncu -i /tmp/plan_metrics.ncu-rep --page raw --csv > gpurun_out/plan_metrics_raw.csv 2>/dev/null
SKIP=$(python - <<'PY'
import json
st = json.load(open("gpurun_out/plan_steps.json"))["steps"]
import csv
rows = list(csv.reader(open("gpurun_out/plan_metrics_raw.csv")))
hdr, data = rows[0], rows[2:]
d = hdr.index("gpu__time_duration.sum"); k = hdr.index("Kernel Name")
best, bi = -1, 0
n = 0
for r, s in zip(data, st):
    if "conv_tc_kernel" in r[k]:
        v = float(r[d].replace(",", ""))
        if v > best: best, bi, name = v, n, s["name"]
        n += 1
open("gpurun_out/top_launch.txt", "w").write(f"{name} skip={bi} dur={best}\n")
print(bi)
PY
)
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name regex:conv_tc_kernel --launch-skip $SKIP --launch-count 1 -f -o /tmp/top_full python tools/ncu_plan.py /tmp/steps2.json >> gpurun_out/ncu_plan.log 2>&1
ncu -i /tmp/top_full.ncu-rep --page raw --csv > gpurun_out/top_full_raw.csv 2>/dev/null
ncu -i /tmp/top_full.ncu-rep --page source --csv > gpurun_out/top_source.csv 2>/dev/null
ncu -i /tmp/top_full.ncu-rep --page details > gpurun_out/top_details.txt 2>/dev/null
ls -la gpurun_out/plan_metrics_raw.csv gpurun_out/top_full_raw.csv gpurun_out/top_source.csv gpurun_out/top_launch.txt
cat gpurun_out/top_launch.txt
