#!/bin/bash
# Round evidence on ONE B200 (run through gpurun): everything lands in gpurun_out/ with an rNN_ prefix.
#   usage: bash tools/gpu_evidence.sh r02
R=${1:-r02}
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__issue_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic
# 1. launch list of the bench command (per-launch times are cold-cache and serialised: compare SHARES, not absolutes)
# (the timed region replays CUDA graphs: --graph-profiling node lists their kernel nodes one by one; the filter keeps the
# repo's kernels and drops the torch element-wise kernels of the set-up)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node \
    -k regex:'conv_tc|gate_logits|cin1|att_gate|sw_finalize|smallcout|conv_act8' -c 1600 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > $O/${R}_bench_under_ncu.log 2>&1
# 2. per-launch metrics of one window group (same plan as the timed region)
export PROFILE_GROUP=8
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/plan_metrics python tools/ncu_plan.py $O/${R}_plan_steps.json > $O/${R}_ncu_plan.log 2>&1
ncu -i /tmp/plan_metrics.ncu-rep --page raw --csv > $O/${R}_ncu_group_metrics_raw.csv 2>/dev/null
python tools/ncu_summarise.py $O/${R}_ncu_group_metrics_raw.csv $O/${R}_plan_steps.json $O/${R}_ncu_group_summary.json
# 3. --set full capture of the longest launch of the group
SKIP=$(python - <<PY
import json
s = json.load(open("$O/${R}_ncu_group_summary.json"))["launches"]
i = max(range(len(s)), key=lambda k: s[k].get("dur_us") or 0)
open("$O/${R}_top_launch.txt", "w").write(f"{s[i]['step']} index={i} dur_us={s[i]['dur_us']} kernel={s[i]['kernel']}\n")
print(i)
PY
)
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off --launch-skip $SKIP --launch-count 1 -f -o /tmp/top_full python tools/ncu_plan.py /tmp/steps2.json >> $O/${R}_ncu_plan.log 2>&1
ncu -i /tmp/top_full.ncu-rep --page raw --csv > $O/${R}_top_full_raw.csv 2>/dev/null
ncu -i /tmp/top_full.ncu-rep --page details > $O/${R}_top_details.txt 2>/dev/null
ncu -i /tmp/top_full.ncu-rep --page source --csv > /tmp/top_source.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("/tmp/top_source.csv")))
if rows:
    hdr = rows[0]
    def col(*names):
        for n in names:
            for i, h in enumerate(hdr):
                if n.lower() in h.lower():
                    return i
        return None
    samp = col("Warp Stall Sampling (All", "# Samples", "Sampling")
    if samp is not None:
        body = [r for r in rows[1:] if len(r) > samp and r[samp].replace(",", "").isdigit()]
        body.sort(key=lambda r: -int(r[samp].replace(",", "")))
        with open("$O/${R}_top_source_hot.csv", "w") as f:
            w = csv.writer(f); w.writerow(hdr)
            for r in body[:60]: w.writerow(r)
PY
# 4. training step kernels (loss, BatchNorm statistics / apply / backward, weight gradients, Adam): metrics pass
timeout 900 ncu --metrics $M --clock-control none -k regex:'dice|bn_|adam|maxpool|gate_bwd|wgrad|act_bwd' -c 400 --csv --log-file $O/${R}_ncu_train_kernels.csv \
    python tools/bench_train.py 2 128 128 128 --steps 1 --native-only > $O/${R}_ncu_train.log 2>&1
ls -la $O/${R}_*
