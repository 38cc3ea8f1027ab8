#!/bin/bash
# Final check of the round on ONE B200: whole GPU suite, the bench line, then the ncu passes over one window group
# (per-launch metrics; --set full of the longest launch).  Most important first: the call may be cut by the budget.
O=gpurun_out; R=r02b; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --no-header -rf > $O/${R}_tests.log 2>&1; tail -4 $O/${R}_tests.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${R}_bench_1gpu.json 2> $O/${R}_bench_1gpu.err; tail -c 300 $O/${R}_bench_1gpu.err
python -c "
import json; d=json.load(open('$O/${R}_bench_1gpu.json')); p=d['parity']; print('bench', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'parity', p['max_abs_err_logits'], p['argmax_flips_margin_gt_1e-4'], d['clocks'], d['roofline']['launch'], round(d['roofline']['frac'],3), 'train', d['train']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])"
PROFILE_GROUP=8 timeout 150 python tools/profile_plan.py $O/${R}_plan_profile_group8.tsv > /dev/null 2>&1; tail -1 $O/${R}_plan_profile_group8.tsv
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__issue_active.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic
export PROFILE_GROUP=8
timeout 240 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/plan_metrics python tools/ncu_plan.py $O/${R}_plan_steps.json > $O/${R}_ncu_plan.log 2>&1
ncu -i /tmp/plan_metrics.ncu-rep --page raw --csv > $O/${R}_ncu_group_metrics_raw.csv 2>/dev/null
python tools/ncu_summarise.py $O/${R}_ncu_group_metrics_raw.csv $O/${R}_plan_steps.json $O/${R}_ncu_group_summary.json
SKIP=$(python - <<PY
import json
s = json.load(open("$O/${R}_ncu_group_summary.json"))["launches"]
i = max(range(len(s)), key=lambda k: s[k].get("dur_us") or 0)
open("$O/${R}_top_launch.txt", "w").write(f"{s[i]['step']} index={i} dur_us={s[i]['dur_us']} kernel={s[i]['kernel']}\n")
print(i)
PY
)
cat $O/${R}_top_launch.txt
timeout 200 ncu --set full --import-source on --clock-control none --profile-from-start off --launch-skip $SKIP --launch-count 1 -f -o /tmp/top_full python tools/ncu_plan.py /tmp/steps2.json >> $O/${R}_ncu_plan.log 2>&1
ncu -i /tmp/top_full.ncu-rep --page raw --csv > $O/${R}_top_full_raw.csv 2>/dev/null
ncu -i /tmp/top_full.ncu-rep --page details > $O/${R}_top_details.txt 2>/dev/null
ls -la $O/${R}_* | cut -c30-
