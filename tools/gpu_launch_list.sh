#!/bin/bash
# ncu launch list of the bench command on ONE B200 (the timed region replays CUDA graphs: --graph-profiling node lists
# the kernel nodes one by one; per-launch times are cold-cache and serialised - compare SHARES, not absolutes)
O=gpurun_out; R=${1:-r02b}; mkdir -p $O
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node \
    -k regex:'conv_tc|gate_logits|cin1|att_gate|sw_finalize|smallcout|conv_act8' -c 1000 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > $O/${R}_bench_under_ncu.log 2>&1
wc -l $O/${R}_launches_bench.csv
