#!/bin/bash
# compute-sanitizer passes over the small end-to-end paths (sizes kept small: the sanitizer slows kernels 10-100x):
#   memcheck  smoke()  (sliding window, 2x2x1 windows of 64x64x16, tensor-core + generic kernels, CUDA graph)
#   memcheck  one native training step at 64x64x16 (forward, loss, backward, FusedAdam)
#   racecheck smoke()  (shared-memory hazards of the generic kernels; TMA/tcgen05 traffic is outside its model)
# plus the launch list of smoke() (which kernels a small geometry uses).
O=gpurun_out; T=${1:-san}; mkdir -p $O
export VSSEG_SW_GRAPH=${VSSEG_SW_GRAPH:-0}   # eager launches: the sanitizer attributes errors to a launch site
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python __graft_entry__.py smoke > $O/${T}_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -4 $O/${T}_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python tools/train_diag.py 1 1 64 64 16 > $O/${T}_memcheck_train.log 2>&1; echo "memcheck train rc=$?"; tail -4 $O/${T}_memcheck_train.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python __graft_entry__.py smoke > $O/${T}_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -4 $O/${T}_racecheck_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'vsseg|conv_|gate_|att_|sw_' -c 400 --csv \
    --log-file $O/${T}_launches_smoke.csv python __graft_entry__.py smoke > $O/${T}_smoke_under_ncu.log 2>&1
grep -o 'conv_act8_kernel\|conv_tc_kernel<[0-9, ]*>\|gate_logits_kernel<[0-9]*>\|[a-z0-9_]*_kernel' $O/${T}_launches_smoke.csv | sort | uniq -c | sort -rn | head -20
