#!/bin/bash
# quick check of a tensor-core kernel change: conv parity tests, group profile (default and forced TS), cycle counters
O=gpurun_out; T=${1:-q}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --no-header -rf -k "tcgen05 or unet_eval or window_group or sliding_window or captured" > $O/${T}_tests.log 2>&1; tail -4 $O/${T}_tests.log
PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/${T}_pp_default.tsv > /dev/null 2> $O/${T}_pp_default.err; tail -1 $O/${T}_pp_default.tsv
VSSEG_TC_TS=2 PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/${T}_pp_ts.tsv > /dev/null 2> $O/${T}_pp_ts.err; tail -1 $O/${T}_pp_ts.tsv
VSSEG_TC_TS=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --no-header -rf -k "tcgen05 or unet_eval or window_group" > $O/${T}_tests_ts.log 2>&1; tail -3 $O/${T}_tests_ts.log
VSSEG_TC_DEBUG=1 timeout 300 python tools/tc_debug.py 8 2> $O/${T}_tcdebug.log > /dev/null
VSSEG_TC_TS=2 VSSEG_TC_DEBUG=1 timeout 300 python tools/tc_debug.py 8 2> $O/${T}_tcdebug_ts.log > /dev/null
