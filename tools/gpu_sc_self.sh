#!/bin/bash
# A/B on one B200: shortcut MMAs on the centre-tap stages (VSSEG_TC_SC_SELF) vs separate shortcut stages, tile sweep of the decoder units
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --no-header -rf -k "tcgen05 or unet_eval or window_group or sliding_window or captured or batch_first" > $O/d_tests.log 2>&1; tail -6 $O/d_tests.log
run() { local n=$1; shift
  env "$@" PROFILE_GROUP=8 timeout 150 python tools/profile_plan.py $O/d_pp_$n.tsv > /dev/null 2> $O/d_pp_$n.err
  echo "== $n: $(tail -1 $O/d_pp_$n.tsv | cut -f5)  ms per group"; grep -E "^dec[1-4].unit0" $O/d_pp_$n.tsv | cut -f1,5 | tr '\n' ' '; echo; }
run sc0 VSSEG_TC_SC_SELF=0
run sc1 VSSEG_TC_SC_SELF=1
run sc1_nohint VSSEG_TC_SC_SELF=1 VSSEG_TC_HINTS=0
AUTOTUNE_ONLY=dec1.unit0,dec2.unit0,dec3.unit0,dec4.unit0 AUTOTUNE_XT=1,2,4,8,16,32 AUTOTUNE_YT=1,2,4,8,16 AUTOTUNE_NST=0,2,3 PROFILE_GROUP=8 timeout 300 python tools/autotune_tiles.py $O/d_autotune_sc.tsv 2> $O/d_autotune_sc.err | cut -c1-400
