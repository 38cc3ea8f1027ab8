#!/bin/bash
# One B200: per-layer tile sweep (tools/autotune_tiles.py), atomic multi-window gate+logits launch, gate batch 8.
# (variant library: tools/build_variant.sh gb8 -DVSSEG_GATE_BATCH=8)
O=gpurun_out; mkdir -p $O
run() { local n=$1; shift
  env "$@" PROFILE_GROUP=8 timeout 150 python tools/profile_plan.py $O/b_pp_$n.tsv > /dev/null 2> $O/b_pp_$n.err
  echo "== $n: $(tail -1 $O/b_pp_$n.tsv | cut -f5)  ms per group"; grep -E "gate" $O/b_pp_$n.tsv | cut -f1,5 | tr '\n' ' ' | cut -c1-500; echo; }
run bf_atomic VSSEG_SW_BATCH_FIRST=1 VSSEG_SW_ATOMIC=1
run bf_fg_gb8 VSSEG_SW_BATCH_FIRST=1 VSSEG_FUSE_GATE=1 VSSEG_LIB_PATH=$PWD/vs_seg_b200/variants/libvsseg_b200_gb8.so
VSSEG_SW_BATCH_FIRST=1 PROFILE_GROUP=8 timeout 400 python tools/autotune_tiles.py $O/b_autotune.tsv 2> $O/b_autotune.err | cut -c1-330
tail -3 $O/b_autotune.err
