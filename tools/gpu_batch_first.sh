#!/bin/bash
# A/B on one B200: first ResidualUnit batched over the windows of a group (VSSEG_SW_BATCH_FIRST=1), fused level-2
# attention gate (VSSEG_FUSE_GATE=1) with 1 or 4 channel groups loaded ahead (variant library), then a short bench.
# (the variant libraries come from: tools/build_variant.sh gb1 -DVSSEG_GATE_BATCH=1 / gb4 -DVSSEG_GATE_BATCH=4; at the time the in-tree default was 1)
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -q --no-header -rf -k "batch_first or window_group_plan" > $O/a_tests.log 2>&1; tail -4 $O/a_tests.log
run() { # name, env...
  local n=$1; shift
  env "$@" PROFILE_GROUP=8 timeout 150 python tools/profile_plan.py $O/a_pp_$n.tsv > /dev/null 2> $O/a_pp_$n.err
  echo "== $n: $(tail -1 $O/a_pp_$n.tsv | cut -f5)  ms per group"; grep -E "^(enc0.unit[01]|dec1.att.conv2|dec1.att.gate)" $O/a_pp_$n.tsv | cut -f1,5 | tr '\n' ' ' | cut -c1-400; echo
}
run base VSSEG_SW_BATCH_FIRST=0
run bf VSSEG_SW_BATCH_FIRST=1
run bf_fg VSSEG_SW_BATCH_FIRST=1 VSSEG_FUSE_GATE=1
run bf_fg_gb4 VSSEG_SW_BATCH_FIRST=1 VSSEG_FUSE_GATE=1 VSSEG_LIB_PATH=$PWD/vs_seg_b200/variants/libvsseg_b200_gb4.so
VSSEG_SW_BATCH_FIRST=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline > $O/a_bench_bf.json 2> $O/a_bench_bf.err; tail -2 $O/a_bench_bf.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/a_bench_bf.json"))
    print("batch-first bench:", round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d["parity"]["max_abs_err_logits"], d["parity"]["argmax_flips_margin_gt_1e-4"], d["clocks"])
except Exception as e: print("bench parse", e)
PY
