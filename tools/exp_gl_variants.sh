#!/bin/bash
# gate_logits_kernel variants (lines per CTA x CTAs per SM): parity test + per-launch time inside a window group
O=gpurun_out; mkdir -p $O
for v in 34_2_1_2 34_3_1_2 34_2_1_2_112 34_3_1_2_112 34_3_1_2_104 66_3_1_1; do
  if [ $v = default ]; then unset VSSEG_LIB_PATH; else export VSSEG_LIB_PATH=$PWD/build/libvsseg_gl_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q --no-header -k "gate_logits" 2>&1 | tail -1
  PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/gl_$v.tsv > /dev/null 2>&1
  echo "$v: $(grep 'gate+logits@w0' $O/gl_$v.tsv | cut -f1,5,8) | $(grep TOTAL $O/gl_$v.tsv | cut -f1,5)"
done
