#!/bin/bash
# One-shot round-2 check on ONE B200: TS microbenchmark, GPU tests, bench (default and forced-TS plans), profiles,
# gradient diagnostics, block bench.  Everything lands in gpurun_out/.
O=gpurun_out; mkdir -p $O
(cd tools/ubench && timeout 60 ./mma_ts) > $O/c_ubench_ts.log 2>&1; tail -8 $O/c_ubench_ts.log
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf > $O/c_tests.log 2>&1; tail -12 $O/c_tests.log
timeout 400 python bench.py --steps 6 --warmup 3 > $O/c_bench.json 2> $O/c_bench.err; tail -3 $O/c_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c_bench.json"))
    print("default:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["max_abs_err_logits"], d["parity"]["argmax_flips_margin_gt_1e-4"], d.get("train"))
except Exception as e: print("bench parse", e)
PY
PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/c_pp_default.tsv > /dev/null 2> $O/c_pp_default.err; tail -1 $O/c_pp_default.tsv
VSSEG_TC_TS=2 PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/c_pp_ts.tsv > /dev/null 2> $O/c_pp_ts.err; tail -1 $O/c_pp_ts.tsv; tail -3 $O/c_pp_ts.err
VSSEG_TC_TS=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --no-header -rf -k "tcgen05 or unet_eval or window_group or sliding_window" > $O/c_tests_ts.log 2>&1; tail -8 $O/c_tests_ts.log
VSSEG_TC_TS=2 timeout 300 python bench.py --steps 6 --warmup 3 --no-train > $O/c_bench_ts.json 2> $O/c_bench_ts.err; tail -3 $O/c_bench_ts.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c_bench_ts.json"))
    print("forced TS:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["max_abs_err_logits"], d["parity"]["argmax_flips_margin_gt_1e-4"])
except Exception as e: print("bench ts parse", e)
PY
timeout 300 python tools/train_diag.py 2 1 64 64 16 > $O/c_train_diag.log 2>&1; tail -2 $O/c_train_diag.log
timeout 300 python tools/bench_block.py > $O/c_block.json 2> $O/c_block.err; tail -c 1500 $O/c_block.json; tail -3 $O/c_block.err
