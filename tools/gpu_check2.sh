#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf > $O/k_tests.log 2>&1; tail -15 $O/k_tests.log
timeout 500 python bench.py --steps 6 --warmup 3 > $O/k_bench.json 2> $O/k_bench.err; tail -3 $O/k_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    line = [l for l in open("gpurun_out/k_bench.json") if l.startswith("{")][-1]
    d = json.loads(line)
    print("bench:", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["parity"]["max_abs_err_logits"], d["parity"]["argmax_flips_margin_gt_1e-4"])
    print("train:", d.get("train"))
except Exception as e: print("bench parse", e)
PY
VSSEG_SW_STREAMS=2 timeout 300 python bench.py --steps 6 --warmup 3 --no-train > $O/k_bench_s2.json 2> $O/k_bench_s2.err; tail -3 $O/k_bench_s2.err | cut -c1-300
python - <<'PY'
import json
try:
    line = [l for l in open("gpurun_out/k_bench_s2.json") if l.startswith("{")][-1]
    d = json.loads(line)
    print("streams=2:", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["parity"]["max_abs_err_logits"], d["parity"]["argmax_flips_margin_gt_1e-4"])
except Exception as e: print("bench s2 parse", e)
PY
