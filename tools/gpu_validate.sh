#!/bin/bash
# One B200: the whole GPU test suite, the group profile and the bench line with the current defaults.
O=gpurun_out; T=${1:-c}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --no-header -rf > $O/${T}_tests.log 2>&1; tail -6 $O/${T}_tests.log
PROFILE_GROUP=8 timeout 150 python tools/profile_plan.py $O/${T}_pp.tsv > /dev/null 2> $O/${T}_pp.err; echo "group of 8: $(tail -1 $O/${T}_pp.tsv | cut -f5) ms"
timeout 500 python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -2 $O/${T}_bench.err | cut -c1-300
python - $O/${T}_bench.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    p = d.get("parity") or {}
    print("bench:", round(d["value"], 1), "patches/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"], 1), "; parity", p.get("max_abs_err_logits"), p.get("argmax_flips_margin_gt_1e-4"), "; clocks", d["clocks"], "; roofline", d["roofline"]["launch"], round(d["roofline"]["frac"], 3), "; train", (d.get("train") or {}).get("ms_per_step"))
except Exception as e: print("bench parse", e)
PY
