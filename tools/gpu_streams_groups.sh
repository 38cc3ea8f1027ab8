#!/bin/bash
# One B200: whole GPU suite, then bench A/B of one vs two streams and window groups of 8 vs 16
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --no-header -rf > $O/f_tests.log 2>&1; tail -6 $O/f_tests.log
b() { local n=$1; shift
  env "$@" timeout 200 python bench.py --steps 12 --warmup 4 --no-train --no-cpu-baseline > $O/f_bench_$n.json 2> $O/f_bench_$n.err
  python -c "
import json; d=json.load(open('$O/f_bench_$n.json')); print('$n bench', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])" || tail -3 $O/f_bench_$n.err; }
b default
b streams1 VSSEG_SW_STREAMS=1
b group16 VSSEG_SW_GROUP=16
b group16_s1 VSSEG_SW_GROUP=16 VSSEG_SW_STREAMS=1
b default2
