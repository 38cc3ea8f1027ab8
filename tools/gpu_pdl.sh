#!/bin/bash
# A/B on one B200: programmatic dependent launch of the tensor-core conv kernels (VSSEG_TC_PDL): parity tests, group profile, bench
O=gpurun_out; mkdir -p $O
VSSEG_TC_PDL=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --no-header -rf -k "tcgen05 or unet_eval or window_group or sliding_window or captured or batch_first or two_stream" > $O/e_tests_pdl.log 2>&1; tail -5 $O/e_tests_pdl.log
for p in 0 1 0 1; do
  VSSEG_TC_PDL=$p PROFILE_GROUP=8 timeout 150 python tools/profile_plan.py $O/e_pp_pdl$p.tsv > /dev/null 2> $O/e_pp_pdl$p.err; echo "pdl=$p group of 8: $(tail -1 $O/e_pp_pdl$p.tsv | cut -f5) ms"
  VSSEG_TC_PDL=$p timeout 200 python bench.py --steps 12 --warmup 4 --no-train --no-cpu-baseline > $O/e_bench_pdl$p.json 2> $O/e_bench_pdl$p.err
  python -c "
import json; d=json.load(open('$O/e_bench_pdl$p.json')); print('pdl=$p bench', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done
