#!/bin/bash
# N-GPU scaling check (run with gpurun --gpus N): bench at N ranks, peer blend and NCCL-reduce variants
N=${1:-8}; O=gpurun_out; mkdir -p $O
for mode in peer nccl; do
  if [ $mode = nccl ]; then export VSSEG_SW_PEER=0; else export VSSEG_SW_PEER=1; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-train > $O/s_bench_n${N}_$mode.json 2> $O/s_bench_n${N}_$mode.err
  python - <<PY
import json
try:
    line = [l for l in open("$O/s_bench_n${N}_$mode.json") if l.startswith("{")][-1]
    d = json.loads(line)
    print("$mode N=$N", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("$mode N=$N parse error", e)
PY
  tail -3 $O/s_bench_n${N}_$mode.err | cut -c1-300
done
