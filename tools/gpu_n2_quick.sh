#!/bin/bash
# 2-GPU check: peer-memory blend test (2 NCCL ranks == 1 GPU) + a short 2-rank bench with the full-volume parity leg
mkdir -p gpurun_out
N=${1:-2}
timeout 500 python -m pytest tests/test_gpu_round2.py -m gpu -q --no-header -rf -k "nccl" 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-train > gpurun_out/q_bench_n$N.log 2> gpurun_out/q_bench_n$N.err
python - <<PY
import json
l=[x for x in open('gpurun_out/q_bench_n$N.log') if x.startswith('{')][-1]
d=json.loads(l); open('gpurun_out/q_bench_n$N.json','w').write(l)
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])
PY
