#!/bin/bash
# 2-GPU check on the GPU box: peer-memory blend test + 2-rank bench (peer blend and NCCL-reduce variants)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_round2.py -m gpu -q --no-header -rf -k "nccl or gate_logits" > gpurun_out/r2_t3.log 2>&1
tail -15 gpurun_out/r2_t3.log
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 1500 gpurun_out/r2_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n$N.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['parity'])"
VSSEG_SW_PEER=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n${N}_nccl.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n${N}_nccl.json')); print('nccl-reduce', d['value'], d['ms_per_step'], d['e2e'])"
