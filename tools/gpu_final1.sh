#!/bin/bash
O=gpurun_out; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; tail -2 $O/r02_bench_1gpu.err | cut -c1-300
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_reference.json 2>/dev/null; tail -c 400 $O/r02_bench_reference.json
timeout 600 ncu --metrics $M --clock-control none -k regex:'dice|bn_|adam|maxpool|gate_bwd|wgrad|act_bwd' -c 400 --csv --log-file $O/r02_ncu_train_kernels.csv python tools/bench_train.py 2 128 128 128 --steps 1 --native-only > $O/r02_ncu_train.log 2>&1
timeout 300 python tools/bench_block.py > $O/r02_block_bench.json 2> $O/r02_block.err
timeout 300 python tools/bench_train.py 2 128 128 128 --steps 5 > $O/r02_train_step.json 2> $O/r02_train_step.err
timeout 900 python tools/train_diag.py 2 1 128 128 128 > $O/r02_train_diag_128.log 2>&1; tail -2 $O/r02_train_diag_128.log | cut -c1-400
timeout 120 python tools/train_diag.py 2 1 64 64 16 > $O/r02_train_diag_64.log 2>&1
(cd tools/ubench && ./mma_ts) > $O/r02_ubench_mma_ts.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; tail -1 $O/r02_smoke.log
