#!/bin/bash
# 8-GPU checks of the one-sided merge: union parity (tools/merge_check.py) and the bench line
# (per-rank c2 shard, final merge push, c5 block)
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/merge_check.py --scale 0.2 --haplotypes 80 > gpurun_out/merge_check_8gpu.log 2>&1
tail -5 gpurun_out/merge_check_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -5 gpurun_out/bench_8gpu.err
cat gpurun_out/bench_8gpu.json | cut -c 1-1500
