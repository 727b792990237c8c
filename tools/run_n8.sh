#!/bin/bash
# 8-GPU bench lines (weak scaling): cfg2 and cfg5 per-rank shard
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 --skip-large > gpurun_out/bench_cfg2_n8.log 2>&1; echo "exit $?" >> gpurun_out/bench_cfg2_n8.log; tail -c 200 gpurun_out/bench_cfg2_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --workload cfg5_rank --skip-large > gpurun_out/bench_cfg5rank_n8.log 2>&1; echo "exit $?" >> gpurun_out/bench_cfg5rank_n8.log; tail -c 200 gpurun_out/bench_cfg5rank_n8.log
