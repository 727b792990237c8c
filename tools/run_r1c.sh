#!/bin/bash
# round-1 GPU pass C: tensor-core MLP self-test (each piece isolated), then tests + bench
mkdir -p gpurun_out
timeout 1500 python tools/tc_selftest.py > gpurun_out/tc_selftest.log 2>&1; echo "selftest exit $?" >> gpurun_out/tc_selftest.log
cat gpurun_out/tc_selftest.log | tail -50
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -10
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log; tail -c 3500 gpurun_out/bench_cfg2.log
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large > gpurun_out/bench_cfg5rank.log 2>&1; tail -c 2500 gpurun_out/bench_cfg5rank.log
