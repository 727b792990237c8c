#!/bin/bash
# quick GPU validation: tests, smoke, short bench at both workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit|Error" gpurun_out/pytest_gpu.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log; tail -c 1500 gpurun_out/bench_cfg2.log
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large > gpurun_out/bench_cfg5rank.log 2>&1; tail -c 900 gpurun_out/bench_cfg5rank.log
