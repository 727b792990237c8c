#!/bin/bash
# r1p: head-list pool_sparse, four-branch step: tests, bench lines, step profiles, ncu traffic pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log; tail -c 200 gpurun_out/bench_cfg2.log
timeout 600 python bench.py --steps 20 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large > gpurun_out/bench_cfg5rank.log 2>&1; tail -c 200 gpurun_out/bench_cfg5rank.log
timeout 300 python tools/step_profile.py cfg2 > gpurun_out/step_profile_cfg2.log 2>&1
timeout 300 python tools/step_profile.py cfg5_rank > gpurun_out/step_profile_cfg5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv --log-file gpurun_out/traffic_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large > gpurun_out/ncu_traffic.log 2>&1; tail -1 gpurun_out/ncu_traffic.log | cut -c1-200
