#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log; tail -c 6000 gpurun_out/bench.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-graph --skip-cpu --skip-large > gpurun_out/bench_nograph.log 2>&1; tail -c 1500 gpurun_out/bench_nograph.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'entropy_|chamfer_nn' -c 24 -o gpurun_out/prof_r1_entropy_chamfer python tools/prof_kernels.py > gpurun_out/ncu1.log 2>&1; tail -3 gpurun_out/ncu1.log
ls -la gpurun_out
