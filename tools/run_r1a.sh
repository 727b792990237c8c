#!/bin/bash
# round-1 GPU pass A: parity tests, smoke, bench (cfg2 + cfg5_rank), launch list, ncu full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log; tail -c 5000 gpurun_out/bench_cfg2.log
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large > gpurun_out/bench_cfg5rank.log 2>&1; tail -c 3000 gpurun_out/bench_cfg5rank.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -c 1000 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg5rank.csv python bench.py --steps 2 --warmup 3 --workload cfg5_rank --no-graph --skip-cpu --skip-large > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'entropy_|chamfer_nn|mlp_fwd' -c 30 -o gpurun_out/prof_r1a python tools/prof_kernels.py > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
