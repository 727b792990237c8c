#!/bin/bash
# r1m: prefiltered Chamfer kernel + list-based pool_sparse: tests, variant probe, bench, traffic pass, ncu of the new kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python tools/chamfer_probe.py > gpurun_out/chamfer_probe.log 2>&1; tail -3 gpurun_out/chamfer_probe.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log; tail -c 300 gpurun_out/bench_cfg2.log
timeout 300 python tools/step_profile.py cfg2 > gpurun_out/step_profile_cfg2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv --log-file gpurun_out/traffic_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large > gpurun_out/ncu_traffic.log 2>&1; tail -1 gpurun_out/ncu_traffic.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'chamfer_nn' -c 4 -o gpurun_out/prof_r1m_chamfer python tools/prof_kernels.py chamfer > gpurun_out/ncu_chamfer.log 2>&1; tail -1 gpurun_out/ncu_chamfer.log
ncu -i gpurun_out/prof_r1m_chamfer.ncu-rep --page raw --csv > gpurun_out/prof_r1m_chamfer_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r1m_chamfer.ncu-rep --page source --csv > gpurun_out/prof_r1m_chamfer_source.csv 2>/dev/null
du -sm gpurun_out
