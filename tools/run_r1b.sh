#!/bin/bash
# round-1 GPU pass B: parity tests (log kept), bench with CUDA graph, launch list, small ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -30
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log; tail -c 4500 gpurun_out/bench_cfg2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_cfg5rank.csv python bench.py --steps 2 --warmup 3 --workload cfg5_rank --no-graph --skip-cpu --skip-large > gpurun_out/ncu_launches5.log 2>&1; tail -2 gpurun_out/ncu_launches5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'entropy_' -c 8 -o gpurun_out/prof_r1b_entropy python tools/prof_kernels.py entropy > gpurun_out/ncu_entropy.log 2>&1; tail -2 gpurun_out/ncu_entropy.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'chamfer_nn' -c 8 -o gpurun_out/prof_r1b_chamfer python tools/prof_kernels.py chamfer > gpurun_out/ncu_chamfer.log 2>&1; tail -2 gpurun_out/ncu_chamfer.log
for f in gpurun_out/*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
  ncu -i $f --page details --csv > ${f%.ncu-rep}_details.csv 2>/dev/null
done
du -sm gpurun_out; ls -la gpurun_out
# keep the merge under 64 MiB
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/*.ncu-rep; fi
