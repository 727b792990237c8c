#!/bin/bash
# round-1 GPU pass H (re-entry): state check — tests, bench, launch lists, ncu full of every hot kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/smi.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log; tail -c 600 gpurun_out/bench_cfg2.log
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large > gpurun_out/bench_cfg5rank.log 2>&1; tail -c 300 gpurun_out/bench_cfg5rank.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -c 300 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_cfg5rank.csv python bench.py --steps 2 --warmup 3 --workload cfg5_rank --no-graph --skip-cpu --skip-large > gpurun_out/ncu_launches5.log 2>&1; tail -2 gpurun_out/ncu_launches5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ws_kernel|pt_kernel' -c 12 -o gpurun_out/prof_r1h_mlp python tools/prof_kernels.py mlp > gpurun_out/ncu_mlp.log 2>&1; tail -2 gpurun_out/ncu_mlp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'chamfer_nn' -c 4 -o gpurun_out/prof_r1h_chamfer python tools/prof_kernels.py chamfer > gpurun_out/ncu_chamfer.log 2>&1; tail -2 gpurun_out/ncu_chamfer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'entropy' -c 8 -o gpurun_out/prof_r1h_entropy python tools/prof_kernels.py entropy > gpurun_out/ncu_entropy.log 2>&1; tail -2 gpurun_out/ncu_entropy.log
for f in gpurun_out/prof_r1h*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
du -sm gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/prof_r1h_entropy.ncu-rep gpurun_out/prof_r1h_chamfer.ncu-rep; fi
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/*.ncu-rep; fi
