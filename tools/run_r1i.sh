#!/bin/bash
# round-1 GPU pass I: fused FC heads, deterministic stats, 16-warp packed forward epilogues
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -12
timeout 300 python tools/tc_stress.py > gpurun_out/tc_stress.log 2>&1; tail -14 gpurun_out/tc_stress.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log; tail -c 300 gpurun_out/bench_cfg2.log
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large > gpurun_out/bench_cfg5rank.log 2>&1; tail -c 300 gpurun_out/bench_cfg5rank.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_cfg5rank.csv python bench.py --steps 2 --warmup 3 --workload cfg5_rank --no-graph --skip-cpu --skip-large > gpurun_out/ncu_launches5.log 2>&1; tail -2 gpurun_out/ncu_launches5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ws_kernel|pt_kernel' -c 12 -o gpurun_out/prof_r1i_mlp python tools/prof_kernels.py mlp > gpurun_out/ncu_mlp.log 2>&1; tail -2 gpurun_out/ncu_mlp.log
for f in gpurun_out/prof_r1i*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
rm -f gpurun_out/prof_r1h*.ncu-rep
du -sm gpurun_out
