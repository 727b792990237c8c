#!/bin/bash
# Full evidence pass for profiles/: tests, bench lines, ncu launch lists, ncu --set full of every hot kernel, warm step profiles.
R=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log; tail -c 150 gpurun_out/bench_cfg2.log
timeout 600 python bench.py --steps 20 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large > gpurun_out/bench_cfg5rank.log 2>&1; tail -c 150 gpurun_out/bench_cfg5rank.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 300 python tools/step_profile.py cfg2 > gpurun_out/step_profile_cfg2.log 2>&1
timeout 300 python tools/step_profile.py cfg5_rank > gpurun_out/step_profile_cfg5.log 2>&1
timeout 300 python tools/tc_stress.py > gpurun_out/tc_stress.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_cfg5rank.csv python bench.py --steps 2 --warmup 3 --workload cfg5_rank --no-graph --skip-cpu --skip-large > gpurun_out/ncu_launches5.log 2>&1; tail -1 gpurun_out/ncu_launches5.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv --log-file gpurun_out/traffic_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large > gpurun_out/ncu_traffic.log 2>&1; tail -1 gpurun_out/ncu_traffic.log | cut -c1-120
timeout 300 python tools/time_layer.py > gpurun_out/time_layer.log 2>&1; grep -c "dbg=" gpurun_out/time_layer.log
timeout 300 python tools/chamfer_probe.py quick > gpurun_out/chamfer_probe.log 2>&1; tail -2 gpurun_out/chamfer_probe.log
(timeout 60 ./tools/micro/atomic_tail; timeout 60 ./tools/micro/carveout_switch) > gpurun_out/micro.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ws_kernel|pt_kernel' -c 12 -o gpurun_out/prof_${R}_mlp python tools/prof_kernels.py mlp > gpurun_out/ncu_mlp.log 2>&1; tail -1 gpurun_out/ncu_mlp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'chamfer_nn' -c 4 -o gpurun_out/prof_${R}_chamfer python tools/prof_kernels.py chamfer > gpurun_out/ncu_chamfer.log 2>&1; tail -1 gpurun_out/ncu_chamfer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'entropy' -c 8 -o gpurun_out/prof_${R}_entropy python tools/prof_kernels.py entropy > gpurun_out/ncu_entropy.log 2>&1; tail -1 gpurun_out/ncu_entropy.log
for f in gpurun_out/prof_${R}_mlp.ncu-rep gpurun_out/prof_${R}_chamfer.ncu-rep gpurun_out/prof_${R}_entropy.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
rm -f gpurun_out/prof_${R}_chamfer.ncu-rep gpurun_out/prof_${R}_entropy.ncu-rep gpurun_out/prof_r1[hikl]*.ncu-rep
du -sm gpurun_out
