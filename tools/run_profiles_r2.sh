#!/bin/bash
# Round-2 evidence pass for profiles/ (one B200): tests, smoke, both bench arms, launch lists, DRAM traffic, ncu --set full of
# the hot kernels.  Every command is bounded by `timeout`.
R=r2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${R}_smi.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${R}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${R}_pytest_gpu.log; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_ref.json.log 2>gpurun_out/${R}_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_cfg2.json.log 2>gpurun_out/${R}_bench_cfg2.err; echo "bench exit $?"; tail -c 200 gpurun_out/${R}_bench_cfg2.json.log
timeout 400 python bench.py --steps 20 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large --skip-scale --skip-eager > gpurun_out/${R}_bench_cfg5rank.json.log 2>/dev/null
timeout 300 python tools/step_profile.py cfg2 > gpurun_out/${R}_step_profile_cfg2_warm.txt 2>&1
timeout 300 python tools/step_profile.py cfg5_rank > gpurun_out/${R}_step_profile_cfg5rank_warm.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/${R}_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large --skip-scale --skip-eager > gpurun_out/${R}_ncu_launches.log 2>&1; tail -c 200 gpurun_out/${R}_ncu_launches.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv --log-file gpurun_out/${R}_traffic_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large --skip-scale --skip-eager > gpurun_out/${R}_ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ws_kernel|pt_kernel' -c 12 -o gpurun_out/prof_${R}_mlp python tools/prof_kernels.py mlp > gpurun_out/${R}_ncu_mlp.log 2>&1; tail -1 gpurun_out/${R}_ncu_mlp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'chamfer_nn' -c 4 -o gpurun_out/prof_${R}_chamfer python tools/prof_kernels.py chamfer > gpurun_out/${R}_ncu_chamfer.log 2>&1; tail -1 gpurun_out/${R}_ncu_chamfer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'entropy' -c 8 -o gpurun_out/prof_${R}_entropy python tools/prof_kernels.py entropy > gpurun_out/${R}_ncu_entropy.log 2>&1; tail -1 gpurun_out/${R}_ncu_entropy.log
for f in gpurun_out/prof_${R}_mlp.ncu-rep gpurun_out/prof_${R}_chamfer.ncu-rep gpurun_out/prof_${R}_entropy.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
rm -f gpurun_out/prof_${R}_*.ncu-rep
du -sm gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "point_transform or fused_input_transform or fps or fc_stack_vs_torch" > gpurun_out/r2_memcheck_new_kernels.txt 2>&1; tail -4 gpurun_out/r2_memcheck_new_kernels.txt
