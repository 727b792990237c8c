#!/bin/bash
# Round-2 final evidence pass on one B200 (every command bounded by `timeout`): GPU suite, smoke, both bench arms, the cfg-5
# shard line, CUPTI step profiles, ncu launch list + DRAM traffic of the eager step, memcheck of one step.
R=r2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${R}_smi.log 2>&1
timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${R}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${R}_pytest_gpu.log; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_ref.json.log 2>gpurun_out/${R}_bench_ref.err
PCUDA_BENCH_WATCHDOG=500 timeout 560 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_cfg2.json.log 2>gpurun_out/${R}_bench_cfg2.err; echo "bench exit $?"; tail -c 150 gpurun_out/${R}_bench_cfg2.json.log; grep "bench rank" gpurun_out/${R}_bench_cfg2.err | tail -12
timeout 200 python bench.py --steps 20 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large --skip-scale --skip-eager > gpurun_out/${R}_bench_cfg5rank.json.log 2>/dev/null
timeout 120 python tools/step_profile.py cfg2 > gpurun_out/${R}_step_profile_cfg2_warm.txt 2>&1
timeout 120 python tools/step_profile.py cfg5_rank > gpurun_out/${R}_step_profile_cfg5rank_warm.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv --log-file gpurun_out/${R}_traffic_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-large --skip-scale --skip-eager > gpurun_out/${R}_ncu_traffic.log 2>&1; tail -c 120 gpurun_out/${R}_ncu_traffic.log
timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python tools/eager_step.py cfg2 > gpurun_out/${R}_memcheck_step.txt 2>&1; tail -3 gpurun_out/${R}_memcheck_step.txt
du -sm gpurun_out
