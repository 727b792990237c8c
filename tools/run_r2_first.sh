#!/bin/bash
# round 2, first GPU pass: new parity tests (measured errors printed), the whole GPU suite, both bench arms
mkdir -p gpurun_out
python -m pytest tests/test_gpu_step_parity.py -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/r2_step_parity.log 2>&1
echo "step parity rc=$?" 
tail -5 gpurun_out/r2_step_parity.log
python -m pytest tests/test_gpu_step_parity.py -q -m gpu -s --no-header -p no:cacheprovider > gpurun_out/r2_step_parity_all.log 2>&1
tail -30 gpurun_out/r2_step_parity_all.log
python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_step_parity.py > gpurun_out/r2_pytest_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2_bench_ref.log 2>&1; tail -c 600 gpurun_out/r2_bench_ref.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg2.log 2>gpurun_out/r2_bench_cfg2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_cfg2.log
