#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -6
timeout 300 python tools/step_profile.py cfg2 > gpurun_out/step_profile_cfg2.log 2>&1; tail -75 gpurun_out/step_profile_cfg2.log
timeout 300 python tools/step_profile.py cfg5_rank > gpurun_out/step_profile_cfg5.log 2>&1; tail -3 gpurun_out/step_profile_cfg5.log
