#!/bin/bash
# round 2: the whole GPU suite + smoke, logs into gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -s --no-header -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1
echo "gpu suite rc=$?"
grep -oE "^step .*|[.F]step .*" gpurun_out/r2_pytest_gpu.log | cut -c1-300 | head -20
grep -E "^E  " gpurun_out/r2_pytest_gpu.log | cut -c1-500 | head -30
tail -6 gpurun_out/r2_pytest_gpu.log | cut -c1-300
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
