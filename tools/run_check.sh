#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit|Error|assert" gpurun_out/pytest_gpu.log | tail -12
timeout 600 python bench.py --steps 50 --warmup 5 --skip-large --skip-cpu > gpurun_out/bench_cfg2_quick.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2_quick.log; tail -c 150 gpurun_out/bench_cfg2_quick.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_cfg2_quick.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches_per_step'])
PY
