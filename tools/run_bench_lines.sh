#!/bin/bash
# tests + the bench lines profiles/ quotes (cfg2 headline with CPU baseline and roofline-sized shapes, cfg5 per-rank shard, reference arm)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_cfg2.log
timeout 600 python bench.py --steps 20 --warmup 3 --workload cfg5_rank --skip-cpu --skip-large > gpurun_out/bench_cfg5rank.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 300 python tools/step_profile.py cfg2 > gpurun_out/step_profile_cfg2.log 2>&1
timeout 300 python tools/step_profile.py cfg5_rank > gpurun_out/step_profile_cfg5.log 2>&1
python - <<'PY'
import json
for fn in ('bench_cfg2.log','bench_cfg5rank.log','bench_ref.log'):
    for l in open('gpurun_out/'+fn):
        if l.startswith('{'):
            d=json.loads(l); print(fn,'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'launches',d.get('gpu_launches_per_step'))
PY
