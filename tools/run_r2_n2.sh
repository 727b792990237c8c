#!/bin/bash
# 2-GPU pass: multi-GPU tests, then the bench at N=2 (both arms bounded by `timeout`)
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/r2_multi_n2.log 2>&1
echo "multi rc=$?"; tail -5 gpurun_out/r2_multi_n2.log | cut -c1-300
PCUDA_BENCH_WATCHDOG=250 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --skip-large --skip-eager > gpurun_out/r2_bench_cfg2_n2.log 2> gpurun_out/r2_bench_cfg2_n2.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_cfg2_n2.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_cfg2_n2.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','exchange','gpu_launches_per_step')}, d['e2e'])
    print(json.dumps(d['scale'])[:1500])
except Exception as e: print('parse failed', e)
PY
