#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --skip-large > gpurun_out/r2_bench_cfg2_n$N.log 2> gpurun_out/r2_bench_cfg2_n$N.err
echo "bench N=$N rc=$?"; tail -c 800 gpurun_out/r2_bench_cfg2_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_cfg2_n$N.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','exchange','gpu_launches_per_step','n_gpus')}, d['e2e'])
    for k,v in d['scale'].items():
        if 'sizes' in v: print(k, {n:round(x['tpairs_s'],2) for n,x in v['sizes'].items()})
        else: print(k, {a:v.get(a) for a in ('value','ms_per_step','per_rank_batch','exchange')}, v.get('e2e'))
except Exception as e: print('parse failed', e)
PY
