#!/bin/bash
# r1o: Chamfer scan with raw shared addressing; pooled-GEMM timing experiments; ncu of pool_sparse at cfg2 size
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python tools/chamfer_probe.py quick > gpurun_out/chamfer_probe.log 2>&1; cat gpurun_out/chamfer_probe.log | tail -30
timeout 300 python tools/time_layer.py > gpurun_out/time_layer.log 2>&1; grep -v Warn gpurun_out/time_layer.log | tail -24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool_sparse' -c 4 -o gpurun_out/prof_r1o_sparse python tools/prof_kernels.py mlp_small > gpurun_out/ncu_sparse.log 2>&1; tail -1 gpurun_out/ncu_sparse.log
ncu -i gpurun_out/prof_r1o_sparse.ncu-rep --page raw --csv > gpurun_out/prof_r1o_sparse_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r1o_sparse.ncu-rep --page source --csv > gpurun_out/prof_r1o_sparse_source.csv 2>/dev/null
rm -f gpurun_out/prof_r1o_sparse.ncu-rep gpurun_out/prof_r1n_chamfer.ncu-rep
