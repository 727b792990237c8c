#!/bin/bash
# r1n: seeded prefilter Chamfer: tests, variant probe, ncu of the new kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python tools/chamfer_probe.py > gpurun_out/chamfer_probe.log 2>&1; grep "var0\|var4128\|var204128" gpurun_out/chamfer_probe.log | grep "N4096\|N16384\|N1024"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'chamfer_nn' -c 4 -o gpurun_out/prof_r1n_chamfer python tools/prof_kernels.py chamfer > gpurun_out/ncu_chamfer.log 2>&1; tail -1 gpurun_out/ncu_chamfer.log
ncu -i gpurun_out/prof_r1n_chamfer.ncu-rep --page raw --csv > gpurun_out/prof_r1n_chamfer_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r1n_chamfer.ncu-rep --page source --csv > gpurun_out/prof_r1n_chamfer_source.csv 2>/dev/null
rm -f gpurun_out/prof_r1m_chamfer.ncu-rep
