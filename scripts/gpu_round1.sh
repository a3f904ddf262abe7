#!/bin/bash
# First GPU contact: per-kernel parity (SIMT + post-processing first, tcgen05 after), then e2e.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tc" -p no:cacheprovider > gpurun_out/k_simt.log 2>&1
echo "simt/postproc exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tc" -p no:cacheprovider > gpurun_out/k_tc.log 2>&1
echo "tc exit $?" >> gpurun_out/summary.txt
timeout -s KILL 1200 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider > gpurun_out/e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/k_simt.log gpurun_out/k_tc.log gpurun_out/e2e.log
cat gpurun_out/summary.txt
