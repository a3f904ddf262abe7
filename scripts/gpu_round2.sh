#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv" -p no:cacheprovider > gpurun_out/k_conv.log 2>&1
echo "conv exit $?" > gpurun_out/summary.txt
timeout -s KILL 1500 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider > gpurun_out/e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt
tail -n 5 gpurun_out/k_conv.log; tail -n 30 gpurun_out/e2e.log
cat gpurun_out/summary.txt
