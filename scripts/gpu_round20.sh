#!/bin/bash
mkdir -p gpurun_out
for m in 3 2 1; do
LUMI_ROI_MODE=$m timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_golden_fixtures.py -m gpu -q -x -k "roi or golden" -p no:cacheprovider > gpurun_out/k_roi_m$m.log 2>&1
echo "roi mode$m exit $?" >> gpurun_out/summary.txt
done
for m in 0 1 2 3 0 2; do
LUMI_ROI_MODE=$m timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_m$m.json 2> gpurun_out/bench_r50_m$m.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r50_m$m.json')); print('mode$m', round(d['value'],1), round(d['ms_per_step'],3), round(d['category_ms_per_step']['roi_pool'],3))"
done
tail -n 2 gpurun_out/k_roi_m*.log; cat gpurun_out/summary.txt
