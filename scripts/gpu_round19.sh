#!/bin/bash
mkdir -p gpurun_out
for nw in 8 4; do
LUMI_ROI_NW=$nw timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_golden_fixtures.py -m gpu -q -x -k "roi or golden" -p no:cacheprovider > gpurun_out/k_roi_nw$nw.log 2>&1
echo "roi nw$nw exit $?" >> gpurun_out/summary.txt
done
for nw in 8 4 8 4; do
LUMI_ROI_NW=$nw timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_nw$nw.json 2> gpurun_out/bench_r50_nw$nw.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r50_nw$nw.json')); print('nw$nw', round(d['value'],1), round(d['ms_per_step'],3), round(d['category_ms_per_step']['roi_pool'],3))"
done
timeout -s KILL 1500 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider > gpurun_out/e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt
tail -n 3 gpurun_out/k_roi_nw*.log; tail -n 4 gpurun_out/e2e.log; cat gpurun_out/summary.txt
