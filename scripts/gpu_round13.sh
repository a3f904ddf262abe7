#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k roi -p no:cacheprovider > gpurun_out/k_roi.log 2>&1
LUMI_ROI_FLAT=1 timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k roi -p no:cacheprovider > gpurun_out/k_roi_flat.log 2>&1
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err
LUMI_ROI_FLAT=1 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_flat.json 2>> gpurun_out/bench_r50.err
tail -n 3 gpurun_out/k_roi.log gpurun_out/k_roi_flat.log
python -c "
import json
for f in ['bench_r50.json','bench_r50_flat.json']:
    d=json.load(open('gpurun_out/'+f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['category_ms_per_step']['roi_pool'])"
