#!/bin/bash
mkdir -p gpurun_out
for v in 2 3 0; do
LUMI_ROI_VARIANT=$v timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py tests/test_golden_fixtures.py -m gpu -q -k "roi" -p no:cacheprovider > gpurun_out/k_roi_v$v.log 2>&1
echo "roi tests v$v exit $?" >> gpurun_out/summary.txt
LUMI_ROI_VARIANT=$v timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_roi_v$v.json 2> gpurun_out/bench_roi_v$v.err
echo "bench v$v exit $?" >> gpurun_out/summary.txt
done
tail -n 3 gpurun_out/k_roi_v*.log; tail -n 5 gpurun_out/e2e.log
python - <<'PY'
import json
for v in (2,3,0):
    try:
        d=json.load(open('gpurun_out/bench_roi_v%d.json'%v)); print(v, d['value'], d['ms_per_step'], d['category_ms_per_step']['roi_pool'])
    except Exception as e: print(v, 'fail', e)
PY
cat gpurun_out/summary.txt
