#!/bin/bash
# Round 2, call B: validate the second batch of kernels (row-walk ROI, NMS mask grid, chunk schedule, 16-warp epilogue, graphs).
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not epi16" --timeout 400 --timeout-method=thread > gpurun_out/b_pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/b_summary.txt
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/b_bench_r50.json 2> gpurun_out/b_bench_r50.err
echo "bench r50 exit $?" >> gpurun_out/b_summary.txt
LUMI_ROI_KERNEL=cols timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_r50_roi_cols.json 2>/dev/null
timeout -s KILL 300 python bench.py --workload ssd --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_ssd.json 2>/dev/null
timeout -s KILL 300 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_r101.json 2>/dev/null
# D1 chunk schedule: accuracy (BASELINE-config parity report) and speed (bench --layers) at 2 / 1 stages per chunk
for t in 2 1; do
  LUMI_CONV_CHUNK_TAIL=$t LUMI_PARITY_TAG=_tail$t timeout -s KILL 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -p no:cacheprovider -k "config2 or config4" --timeout 600 --timeout-method=thread > gpurun_out/b_pytest_tail$t.log 2>&1
  echo "pytest tail$t exit $?" >> gpurun_out/b_summary.txt
  LUMI_CONV_CHUNK_TAIL=$t timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/b_bench_r50_tail$t.json 2>/dev/null
done
LUMI_CONV_CHUNK_TAIL=1 timeout -s KILL 300 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_r101_tail1.json 2>/dev/null
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:roi_pool -c 1 -f -o gpurun_out/b_prof_roi python bench.py --ncu-range --ncu-unpiped --no-cpu-baseline > gpurun_out/b_ncu_roi.log 2>&1
echo "ncu roi exit $?" >> gpurun_out/b_summary.txt
# LAST (new, untested kernels: a hang must not cost the rest of the call): 16-epilogue-warp conv kernels
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "epi16" --timeout 120 --timeout-method=thread > gpurun_out/b_pytest_epi16.log 2>&1
echo "pytest epi16 exit $?" >> gpurun_out/b_summary.txt
LUMI_CONV_EPI16=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/b_bench_r50_epi16.json 2> gpurun_out/b_bench_r50_epi16.err
echo "bench epi16 exit $?" >> gpurun_out/b_summary.txt
LUMI_CONV_EPI16=1 LUMI_PARITY_TAG=_epi16 timeout -s KILL 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -p no:cacheprovider -k "config2" --timeout 500 --timeout-method=thread > gpurun_out/b_pytest_epi16_parity.log 2>&1
echo "pytest epi16 parity exit $?" >> gpurun_out/b_summary.txt
tail -n 12 gpurun_out/b_pytest_gpu.log
tail -n 6 gpurun_out/b_pytest_epi16.log
cat gpurun_out/b_summary.txt
python - <<'PY'
import json
for wl in ('r50','r50_roi_cols','r50_tail2','r50_tail1','r50_epi16','ssd','r101','r101_tail1'):
    try:
        d=json.load(open('gpurun_out/b_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4))
    except Exception as e: print(wl, 'ERR', e)
PY
