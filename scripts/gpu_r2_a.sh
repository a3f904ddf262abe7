#!/bin/bash
# Round 2, call A: full GPU test suite (incl. the BASELINE-config parity tests), clean-exit check of bench.py,
# bench lines for the three workloads, one-step ncu --set full for SSD and R101 (r1 only had R50).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not epi16" --timeout 400 --timeout-method=thread > gpurun_out/a_pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/a_summary.txt
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/a_bench_r50.json 2> gpurun_out/a_bench_r50.err
echo "bench r50 exit $? (must be 0: normal interpreter exit, no os._exit)" >> gpurun_out/a_summary.txt
timeout -s KILL 600 python bench.py --workload ssd --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/a_bench_ssd.json 2> gpurun_out/a_bench_ssd.err
echo "bench ssd exit $?" >> gpurun_out/a_summary.txt
timeout -s KILL 600 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --layers --no-cpu-baseline > gpurun_out/a_bench_r101.json 2> gpurun_out/a_bench_r101.err
echo "bench r101 exit $?" >> gpurun_out/a_summary.txt
# ROI kernel A/B (category_ms_per_step.roi_pool): round-1 cell kernel vs the column-walk kernel at 4 / 8 channels per lane
LUMI_ROI_KERNEL=cells timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_r50_roi_cells.json 2>/dev/null
LUMI_ROI_COLS_CPL=8 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_r50_roi_cols8.json 2>/dev/null
LUMI_CONV_SERPENTINE=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/a_bench_r50_serp.json 2>/dev/null
# CUDA graphs A/B (default off until validated): batch 8 and batch-1 / batch-2 latency
LUMI_GRAPHS=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/a_bench_r50_graphs.json 2> gpurun_out/a_bench_r50_graphs.err
for b in 1 2; do
  timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --per-gpu-batch $b --no-cpu-baseline > gpurun_out/a_bench_r50_b${b}.json 2>/dev/null
  LUMI_GRAPHS=1 timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --per-gpu-batch $b --no-cpu-baseline > gpurun_out/a_bench_r50_b${b}_graphs.json 2>/dev/null
done
for wl in ssd frcnn_r101; do
  timeout -s KILL 1200 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/ncu_$wl python bench.py --workload $wl --ncu-range --ncu-unpiped --no-cpu-baseline > gpurun_out/a_ncu_$wl.log 2>&1
  echo "ncu $wl exit $?" >> gpurun_out/a_summary.txt
  python scripts/ncu_step_summary.py /tmp/ncu_$wl.ncu-rep $wl gpurun_out/a_ncu_$wl > gpurun_out/a_ncu_${wl}_summary.txt 2>&1
done
# source-level profile of the first six tcgen05 conv launches of a step (stem, block1 shortcut / conv1 / conv2 / conv3+res / conv1):
# the short-K, epilogue-bound layers -- read here with `ncu -i ... --page source --csv`
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc_kernel -c 6 -f -o gpurun_out/a_prof_conv_first6 python bench.py --ncu-range --ncu-unpiped --no-cpu-baseline > gpurun_out/a_ncu_conv6.log 2>&1
echo "ncu conv6 exit $?" >> gpurun_out/a_summary.txt
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:roi_pool -c 1 -f -o gpurun_out/a_prof_roi python bench.py --ncu-range --ncu-unpiped --no-cpu-baseline > gpurun_out/a_ncu_roi.log 2>&1
echo "ncu roi exit $?" >> gpurun_out/a_summary.txt
# D1 chunk schedule: accuracy (BASELINE-config parity report) and speed (bench --layers) at 1 / 2 stages per chunk
for t in 2 1; do
  LUMI_CONV_CHUNK_TAIL=$t LUMI_PARITY_TAG=_tail$t timeout -s KILL 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -p no:cacheprovider -k "config2 or config4" --timeout 600 --timeout-method=thread > gpurun_out/a_pytest_tail$t.log 2>&1
  echo "pytest tail$t exit $?" >> gpurun_out/a_summary.txt
  LUMI_CONV_CHUNK_TAIL=$t timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/a_bench_r50_tail$t.json 2>/dev/null
done
# LAST (new, untested kernels: a hang must not cost the rest of the call): 16-epilogue-warp conv kernels
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "epi16" --timeout 120 --timeout-method=thread > gpurun_out/a_pytest_epi16.log 2>&1
echo "pytest epi16 exit $?" >> gpurun_out/a_summary.txt
LUMI_CONV_EPI16=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/a_bench_r50_epi16.json 2> gpurun_out/a_bench_r50_epi16.err
echo "bench epi16 exit $?" >> gpurun_out/a_summary.txt
tail -n 15 gpurun_out/a_pytest_gpu.log
tail -n 8 gpurun_out/a_pytest_epi16.log
cat gpurun_out/a_summary.txt
python - <<'PY'
import json
for wl in ('r50','r50_roi_cells','r50_roi_cols8','r50_serp','r50_tail2','r50_tail1','r50_epi16','r50_graphs','r50_b1','r50_b1_graphs','r50_b2','r50_b2_graphs','ssd','r101'):
    try:
        d=json.load(open('gpurun_out/a_bench_%s.json'%wl)); print(wl, d['value'], d['ms_per_step'], d['e2e']['value'], d['category_ms_per_step'], d['roofline']['frac'])
    except Exception as e: print(wl, 'ERR', e)
PY
cat gpurun_out/parity_report_baseline.json 2>/dev/null | head -120
