#!/bin/bash
# Round 2, call C: defaults after call B (graphs on, chunk schedule, epi16 on one-stage tiles, ROI row-walk) + two-phase NMS validation.
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 --timeout-method=thread > gpurun_out/c_pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/c_summary.txt
cp gpurun_out/parity_report_baseline.json gpurun_out/c_parity_report_baseline_default.json 2>/dev/null
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/c_bench_r50.json 2> gpurun_out/c_bench_r50.err
echo "bench r50 exit $?" >> gpurun_out/c_summary.txt
for m in 4 6; do LUMI_ROI_MINB=$m timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_r50_minb$m.json 2>/dev/null; done
LUMI_GRAPHS=0 timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_r50_nographs.json 2>/dev/null
# two-phase NMS: kernel-level bit-exactness tests, BASELINE-config parity, speed
LUMI_NMS_LAZY=1 timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py tests/test_golden_fixtures.py -m gpu -q -p no:cacheprovider -k "nms or rpn or proposal or golden" --timeout 200 --timeout-method=thread > gpurun_out/c_pytest_lazy.log 2>&1
echo "pytest lazy kernels exit $?" >> gpurun_out/c_summary.txt
LUMI_NMS_LAZY=1 LUMI_PARITY_TAG=_lazy timeout -s KILL 600 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider -k "config2 or full_size or stages_and_detections" --timeout 500 --timeout-method=thread > gpurun_out/c_pytest_lazy_e2e.log 2>&1
echo "pytest lazy e2e exit $?" >> gpurun_out/c_summary.txt
LUMI_NMS_LAZY=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_r50_lazy.json 2>/dev/null
for b in 1 2; do
  timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --per-gpu-batch $b --no-cpu-baseline > gpurun_out/c_bench_r50_b${b}.json 2>/dev/null
  LUMI_NMS_LAZY=1 timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --per-gpu-batch $b --no-cpu-baseline > gpurun_out/c_bench_r50_b${b}_lazy.json 2>/dev/null
done
timeout -s KILL 300 python bench.py --workload ssd --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_ssd.json 2>/dev/null
timeout -s KILL 300 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_r101.json 2>/dev/null
tail -n 12 gpurun_out/c_pytest_gpu.log; tail -n 5 gpurun_out/c_pytest_lazy.log; tail -n 5 gpurun_out/c_pytest_lazy_e2e.log
cat gpurun_out/c_summary.txt
python - <<'PY'
import json
for wl in ('r50','r50_minb4','r50_minb6','r50_nographs','r50_lazy','r50_b1','r50_b1_lazy','r50_b2','r50_b2_lazy','ssd','r101'):
    try:
        d=json.load(open('gpurun_out/c_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4), d['gpu_launches'])
    except Exception as e: print(wl, 'ERR', e)
PY
