#!/bin/bash
# Round 2, call G: warp-uniform TMA / MMA issue (operands in uniform registers instead of a 5-R2UR waterfall loop
# around every UTCHMMA): conv parity tests, then per-layer timings under the existing switches.
mkdir -p gpurun_out
: > gpurun_out/g_summary.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "conv" --timeout 120 --timeout-method=thread -x > gpurun_out/g_pytest_conv.log 2>&1
echo "pytest conv exit $?" >> gpurun_out/g_summary.txt
tail -n 6 gpurun_out/g_pytest_conv.log
if grep -q " passed" gpurun_out/g_pytest_conv.log && ! grep -q "failed" gpurun_out/g_pytest_conv.log; then
  timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/g_bench_r50_base.json 2> gpurun_out/g_bench_r50_base.err
  echo "bench base exit $?" >> gpurun_out/g_summary.txt
  LUMI_CONV_2CTA=9 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/g_bench_r50_cta2_9.json 2>/dev/null
  LUMI_CONV_HALO=2 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/g_bench_r50_halo2.json 2>/dev/null
  LUMI_CONV_HALO=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/g_bench_r50_halo1.json 2>/dev/null
  timeout -s KILL 300 python bench.py --workload ssd --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/g_bench_ssd_base.json 2>/dev/null
  timeout -s KILL 300 python bench.py --per-gpu-batch 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/g_bench_r50_b1.json 2>/dev/null
  timeout -s KILL 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_engine_state.py -m gpu -q -p no:cacheprovider --timeout 500 --timeout-method=thread > gpurun_out/g_pytest_e2e.log 2>&1
  echo "pytest e2e exit $?" >> gpurun_out/g_summary.txt
  tail -n 3 gpurun_out/g_pytest_e2e.log
fi
cat gpurun_out/g_summary.txt
python - <<'PY'
import json
for wl in ('r50_base','r50_cta2_9','r50_halo2','r50_halo1','ssd_base','r50_b1'):
    try:
        d=json.load(open('gpurun_out/g_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4))
    except Exception as e: print(wl, 'ERR', e)
PY
