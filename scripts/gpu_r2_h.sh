#!/bin/bash
# Round 2, call H: elect.sync issue (twelve back-to-back UTCHMMA per stage), tcgen05.mma issue-rate probe.
mkdir -p gpurun_out
: > gpurun_out/h_summary.txt
timeout -s KILL 300 python scripts/mma_probe.py > gpurun_out/h_mma_probe.txt 2>&1
echo "mma probe exit $?" >> gpurun_out/h_summary.txt
cat gpurun_out/h_mma_probe.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "conv" --timeout 120 --timeout-method=thread -x > gpurun_out/h_pytest_conv.log 2>&1
echo "pytest conv exit $?" >> gpurun_out/h_summary.txt
tail -n 4 gpurun_out/h_pytest_conv.log
if grep -q " passed" gpurun_out/h_pytest_conv.log && ! grep -q "failed" gpurun_out/h_pytest_conv.log; then
  timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/h_bench_r50_base.json 2> gpurun_out/h_bench_r50_base.err
  echo "bench base exit $?" >> gpurun_out/h_summary.txt
  LUMI_CONV_2CTA=9 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/h_bench_r50_cta2_9.json 2>/dev/null
  LUMI_CONV_2CTA=0 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/h_bench_r50_cta2_0.json 2>/dev/null
  LUMI_CONV_HALO=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/h_bench_r50_halo1.json 2>/dev/null
  timeout -s KILL 300 python bench.py --workload ssd --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/h_bench_ssd_base.json 2>/dev/null
  timeout -s KILL 300 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_r101_base.json 2>/dev/null
  timeout -s KILL 300 python bench.py --per-gpu-batch 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_r50_b1.json 2>/dev/null
fi
cat gpurun_out/h_summary.txt
python - <<'PY'
import json
for wl in ('r50_base','r50_cta2_9','r50_cta2_0','r50_halo1','ssd_base','r101_base','r50_b1'):
    try:
        d=json.load(open('gpurun_out/h_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4))
    except Exception as e: print(wl, 'ERR', e)
PY
