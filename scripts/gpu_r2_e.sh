#!/bin/bash
# Round 2, call E: first contact of the CTA-pair (cta_group::2) conv kernel -- op-level parity tests only, short timeouts.
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "cta2" --timeout 60 --timeout-method=thread -x > gpurun_out/e_pytest_cta2.log 2>&1
echo "pytest cta2 exit $?" > gpurun_out/e_summary.txt
tail -n 30 gpurun_out/e_pytest_cta2.log
if grep -q " passed" gpurun_out/e_pytest_cta2.log && ! grep -q "failed" gpurun_out/e_pytest_cta2.log; then
  for m in 9 36; do
    LUMI_CONV_2CTA=$m timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/e_bench_r50_cta2_$m.json 2> gpurun_out/e_bench_r50_cta2_$m.err
    echo "bench cta2 $m exit $?" >> gpurun_out/e_summary.txt
  done
  timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/e_bench_r50_base.json 2>/dev/null
  LUMI_CONV_2CTA=9 LUMI_PARITY_TAG=_cta2 timeout -s KILL 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -p no:cacheprovider -k "config2" --timeout 500 --timeout-method=thread > gpurun_out/e_pytest_cta2_parity.log 2>&1
  echo "pytest cta2 parity exit $?" >> gpurun_out/e_summary.txt
  LUMI_CONV_2CTA=9 timeout -s KILL 300 python bench.py --workload ssd --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_ssd_cta2.json 2>/dev/null
  LUMI_CONV_2CTA=9 timeout -s KILL 300 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_r101_cta2.json 2>/dev/null
fi
cat gpurun_out/e_summary.txt
python - <<'PY'
import json
for wl in ('r50_base','r50_cta2_9','r50_cta2_36','ssd_cta2','r101_cta2'):
    try:
        d=json.load(open('gpurun_out/e_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4))
    except Exception as e: print(wl, 'ERR', e)
PY
