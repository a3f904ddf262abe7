#!/bin/bash
# Round 2, call F: first contact of the halo-patch conv kernels (shifted SWIZZLE_128B views into one patch per 64-channel
# slice).  Step 1 decides how the matrix descriptor wants the start address' swizzle phase (base offset 0 or (addr>>7)&7)
# on one small layer, short timeouts; the rest only runs with the setting that passes.
mkdir -p gpurun_out
: > gpurun_out/f_summary.txt
OK=""
for bo in 0 1; do
  LUMI_HALO_BASEOFF=$bo timeout -s KILL 180 python -m pytest "tests/test_gpu_kernels.py::test_conv2d_matches_oracle[halo_38x64_128-tc_split_halo]" \
    "tests/test_gpu_kernels.py::test_conv2d_matches_oracle[relu6-tc_split_halo]" -m gpu -q -p no:cacheprovider --timeout 60 --timeout-method=thread > gpurun_out/f_pytest_halo_probe_bo$bo.log 2>&1
  echo "probe baseoff=$bo exit $?" >> gpurun_out/f_summary.txt
  tail -n 12 gpurun_out/f_pytest_halo_probe_bo$bo.log
  if grep -q " passed" gpurun_out/f_pytest_halo_probe_bo$bo.log && ! grep -q "failed\|error" gpurun_out/f_pytest_halo_probe_bo$bo.log; then OK=$bo; break; fi
done
echo "working baseoff: '$OK'" >> gpurun_out/f_summary.txt
if [ -z "$OK" ]; then
  for bo in 0 1; do LUMI_HALO_BASEOFF=$bo timeout -s KILL 120 python scripts/halo_probe.py > gpurun_out/f_halo_probe_bo$bo.txt 2>&1; done
  head -n 60 gpurun_out/f_halo_probe_bo0.txt
fi
if [ -n "$OK" ]; then
  export LUMI_HALO_BASEOFF=$OK
  timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "halo" --timeout 60 --timeout-method=thread > gpurun_out/f_pytest_halo.log 2>&1
  echo "pytest halo exit $?" >> gpurun_out/f_summary.txt
  tail -n 15 gpurun_out/f_pytest_halo.log
  timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/f_bench_r50_base.json 2>/dev/null
  for m in 1 2; do
    LUMI_CONV_HALO=$m timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/f_bench_r50_halo$m.json 2> gpurun_out/f_bench_r50_halo$m.err
    echo "bench halo $m exit $?" >> gpurun_out/f_summary.txt
  done
  LUMI_CONV_HALO=2 LUMI_PARITY_TAG=_halo timeout -s KILL 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -p no:cacheprovider -k "config2" --timeout 500 --timeout-method=thread > gpurun_out/f_pytest_halo_parity.log 2>&1
  echo "pytest halo parity exit $?" >> gpurun_out/f_summary.txt
  for m in 0 2; do
    LUMI_CONV_HALO=$m timeout -s KILL 300 python bench.py --workload ssd --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/f_bench_ssd_halo$m.json 2>/dev/null
  done
  LUMI_CONV_HALO=2 timeout -s KILL 300 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_r101_halo2.json 2>/dev/null
fi
cat gpurun_out/f_summary.txt
python - <<'PY'
import json
for wl in ('r50_base','r50_halo1','r50_halo2','ssd_halo0','ssd_halo2','r101_halo2'):
    try:
        d=json.load(open('gpurun_out/f_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4))
    except Exception as e: print(wl, 'ERR', e)
PY
