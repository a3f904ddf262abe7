#!/bin/bash
# Round 2, call J: halo-patch kernels again, now that the issue path is cheap (pairs: 16 KB of weights per stage and CTA).
mkdir -p gpurun_out
: > gpurun_out/j_summary.txt
for m in 2 1 0; do
  LUMI_CONV_HALO=$m timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/j_bench_r50_halo$m.json 2>/dev/null
  echo "bench halo $m exit $?" >> gpurun_out/j_summary.txt
done
LUMI_CONV_HALO=2 LUMI_CONV_HALO_PCT=110 timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/j_bench_r50_halo2_pct110.json 2>/dev/null
for m in 2 0; do
  LUMI_CONV_HALO=$m timeout -s KILL 300 python bench.py --workload ssd --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/j_bench_ssd_halo$m.json 2>/dev/null
done
cat gpurun_out/j_summary.txt
python - <<'PY'
import json
for wl in ('r50_halo0','r50_halo1','r50_halo2','r50_halo2_pct110','ssd_halo0','ssd_halo2'):
    try:
        d=json.load(open('gpurun_out/j_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4))
        print('   ', [(l['layer'][-12:], round(l['us'],1)) for l in d['conv_layers'] if 'unit_2/bottleneck_v1/conv2' in l['layer'] or 'rpn/conv' in l['layer'] or 'conv1_2' in l['layer'] or 'conv3_2' in l['layer'] or 'conv4_2' in l['layer']])
    except Exception as e: print(wl, 'ERR', e)
PY
