#!/bin/bash
# Round 2, call I: does tcgen05.ld traffic (the D1 chunk drain) stall the MMA pipe?  Probe + the conv under 1 / 2 / 4-stage
# D1 chunks with the pair kernel from 9 K stages.
mkdir -p gpurun_out
: > gpurun_out/i_summary.txt
timeout -s KILL 300 python scripts/mma_probe.py > gpurun_out/i_mma_probe.txt 2>&1
echo "mma probe exit $?" >> gpurun_out/i_summary.txt
tail -n 22 gpurun_out/i_mma_probe.txt
for t in 2 4 1; do
  LUMI_CONV_CHUNK_TAIL=$t timeout -s KILL 300 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/i_bench_r50_tail$t.json 2>/dev/null
  echo "bench tail $t exit $?" >> gpurun_out/i_summary.txt
done
cat gpurun_out/i_summary.txt
python - <<'PY'
import json
for wl in ('r50_tail2','r50_tail4','r50_tail1'):
    try:
        d=json.load(open('gpurun_out/i_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4))
        print('   ', [(l['layer'][-12:], round(l['us'],1)) for l in d['conv_layers'] if 'unit_2/bottleneck_v1/conv2' in l['layer'] or 'rpn/conv' in l['layer']])
    except Exception as e: print(wl, 'ERR', e)
PY
