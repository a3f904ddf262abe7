#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --workload ssd --steps 20 --warmup 3 --no-cpu-baseline --layers > gpurun_out/bench_ssd.json 2> gpurun_out/bench_ssd.err
timeout -s KILL 300 python bench.py --workload frcnn_r101 --steps 20 --warmup 3 --no-cpu-baseline --layers > gpurun_out/bench_r101.json 2> gpurun_out/bench_r101.err
for b in 1 2; do
timeout -s KILL 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --per-gpu-batch $b > gpurun_out/bench_r50_b$b.json 2> gpurun_out/bench_r50_b$b.err
done
python - <<'PY'
import json
for f in ['bench_ssd','bench_r101','bench_r50_b1','bench_r50_b2']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items()}, round(d['roofline']['frac'],4))
    except Exception as e: print(f,'fail',e)
PY
