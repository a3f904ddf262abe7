#!/bin/bash
mkdir -p gpurun_out
for r in 0 4 8 16; do
LUMI_CONV_SM_RESERVE=$r timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_res$r.json 2> gpurun_out/bench_r50.err
done
python -c "
import json
for r in [0,4,8,16]:
    d=json.load(open('gpurun_out/bench_r50_res%d.json'%r)); print(r, d['value'], d['ms_per_step'], d['e2e']['value'])"
