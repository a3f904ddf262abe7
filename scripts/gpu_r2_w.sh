#!/bin/bash
# Round 2, call W: confirmation of the final binary (experiment code removed): full GPU suite + headline bench.
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 --timeout-method=thread > gpurun_out/w_pytest_gpu.log 2>&1
echo "pytest gpu exit $?"
tail -n 3 gpurun_out/w_pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/w_bench_r50.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/w_bench_r50.json')); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, d['roofline']['traffic_note'][-60:])"
