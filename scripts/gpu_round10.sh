#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider > gpurun_out/k_all.log 2>&1
echo "kernels exit $?" > gpurun_out/summary.txt
timeout -s KILL 1500 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider > gpurun_out/e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt
LUMI_ROI_CPL=8 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_cpl8.json 2> gpurun_out/bench_r50.err
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50.json 2>> gpurun_out/bench_r50.err
echo "bench exit $?" >> gpurun_out/summary.txt
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r50.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
tail -n 5 gpurun_out/k_all.log; tail -n 8 gpurun_out/e2e.log; python -c "
import json
for f in ['bench_r50_cpl8.json','bench_r50.json']:
    d=json.load(open('gpurun_out/'+f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['category_ms_per_step'], d['roofline']['frac'])"
tail -n 3 gpurun_out/bench_r50.err; cat gpurun_out/summary.txt
