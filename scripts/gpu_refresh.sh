#!/bin/bash
# short end-of-round refresh after a kernel change: all GPU tests, smoke, bench (N=1, with the CPU arm), launch list
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/summary.txt
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 --layers > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err
echo "bench exit $?" >> gpurun_out/summary.txt
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r50.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log
python -c "
import json; d=json.load(open('gpurun_out/bench_r50.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['category_ms_per_step'], d['roofline']['frac'], d.get('cpu_baseline'))"
cat gpurun_out/summary.txt
