#!/bin/bash
# quick validation after a kernel change: kernel tests, e2e tests, smoke, bench (no CPU baseline)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_golden_fixtures.py -m gpu -q -x -p no:cacheprovider > gpurun_out/k_all.log 2>&1
echo "kernels exit $?" > gpurun_out/summary.txt
timeout -s KILL 1500 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider > gpurun_out/e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err
echo "bench exit $?" >> gpurun_out/summary.txt
tail -n 6 gpurun_out/k_all.log; tail -n 8 gpurun_out/e2e.log; tail -n 3 gpurun_out/smoke.log
python -c "
import json; d=json.load(open('gpurun_out/bench_r50.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['category_ms_per_step'], d['roofline']['frac'])"
tail -n 3 gpurun_out/bench_r50.err; cat gpurun_out/summary.txt
