#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -k "ssd or predictor" -p no:cacheprovider > gpurun_out/e2e_ssd.log 2>&1
echo "e2e ssd exit $?" > gpurun_out/summary.txt
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err
echo "bench exit $?" >> gpurun_out/summary.txt
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r50.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
tail -n 8 gpurun_out/e2e_ssd.log; cat gpurun_out/bench_r50.json; tail -n 5 gpurun_out/bench_r50.err; cat gpurun_out/summary.txt
