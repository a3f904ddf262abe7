#!/bin/bash
# Round-end evidence run: full GPU test suite, smoke, one-step ncu --set full summary, bench (N=1), ncu launch list.
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/summary.txt
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
# one step under ncu --set full (single stream, whole batch: the configuration of bench.py's event profile)
timeout -s KILL 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/ncu_step python bench.py --ncu-range --ncu-unpiped --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?" >> gpurun_out/summary.txt
ncu -i /tmp/ncu_step.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_step_raw.csv.gz
python scripts/ncu_step_summary.py /tmp/ncu_step.ncu-rep frcnn_r50 gpurun_out/ncu_step > gpurun_out/ncu_step_summary.txt 2>&1
echo "summary exit $?" >> gpurun_out/summary.txt
cp gpurun_out/ncu_step_summary.json profiles/r1_ncu_step_summary.json     # bench.py reads roofline.traffic from here
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 --layers > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err
echo "bench exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref exit $?" >> gpurun_out/summary.txt
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r50.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
# a small report with source correlation for local inspection (roi + nms + 3 conv launches)
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:roi_pool|nms_mask|nms_scan|sort_desc" -c 5 -f -o gpurun_out/prof_post_small python bench.py --ncu-range --ncu-unpiped --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log
python -c "
import json; d=json.load(open('gpurun_out/bench_r50.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['category_ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d.get('cpu_baseline'))"
tail -n 3 gpurun_out/bench_r50.err; cat gpurun_out/ncu_step_summary.txt | tail -n 30; cat gpurun_out/summary.txt; du -sm gpurun_out
