#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 --timeout-method=thread > gpurun_out/d_pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/d_summary.txt
cp gpurun_out/parity_report_baseline.json gpurun_out/d_parity_report_baseline.json 2>/dev/null; cp gpurun_out/parity_report.json gpurun_out/d_parity_report.json 2>/dev/null
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/d_bench_r50.json 2> gpurun_out/d_bench_r50.err
echo "bench r50 exit $?" >> gpurun_out/d_summary.txt
timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --per-gpu-batch 1 --no-cpu-baseline > gpurun_out/d_bench_r50_b1.json 2>/dev/null
timeout -s KILL 300 python bench.py --workload ssd --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_ssd.json 2>/dev/null
timeout -s KILL 300 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_r101.json 2>/dev/null
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/d_summary.txt
tail -n 8 gpurun_out/d_pytest_gpu.log; tail -2 gpurun_out/d_smoke.log; cat gpurun_out/d_summary.txt
python - <<'PY'
import json
for wl in ('r50','r50_b1','ssd','r101'):
    try:
        d=json.load(open('gpurun_out/d_bench_%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4), d['gpu_launches'])
    except Exception as e: print(wl, 'ERR', e)
PY
