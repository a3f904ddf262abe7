#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_golden_fixtures.py -m gpu -q -p no:cacheprovider > gpurun_out/k_all.log 2>&1
echo "kernels exit $?" > gpurun_out/summary.txt
timeout -s KILL 1800 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider > gpurun_out/e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err
echo "bench exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --workload ssd --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ssd.json 2> gpurun_out/bench_ssd.err
echo "ssd exit $?" >> gpurun_out/summary.txt
tail -n 12 gpurun_out/k_all.log; tail -n 12 gpurun_out/e2e.log; tail -n 2 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ['bench_r50','bench_ssd']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items()}, round(d['roofline']['frac'],4))
    except Exception as e: print(f,'fail',e)
PY
cat gpurun_out/summary.txt
