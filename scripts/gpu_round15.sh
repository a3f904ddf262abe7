#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_golden_fixtures.py -m gpu -q -x -p no:cacheprovider > gpurun_out/k_all.log 2>&1
echo "kernels exit $?" > gpurun_out/summary.txt
timeout -s KILL 1500 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider > gpurun_out/e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt
for rb in 4 1; do
LUMI_ROI_RB=$rb timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_rb$rb.json 2> gpurun_out/bench_r50_rb$rb.err
echo "bench rb$rb exit $?" >> gpurun_out/summary.txt
done
for b in 1 2; do
timeout -s KILL 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --per-gpu-batch $b > gpurun_out/bench_r50_b$b.json 2> gpurun_out/bench_r50_b$b.err
done
tail -n 4 gpurun_out/k_all.log; tail -n 6 gpurun_out/e2e.log
python - <<'PY'
import json
for f in ['bench_r50_rb4','bench_r50_rb1','bench_r50_b1','bench_r50_b2']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items()}, round(d['roofline']['frac'],4))
    except Exception as e: print(f,'fail',e)
PY
cat gpurun_out/summary.txt
