#!/bin/bash
# Round 2, call V: final validation after the residual-producer fix: determinism, full GPU suite, smoke, bench lines.
mkdir -p gpurun_out
O=gpurun_out/ev
REPS=6 timeout -s KILL 300 python scripts/determinism_diag.py > gpurun_out/v_determinism.txt 2>&1
cat gpurun_out/v_determinism.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 --timeout-method=thread > ${O}_pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > ${O}_summary.txt
cp gpurun_out/parity_report.json ${O}_parity_report.json 2>/dev/null; cp gpurun_out/parity_report_baseline.json ${O}_parity_report_baseline.json 2>/dev/null
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1
echo "smoke exit $?" >> ${O}_summary.txt
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 --layers > ${O}_bench_r50.json 2> ${O}_bench_r50.err
echo "bench r50 exit $? (normal interpreter exit)" >> ${O}_summary.txt
timeout -s KILL 600 python bench.py --impl reference --steps 1 --warmup 0 > ${O}_bench_ref.json 2> ${O}_bench_ref.err
echo "bench ref exit $?" >> ${O}_summary.txt
timeout -s KILL 400 python bench.py --workload ssd --steps 20 --warmup 3 --layers --no-cpu-baseline > ${O}_bench_ssd.json 2>/dev/null
timeout -s KILL 400 python bench.py --workload frcnn_r101 --steps 10 --warmup 3 --layers --no-cpu-baseline > ${O}_bench_r101.json 2>/dev/null
for b in 1 2; do timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --per-gpu-batch $b --no-cpu-baseline > ${O}_bench_r50_latency_b$b.json 2>/dev/null; done
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file ${O}_launches_r50.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${O}_ncu_bench.log 2>&1
echo "ncu launches exit $?" >> ${O}_summary.txt
LUMI_CONV_DBG=32 timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "conv" --timeout 120 --timeout-method=thread > gpurun_out/v_pytest_conv_dbg32.log 2>&1
echo "pytest conv dbg32 exit $?" >> ${O}_summary.txt
LUMI_CONV_DBG=32 timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/v_bench_r50_dbg32.json 2>/dev/null
tail -n 6 ${O}_pytest_gpu.log; tail -n 2 ${O}_smoke.log; tail -n 2 gpurun_out/v_pytest_conv_dbg32.log
python - <<'PY'
import json
for wl in ('ev_bench_r50','ev_bench_ssd','ev_bench_r101','ev_bench_r50_latency_b1','ev_bench_r50_latency_b2','v_bench_r50_dbg32'):
    try:
        d=json.load(open('gpurun_out/%s.json'%wl)); print(wl, round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items() if v>0}, round(d['roofline']['frac'],4), d['roofline'].get('traffic'))
    except Exception as e: print(wl,'ERR',e)
PY
cat ${O}_summary.txt
