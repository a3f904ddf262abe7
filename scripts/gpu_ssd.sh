#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -k "ssd or predictor" -p no:cacheprovider > gpurun_out/e2e_ssd.log 2>&1
echo "e2e ssd exit $?" > gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --workload ssd --steps 20 --warmup 3 --no-cpu-baseline --layers > gpurun_out/bench_ssd.json 2> gpurun_out/bench_ssd.err
echo "ssd exit $?" >> gpurun_out/summary.txt
tail -n 5 gpurun_out/e2e_ssd.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ssd.json')); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['category_ms_per_step'].items()}, round(d['roofline']['frac'],4))
for x in d['conv_layers']: print('%-70s %8.1f us %7.1f TF' % (x['layer'][-70:], x['us'], x['tflops'] or 0))
PY
cat gpurun_out/summary.txt
