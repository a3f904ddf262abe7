#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv" -p no:cacheprovider > gpurun_out/k_conv.log 2>&1
echo "kernels exit $?" > gpurun_out/summary.txt
timeout -s KILL 1500 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider > gpurun_out/e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layers > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err
tail -n 4 gpurun_out/k_conv.log; tail -n 6 gpurun_out/e2e.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r50.json')); print(round(d['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['category_ms_per_step'].items()}, round(d['roofline']['frac'],4))
for x in d['conv_layers']: print('%-62s %7.1f us %6.1f TF' % (x['layer'].replace('truncated_base_network/resnet_v1_50/','').replace('/bottleneck_v1',''), x['us'], x['tflops']))
r=json.load(open('gpurun_out/parity_report.json'))
for k,v in r.items():
    if 'tc' in k: print(k, {a:('%.2e'%b) for a,b in v.items() if 'engine' in a or 'oracle' in a})
PY
cat gpurun_out/summary.txt
