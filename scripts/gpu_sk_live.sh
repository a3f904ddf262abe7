#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
LUMI_CONV_STREAMK=$m timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layers > gpurun_out/bench_sk$m.json 2> gpurun_out/bench_sk$m.err
echo "sk$m exit $?" >> gpurun_out/summary.txt
done
python - <<'PY'
import json
a=json.load(open('gpurun_out/bench_sk0.json')); b=json.load(open('gpurun_out/bench_sk1.json'))
print('value', a['value'], b['value'], 'conv_tc', a['category_ms_per_step']['conv_tc'], b['category_ms_per_step']['conv_tc'])
for x,y in zip(a['conv_layers'], b['conv_layers']):
    print('%-75s %7.1f %7.1f %+5.0f%%  %6.1f TF' % (x['layer'][-75:], x['us'], y['us'], 100*(y['us']/x['us']-1), y['tflops'] or 0))
PY
cat gpurun_out/summary.txt
