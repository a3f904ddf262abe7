#!/bin/bash
# Round 2, call K: where does a conv stage's time go?  Timing-only experiments (LUMI_CONV_DBG; results are wrong by design):
# 1 no operand loads after the first ring fill, 2 no tcgen05.ld in the D1 drains, 4 no cross-term MMAs, 8 no output stores.
mkdir -p gpurun_out
for d in 0 1 2 4 8 3 5 7 15; do
  LUMI_ALLOW_WRONG_RESULTS=1 LUMI_CONV_DBG=$d LUMI_GRAPHS=0 timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --layers --no-cpu-baseline > gpurun_out/k_bench_r50_dbg$d.json 2>/dev/null
done
python - <<'PY'
import json
names=['block1/unit_2/bottleneck_v1/conv1','block1/unit_2/bottleneck_v1/conv2','block1/unit_2/bottleneck_v1/conv3','block2/unit_2/bottleneck_v1/conv1','block2/unit_2/bottleneck_v1/conv2','block2/unit_2/bottleneck_v1/conv3','block3/unit_2/bottleneck_v1/conv1','block3/unit_2/bottleneck_v1/conv2','block3/unit_2/bottleneck_v1/conv3','rpn/conv','conv1#s2d']
print('%-8s %8s '%('dbg','conv ms')+' '.join('%9s'%n.replace('/bottleneck_v1','').replace('unit_2/','')[-9:] for n in names))
for d in (0,1,2,4,8,3,5,7,15):
    try:
        j=json.load(open('gpurun_out/k_bench_r50_dbg%d.json'%d))
        L={l['layer']:l['us'] for l in j['conv_layers']}
        row=[]
        for n in names:
            k=[x for x in L if x.endswith(n)]
            row.append(L[k[0]] if k else float('nan'))
        print('%-8d %8.3f '%(d, j['category_ms_per_step']['conv_tc'])+' '.join('%9.1f'%v for v in row))
    except Exception as e: print(d,'ERR',e)
PY
