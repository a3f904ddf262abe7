#!/bin/bash
# 2-GPU run: NCCL weight broadcast + detection all-gather (bench contract launch line)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r50_n2.json 2> gpurun_out/bench_r50_n2.err
echo "n2 exit $?" > gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --workload ssd --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ssd.json 2> gpurun_out/bench_ssd.err
echo "ssd exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --workload frcnn_r101 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r101.json 2> gpurun_out/bench_r101.err
echo "r101 exit $?" >> gpurun_out/summary.txt
python -c "
import json
for f in ['bench_r50_n2.json','bench_ssd.json','bench_r101.json']:
    try:
        d=json.load(open('gpurun_out/'+f)); print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['category_ms_per_step'], d['roofline']['frac'] if d['roofline'] else None, d.get('weight_bcast_ms'))
    except Exception as e: print(f, 'ERR', e)
"
tail -n 5 gpurun_out/*.err; cat gpurun_out/summary.txt
