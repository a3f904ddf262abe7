#!/bin/bash
# Round 2: the N = 8 point of the weak-scaling curve alone (8-GPU box time is charged 8x).
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/s_nccl_n8_%h_%p.log timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/s_bench_n8.json 2> gpurun_out/s_bench_n8.err
echo "bench n8 exit $?"
cat gpurun_out/s_nccl_n8_*.log 2>/dev/null | grep -E "nranks|NVLS|Connected all|Channel 00/" | head -40 > gpurun_out/s_nccl_n8_digest.txt; rm -f gpurun_out/s_nccl_n8_*.log
python -c "
import json; d=json.load(open('gpurun_out/s_bench_n8.json')); print(8, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1)); [print('   ', r) for r in d.get('per_rank') or []]"
