#!/bin/bash
# 2-GPU validation of the multi-rank bench path (NCCL weight broadcast, record all-gather, per-rank report, clean exit)
mkdir -p gpurun_out
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "bench n2 exit $?" > gpurun_out/n2_summary.txt
timeout -s KILL 200 python -m pytest tests/test_gpu_engine_state.py -m gpu -q -p no:cacheprovider -k "two_engines or second_device" --timeout 150 --timeout-method=thread > gpurun_out/n2_pytest_state.log 2>&1
echo "pytest engine state exit $?" >> gpurun_out/n2_summary.txt
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/n2_bench_n1.json 2>/dev/null
echo "bench n1 exit $?" >> gpurun_out/n2_summary.txt
cat gpurun_out/n2_summary.txt; tail -3 gpurun_out/n2_pytest_state.log; tail -5 gpurun_out/n2_bench.err
python - <<'PY'
import json
d1=json.load(open('gpurun_out/n2_bench_n1.json')); print('n1', round(d1['value'],1), round(d1['ms_per_step'],3), {k:round(v,3) for k,v in d1['category_ms_per_step'].items() if v>0})
d=json.load(open('gpurun_out/n2_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'bcast', d.get('weight_bcast_ms'))
for r in d.get('per_rank') or []: print('   ', r)
PY
