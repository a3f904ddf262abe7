#!/bin/bash
# Round 2: 1 -> 8 GPU weak-scaling curve on ONE 8-GPU box (per-GPU batch 8), per-rank category times, plus the
# multi-device engine tests (two engines / two devices in one process).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/s_smi.txt 2>&1
: > gpurun_out/s_summary.txt
if [ -z "$SCALE_SKIP_TESTS" ]; then
timeout -s KILL 200 python -m pytest tests/test_gpu_engine_state.py -m gpu -q -p no:cacheprovider -k "two_engines or second_device" --timeout 150 --timeout-method=thread > gpurun_out/s_pytest_state.log 2>&1
echo "pytest engine state exit $?" >> gpurun_out/s_summary.txt
fi
timeout -s KILL 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench_n1.json 2> gpurun_out/s_bench_n1.err
echo "bench n1 exit $?" >> gpurun_out/s_summary.txt
for n in ${SCALE_NS:-2 4 8}; do
  NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/s_nccl_n${n}_%h_%p.log timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/s_bench_n$n.json 2> gpurun_out/s_bench_n$n.err
  rc=$?
  echo "bench n$n exit $rc" >> gpurun_out/s_summary.txt
  if [ $rc -ne 0 ]; then echo "N=$n failed: stopping the multi-rank runs" >> gpurun_out/s_summary.txt; tail -20 gpurun_out/s_bench_n$n.err; break; fi
done
# keep only a digest of the NCCL logs (rank count, algorithm / transport lines)
for n in 2 4 8; do cat gpurun_out/s_nccl_n${n}_*.log 2>/dev/null | grep -E "nranks|NVLS|Connected all|comm 0x.* rank .* nranks|Channel 00/" | head -40 > gpurun_out/s_nccl_n${n}_digest.txt; rm -f gpurun_out/s_nccl_n${n}_*.log; done
cat gpurun_out/s_summary.txt; tail -3 gpurun_out/s_pytest_state.log
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.load(open('gpurun_out/s_bench_n%d.json'%n))
        if n==1: base=d['value']
        print(n, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'eff', round(d['value']/(n*base),3) if base else None, 'bcast ms', round(d.get('weight_bcast_ms',0),1))
        for r in (d.get('per_rank') or []): print('    ', r)
    except Exception as e: print(n,'ERR',e)
PY
