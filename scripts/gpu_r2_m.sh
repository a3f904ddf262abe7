#!/bin/bash
# Round 2, call M: what does the per-stage synchronisation of the MMA issuer cost (probe), and the conv without the
# tcgen05.fence after the operand-ring wait.
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/mma_probe.py --quick > gpurun_out/m_mma_probe.txt 2>&1
tail -n 12 gpurun_out/m_mma_probe.txt | cut -c1-60,118-
LUMI_CONV_DBG=16 timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "conv" --timeout 120 --timeout-method=thread -x > gpurun_out/m_pytest_conv_dbg16.log 2>&1
tail -n 2 gpurun_out/m_pytest_conv_dbg16.log
for d in 0 16; do
  LUMI_CONV_DBG=$d timeout -s KILL 200 python bench.py --steps 20 --warmup 4 --layers --no-cpu-baseline > gpurun_out/m_bench_r50_dbg$d.json 2>/dev/null
done
python - <<'PY'
import json
for d in (0,16):
    j=json.load(open('gpurun_out/m_bench_r50_dbg%d.json'%d))
    L={l['layer']:l['us'] for l in j['conv_layers']}
    print(d, round(j['value'],1), round(j['category_ms_per_step']['conv_tc'],3), [(k[-12:], round(v,1)) for k,v in L.items() if 'unit_2/bottleneck_v1/conv' in k or 'rpn/conv' in k])
PY
