#!/bin/bash
# Round 2, call O: which of the last changes breaks the debug-vs-production agreement of the config-2 test?
mkdir -p gpurun_out
T="tests/test_gpu_baseline_configs.py::test_config2_frcnn_r50_batch8_600x1024"
run() { echo "=== $1"; env $1 LUMI_PARITY_TAG=_o timeout -s KILL 300 python -m pytest "$T" -m gpu -q -p no:cacheprovider --timeout 250 --timeout-method=thread 2>&1 | grep -E "passed|failed|assert [0-9]" | head -4; }
run "LUMI_X=0" > gpurun_out/o_bisect.txt 2>&1
run "LUMI_FMAP_F32_FUSED=0" >> gpurun_out/o_bisect.txt 2>&1
run "LUMI_CONV_2CTA=64" >> gpurun_out/o_bisect.txt 2>&1
run "LUMI_CONV_2CTA=0" >> gpurun_out/o_bisect.txt 2>&1
run "LUMI_CONV_2CTA=64 LUMI_FMAP_F32_FUSED=0" >> gpurun_out/o_bisect.txt 2>&1
run "LUMI_CONV_STREAMK=0" >> gpurun_out/o_bisect.txt 2>&1
run "LUMI_GRAPHS=0" >> gpurun_out/o_bisect.txt 2>&1
cat gpurun_out/o_bisect.txt
