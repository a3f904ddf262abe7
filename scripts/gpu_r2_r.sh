#!/bin/bash
# Round 2, call R: why does the fused fp32 copy break the endpoint-sized conv?  (op level, LUMI_OP_ALSO=2 returns the split planes)
mkdir -p gpurun_out
T="tests/test_gpu_kernels.py::test_conv2d_matches_oracle[b3_conv3_endpoint-tc_split]"
run() { echo "=== $*"; env "$@" timeout -s KILL 200 python -m pytest "$T" -m gpu -q -p no:cacheprovider --timeout 150 --timeout-method=thread 2>&1 | grep -E "passed|failed|AssertionError:" | head -3; }
run LUMI_OP_ALSO=0
run LUMI_OP_ALSO=2
run LUMI_OP_ALSO=0 LUMI_CONV_DBG=32
run LUMI_OP_ALSO=2 LUMI_CONV_DBG=32
run LUMI_OP_ALSO=2 LUMI_CONV_DBG=64
run LUMI_OP_ALSO=2 LUMI_CONV_DBG=128
run LUMI_OP_ALSO=1 LUMI_CONV_DBG=128
run LUMI_OP_ALSO=2 LUMI_CONV_DBG=16
echo "=== memcheck (LUMI_OP_ALSO=2)"
LUMI_OP_ALSO=2 timeout -s KILL 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "$T" -m gpu -q -p no:cacheprovider --timeout 350 --timeout-method=thread 2>&1 | grep -E "passed|failed|Invalid|ERROR SUMMARY|at 0x|by thread|Address" | head -20
