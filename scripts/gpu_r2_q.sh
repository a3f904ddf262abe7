#!/bin/bash
# Round 2, call Q: op-level check of the fused fp32 copy of a split-plane conv output.
mkdir -p gpurun_out
for m in 0 1 2; do
  echo "=== LUMI_OP_ALSO=$m"
  LUMI_OP_ALSO=$m timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "conv2d_matches and tc_split and (endpoint or b1_conv3 or b2_conv3 or sk_many or 1x1_64_256)" --timeout 200 --timeout-method=thread 2>&1 | grep -E "passed|failed|Error|assert|mismatch|max" | head -12
done
