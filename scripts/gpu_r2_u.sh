#!/bin/bash
# Round 2, call U: which build introduced the run-to-run differences of the two-stream pipeline?
mkdir -p gpurun_out
: > gpurun_out/u_bisect.txt
for v in res_single res_lane0; do
  cp build_variants/lib_$v.so luminoth_b200/libluminoth_b200.so
  echo "=== $v" >> gpurun_out/u_bisect.txt
  timeout -s KILL 200 python scripts/determinism_quick.py 2>&1 | grep -E "distinct|Error|error" >> gpurun_out/u_bisect.txt
done
cp build_variants/lib_head.so luminoth_b200/libluminoth_b200.so
cat gpurun_out/u_bisect.txt
