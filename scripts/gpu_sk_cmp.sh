#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
LUMI_CONV_STREAMK=$m timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_sk$m.csv python bench.py --ncu-range --no-cpu-baseline > gpurun_out/ncu_sk$m.log 2>&1
echo "sk$m exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
