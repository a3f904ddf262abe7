#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r50_n2.json 2> gpurun_out/bench_r50_n2.err
echo "n2 exit $?" > gpurun_out/summary.txt
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "ref n2 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_r50_n2.json | cut -c1-600; echo; cat gpurun_out/bench_ref_n2.json | cut -c1-400; echo
grep -v "^frame" gpurun_out/bench_r50_n2.err | tail -n 8 | cut -c1-200; cat gpurun_out/summary.txt
