#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "ref n$N exit $?" > gpurun_out/summary.txt
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r50_n$N.json 2> gpurun_out/bench_r50_n$N.err
echo "n$N exit $?" >> gpurun_out/summary.txt
head -c 600 gpurun_out/bench_r50_n$N.json; echo; head -c 300 gpurun_out/bench_ref_n$N.json; echo
tail -n 5 gpurun_out/bench_r50_n$N.err; cat gpurun_out/summary.txt
