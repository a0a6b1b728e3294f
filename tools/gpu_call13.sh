#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}_fused.json 2>&1
echo "exit $?" >> gpurun_out/bench_n${N}_fused.json
tail -n 2 gpurun_out/bench_n${N}_fused.json | cut -c1-300
