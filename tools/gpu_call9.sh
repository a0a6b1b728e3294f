#!/usr/bin/env bash
# N GPUs (NGPU): the NCCL-group plan tests that need N ranks, then the default bench line (image-parallel + strong)
set -u
mkdir -p gpurun_out
N=${NGPU:-4}
run() { local t=$1 log=$2; shift 2; echo "=== $* -> $log"; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "exit $?" >> "gpurun_out/$log"; tail -n 3 "gpurun_out/$log" | cut -c1-300; }
if [ "$N" = "4" ]; then
  run 400 viewshard_n4_tests.log python -m pytest tests/test_viewshard_gpu.py -m gpu -q -s -x -k "cfg_views_engine_matches_unsharded_nccl or cfg_views_engine_peer_transport"
fi
run 900 bench_n$N.json python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3
