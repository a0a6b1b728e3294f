#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_viewshard_gpu.py -m gpu -q -s -x -k "view_sharded_engine_matches_unsharded_nccl and not collectives" > gpurun_out/tests_nccl_fused.log 2>&1
echo "exit $?" >> gpurun_out/tests_nccl_fused.log
tail -n 4 gpurun_out/tests_nccl_fused.log | cut -c1-600
