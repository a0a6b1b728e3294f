#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; echo "=== $* -> $log"; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "exit $?" >> "gpurun_out/$log"; tail -n 3 "gpurun_out/$log" | cut -c1-300; }
run 600 tests_clip.log python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "clip"
run 700 tests_cfgpeer.log python -m pytest tests/test_viewshard_gpu.py -m gpu -x -q -s -k "cfg_split_engine_peer_transport_one_gpu"
