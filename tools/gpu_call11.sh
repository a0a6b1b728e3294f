#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; echo "=== $* -> $log"; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "exit $?" >> "gpurun_out/$log"; tail -n 3 "gpurun_out/$log" | cut -c1-300; }
run 500 tests_sub.log python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "subprocess or concat_timestep"
run 700 tests_vs.log python -m pytest tests/test_viewshard_gpu.py -m gpu -x -q -s -k "kv_scatter or peer_transport_one_gpu"
run 600 tests_parity.log python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s
run 300 ncu_set.log ncu --set full --clock-control none -k "regex:gn_|layernorm|attn_|gemm_tc|softmax_rows" -s 13 -c 13 -o gpurun_out/ncu_set_final -f python tools/microbench.py ncu_set
