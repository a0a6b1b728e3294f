#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; echo "=== $* -> $log"; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "exit $?" >> "gpurun_out/$log"; tail -n 3 "gpurun_out/$log" | cut -c1-300; }
run 1200 tests_subset.log python -m pytest tests/test_kernels_gpu.py tests/test_viewshard_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "not one_gpu_gloo"
run 900 bench.json python bench.py --steps 3 --warmup 3
run 600 ncu_set.log ncu --set full --clock-control none --import-source on -k "regex:gn_|layernorm|attn_|gemm_tc|softmax_rows" -c 44 -o gpurun_out/ncu_set -f python tools/microbench.py ncu_set
run 900 sweep.log python tools/sweep.py --frames 24 --edm-steps 25 --steps 1 --warmup 3 --extra --no-parity
mv gpurun_out/sweep.json gpurun_out/sweep_t24.json 2>/dev/null
run 900 sweep2.log python tools/sweep.py --frames 14 25 --edm-steps 10 50 --steps 1 --warmup 3 --extra --no-parity
