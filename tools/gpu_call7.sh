#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() { local t=$1 log=$2; shift 2; echo "=== $* -> $log"; timeout "$t" "$@" > "gpurun_out/$log" 2>&1; echo "exit $?" >> "gpurun_out/$log"; tail -n 4 "gpurun_out/$log" | cut -c1-300; }
run 600 attn3_tests.log python -m pytest tests/test_kernels_gpu.py tests/test_zz_attention_rescale_gpu.py -q -x -m gpu -k "attention_spatial or groupnorm"
V3D_ATTN_TILES=2 run 600 attn2_tests.log python -m pytest tests/test_kernels_gpu.py tests/test_zz_attention_rescale_gpu.py -q -x -m gpu -k "attention_spatial"
run 300 micro_attn_3tile.log python tools/microbench.py attn
V3D_ATTN_TILES=2 run 300 micro_attn_2tile.log python tools/microbench.py attn
run 1500 tests_gpu.log python -m pytest tests -m gpu -x -q
run 900 bench.json python bench.py --steps 3 --warmup 3
