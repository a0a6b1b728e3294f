#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_fullsize_gpu.py tests/test_standins_gpu.py tests/test_zz_attention_rescale_gpu.py -m gpu -q > gpurun_out/tests_rest.log 2>&1
echo "exit $?" >> gpurun_out/tests_rest.log
tail -n 4 gpurun_out/tests_rest.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "exit $?" >> gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log | cut -c1-300
