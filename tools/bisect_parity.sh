#!/usr/bin/env bash
# which opt-in variant breaks the full-size parity?  (one child per configuration; logs under gpurun_out/)
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  echo "=== $name: $*"
  env "$@" timeout 600 python -m pytest tests/test_parity_gpu.py -q -s -x -k "unet_v3d512" > gpurun_out/bisect_$name.log 2>&1
  echo "exit $?" >> gpurun_out/bisect_$name.log
  grep -E "rel-l2|passed|failed|Error" gpurun_out/bisect_$name.log | cut -c1-600
}
run alloff V3D_GN_FUSED=0 V3D_GEMM_2CTA=0 V3D_GEMM_RTMA=0
run gnfused V3D_GN_FUSED=1 V3D_GEMM_2CTA=0 V3D_GEMM_RTMA=0
run pair V3D_GN_FUSED=0 V3D_GEMM_2CTA=auto V3D_GEMM_RTMA=0
run rtma V3D_GN_FUSED=0 V3D_GEMM_2CTA=0 V3D_GEMM_RTMA=auto
run nograph V3D_GN_FUSED=0 V3D_GEMM_2CTA=0 V3D_GEMM_RTMA=0 V3D_CUDA_GRAPH=0
