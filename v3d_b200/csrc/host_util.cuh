// Host-side helpers shared by the C-ABI translation units: error slot, launch counter,
// lazily-resolved cuTensorMapEncodeTiled (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace v3d {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// bf16 tensor map, zero OOB fill. dims/strides innermost first; strides[i] is the byte stride of dim i+1.
// swizzle_bytes: 128 / 64 / 32, or 0 for a dense box. Returns 0 on success.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128);

#define V3D_CHECK_LAUNCH(name)                                            \
  do {                                                                    \
    cudaError_t e__ = cudaGetLastError();                                 \
    if (e__ != cudaSuccess) {                                             \
      v3d::set_error("%s launch failed: %s", name, cudaGetErrorString(e__)); \
      return V3D_ERR_CUDA;                                                \
    }                                                                     \
    v3d::count_launch();                                                  \
  } while (0)

}  // namespace v3d
