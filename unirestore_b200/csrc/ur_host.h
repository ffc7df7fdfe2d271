// Host-side helpers shared by the C-ABI translation units: thread-local error text,
// CUDA error mapping, cuTensorMapEncodeTiled resolved at run time (no link-time libcuda
// dependency, so the library loads on a CPU-only box for the symbol checks).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/unirestore_b200.h"

namespace ur {
int set_error(int code, const char* fmt, ...);
int set_cuda_error(cudaError_t e, const char* what);
// bf16 tiled tensor map with 128-byte swizzle and zero OOB fill.
int encode_tensor_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes = 128);
int num_sms();
// 1: kernels are launched with programmatic stream serialization (PDL); UR_PDL=0 in the environment turns it off
int pdl_enabled();

// <<<...>>> replacement: cudaLaunchKernelEx with the PDL attribute (see ur_common.cuh)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace ur
