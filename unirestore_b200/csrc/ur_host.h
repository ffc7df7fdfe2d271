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
}  // namespace ur
