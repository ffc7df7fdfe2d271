// Host runtime of the C-ABI: init, error reporting, tensor-map encoding.
#include "ur_host.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace ur {

static thread_local char g_err[512] = "";
static int g_num_sms = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int set_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return UR_ERR_CUDA;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn g_encode = nullptr;

static int resolve_encode() {
  if (g_encode) return UR_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return set_error(UR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  g_encode = reinterpret_cast<encode_tiled_fn>(fn);
  return UR_OK;
}

int encode_tensor_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
  int rc = resolve_encode();
  if (rc) return rc;
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides[i];
    if (i < rank - 1) gs[i] = strides_bytes[i];
  }
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), base, gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(UR_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]",
                     static_cast<int>(r), rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                     (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0), bx[0],
                     rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
  }
  return UR_OK;
}

int num_sms() { return g_num_sms ? g_num_sms : 148; }

int pdl_enabled() {
  static const int on = []() {
    const char* e = getenv("UR_PDL");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on;
}

}  // namespace ur

using namespace ur;

extern "C" int ur_init(int device) {
  // no cudaSetDevice here: the caller (torch) owns the current device; callers make the tensor's device current
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDeviceProperties");
  if (prop.major != 10) return set_error(UR_ERR_CUDA, "unirestore_b200 requires an sm_100 GPU, found sm_%d%d", prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
  return resolve_encode();
}

extern "C" const char* ur_last_error(void) { return g_err; }
extern "C" int ur_version(void) { return 100; }
