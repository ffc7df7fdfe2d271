// Store-path micro-benchmark for the GEMM epilogue (development tool, not part of the library):
// 148 CTAs x 8 warps write a [M][N] bf16 matrix tile by tile (128 x 160 tiles, round-robin over CTAs) with the
// access patterns the epilogue could use; reports time per pass and GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void stg256(void* p, uint32_t v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg128(void* p, uint32_t v) {
  asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

constexpr int TM = 128;

// pattern 0: direct, lane = row, 2 x 32 B per 32-column sub-block
// pattern 1: 8 rows x 64 B per instruction (16 B per lane), 32-column sub-blocks
// pattern 2: 4 rows x 128 B per instruction (16 B per lane), 64-column sub-blocks
// pattern 3: 2 rows x 256 B per instruction (16 B per lane), 128-column sub-blocks
// pattern 4: direct, lane = row, 4 x 32 B per 64-column sub-block (same as 0, different order)
template <int PAT>
__global__ void __launch_bounds__(256, 1) store_kernel(uint8_t* out, int M, int N, int TN, int passes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, grp = warp >> 2;
  const int m_tiles = M / TM, n_tiles = N / TN, tiles = m_tiles * n_tiles;
  const size_t ld = static_cast<size_t>(N) * 2;
  for (int pass = 0; pass < passes; ++pass) {
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int mt = t / n_tiles, nt = t - mt * n_tiles;
      uint8_t* tile = out + (static_cast<size_t>(mt) * TM + q * 32) * ld + static_cast<size_t>(nt) * TN * 2;
      const uint32_t v = t + pass;
      if (PAT == 0 || PAT == 4) {
        for (int c = grp * 32; c < TN; c += 64) {
          uint8_t* p = tile + lane * ld + c * 2;
          stg256(p, v);
          stg256(p + 32, v);
        }
      } else if (PAT == 1) {
        for (int c = grp * 32; c < TN; c += 64) {
#pragma unroll
          for (int i = 0; i < 4; ++i) stg128(tile + (i * 8 + (lane >> 2)) * ld + c * 2 + (lane & 3) * 16, v);
        }
      } else if (PAT == 2) {
        for (int c = grp * 64; c < TN; c += 128) {
          const int w = (TN - c) >= 64 ? 64 : (TN - c);           // columns of this sub-block
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if ((lane & 7) * 8 < w) stg128(tile + (i * 4 + (lane >> 3)) * ld + c * 2 + (lane & 7) * 16, v);
        }
      } else if (PAT == 3) {
        for (int c = grp * 128; c < TN; c += 256) {
          const int w = (TN - c) >= 128 ? 128 : (TN - c);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if ((lane & 15) * 8 < w) stg128(tile + (i * 2 + (lane >> 4)) * ld + c * 2 + (lane & 15) * 16, v);
        }
      }
    }
  }
}

// TMA stores: every warp stores its 32 rows x BW columns boxes from a (never rewritten) shared-memory buffer
template <int BW>
__global__ void __launch_bounds__(256, 1) tma_store_kernel(const __grid_constant__ CUtensorMap map, int M, int N, int TN, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, grp = warp >> 2;
  for (int i = threadIdx.x; i < 8 * 32 * BW * 2 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int m_tiles = M / TM, n_tiles = N / TN, tiles = m_tiles * n_tiles;
  uint8_t* buf = smem + warp * 32 * BW * 2;
  for (int pass = 0; pass < passes; ++pass) {
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int mt = t / n_tiles, nt = t - mt * n_tiles;
      for (int c = grp * BW; c < TN; c += 2 * BW) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map),
                       "r"(smem_u32(buf)), "r"(nt * TN + c), "r"(mt * TM + q * 32)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int M = 32768;
  const int passes = 20;
  uint8_t* out;
  CK(cudaMalloc(&out, static_cast<size_t>(M) * 2560 * 2));
  uint8_t* flush;
  CK(cudaMalloc(&flush, 256 << 20));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  const int cfgs[][2] = {{320, 160}, {960, 160}, {1280, 128}, {2560, 256}};
  for (auto& cfg : cfgs) {
    const int N = cfg[0], TN = cfg[1];
    const double bytes = static_cast<double>(M) * N * 2;
    auto run = [&](const char* name, auto launch) {
      launch(1);
      CK(cudaDeviceSynchronize());
      CK(cudaMemsetAsync(flush, 0, 256 << 20));
      CK(cudaEventRecord(e0));
      launch(passes);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double us = ms * 1e3 / passes;
      const double tiles_per_cta = static_cast<double>(M / TM) * (N / TN) / 148.0;
      printf("N=%4d TN=%3d %-44s %8.1f us/pass %7.0f GB/s  %6.0f clk/tile/SM @1.9GHz\n", N, TN, name, us, bytes / us * 1e-3,
             us * 1900.0 / tiles_per_cta);
    };
    run("direct 32 rows x 32 B (STG.256)", [&](int p) { store_kernel<0><<<148, 256>>>(out, M, N, TN, p); });
    run("8 rows x 64 B per instr (STG.128)", [&](int p) { store_kernel<1><<<148, 256>>>(out, M, N, TN, p); });
    run("4 rows x 128 B per instr (STG.128)", [&](int p) { store_kernel<2><<<148, 256>>>(out, M, N, TN, p); });
    run("2 rows x 256 B per instr (STG.128)", [&](int p) { store_kernel<3><<<148, 256>>>(out, M, N, TN, p); });
    auto tma = [&](int bw, auto kern, const char* name) {
      CUtensorMap map;
      const cuuint64_t dims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(M)};
      const cuuint64_t str[1] = {static_cast<cuuint64_t>(N) * 2};
      const cuuint32_t box[2] = {static_cast<cuuint32_t>(bw), 32u};
      const cuuint32_t es[2] = {1u, 1u};
      CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); return; }
      const int smem = 8 * 32 * bw * 2;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      run(name, [&](int p) { kern<<<148, 256, smem>>>(map, M, N, TN, p); });
    };
    tma(32, tma_store_kernel<32>, "TMA store 32 rows x 64 B per warp");
    tma(64, tma_store_kernel<64>, "TMA store 32 rows x 128 B per warp");
    if (TN % 128 == 0) tma(128, tma_store_kernel<128>, "TMA store 32 rows x 256 B per warp");
  }
  return 0;
}
