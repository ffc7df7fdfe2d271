// ur_group_norm: one-launch GroupNorm (+SiLU) for tensors that stay L2-resident (every per-DDIM-step GroupNorm).
//
// One thread-block CLUSTER per image (8 or 16 CTAs, distributed shared memory):
//   phase 1  each CTA reads its slab of pixels (128-bit loads, all channels) and accumulates per-channel
//            (sum, sumsq) in shared memory
//   phase 2  CTA r reduces channel slice r over all peers through DSMEM, turns its G/CL groups into (mean, rstd)
//   phase 3  every CTA gathers the G (mean, rstd) pairs from their owners through DSMEM
//   phase 4  each CTA re-reads its slab (L2 hit), applies scale/shift (+SiLU) and writes the concatenated output
// This replaces ur_chan_stats (x2 for a concatenated skip) + ur_norm_apply = 2-3 launches and the fp64 global
// atomics on the per-step path.  Same arithmetic: fp32 per-thread partials, fp64 cross-CTA / group reduction,
// var = E[x^2] - mean^2 clamped at 0 (reference: nn.GroupNorm inside diffusers ResnetBlock2D / Transformer2DModel /
// Attention.group_norm, base_model.py:54,138, controller.py:161-170).
#include "ur_common.cuh"
#include "ur_host.h"

#include <stdlib.h>

namespace ur {

constexpr int kGnThreads = 1024;

__device__ __forceinline__ uint32_t map_to_rank(const void* smem_ptr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

struct GnParams {
  const bf16* x1;
  long long ld1, is1;
  int C1;
  const bf16* x2;
  long long ld2, is2;
  int C2;
  int G, P;
  const float* gamma;
  const float* beta;
  float eps;
  int silu;
  bf16* out;
  long long ldo, iso;
};

__global__ void __launch_bounds__(kGnThreads, 1) group_norm_cluster_kernel(const GnParams p) {
  extern __shared__ __align__(16) float sh[];
  const int C = p.C1 + p.C2;
  const int CL = static_cast<int>(cluster_nctarank());
  const int rank = static_cast<int>(cluster_ctarank());
  const int b = blockIdx.y;
  float* s_sum = sh;                 // [C]   this CTA's partial sums
  float* s_sq = sh + C;              // [C]
  float* s_own = s_sq + C;           // [2 * G / CL]  (mean, rstd) of the groups this CTA owns
  float* s_grp = s_own + 2 * (p.G / CL);   // [2 * G]  all groups

  const int CV = C >> 3;
  const int PL = kGnThreads / CV;                      // pixel lanes (threads beyond CV * PL idle in the pixel loops)
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const bool active = pl < PL;
  // [PL][2][C] per-pixel-lane partials, 16-byte aligned (shared-memory float atomics would serialise PL-fold)
  float* s_part = sh + ((2 * C + 2 * (p.G / CL) + 2 * p.G + 3) & ~3);
  const int c0 = cv * 8;
  const int slab = (p.P + CL - 1) / CL;
  const int p0 = rank * slab;
  const int p1 = min(p.P, p0 + slab);
  const bf16* src = (c0 < p.C1) ? (p.x1 + b * p.is1 + c0) : (p.x2 + b * p.is2 + (c0 - p.C1));
  const long long lds = (c0 < p.C1) ? p.ld1 : p.ld2;

  // ---- phase 1: per-channel partial sums of this CTA's slab
  if (active) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    for (int px = p0 + pl; px < p1; px += 4 * PL) {
      uint4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        v[i] = (px + i * PL < p1) ? __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(px + i * PL) * lds))
                                  : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a, c;
          unpack_bf16(u[j], a, c);
          s[2 * j] += a;
          q[2 * j] += a * a;
          s[2 * j + 1] += c;
          q[2 * j + 1] += c * c;
        }
      }
    }
    float* row = s_part + static_cast<size_t>(pl) * 2 * C + c0;
    *reinterpret_cast<float4*>(row) = make_float4(s[0], s[1], s[2], s[3]);
    *reinterpret_cast<float4*>(row + 4) = make_float4(s[4], s[5], s[6], s[7]);
    *reinterpret_cast<float4*>(row + C) = make_float4(q[0], q[1], q[2], q[3]);
    *reinterpret_cast<float4*>(row + C + 4) = make_float4(q[4], q[5], q[6], q[7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {      // [0, C): sums, [C, 2C): sums of squares (= s_sum | s_sq)
    float t = 0.f;
    for (int l = 0; l < PL; ++l) t += s_part[static_cast<size_t>(l) * 2 * C + i];
    sh[i] = t;
  }
  cluster_sync_all();

  // ---- phase 2: this CTA owns groups [rank * G/CL, (rank+1) * G/CL): one warp per group sums cg channels x CL peers
  const int cg = C / p.G;
  const int gper = p.G / CL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int gi = warp; gi < gper; gi += kGnThreads / 32) {
    const int g = rank * gper + gi;
    double s = 0.0, q = 0.0;
    for (int i = lane; i < cg * CL; i += 32) {
      const int c = g * cg + i % cg;
      const uint32_t r = static_cast<uint32_t>(i / cg);
      s += static_cast<double>(ld_dsmem_f32(map_to_rank(&s_sum[c], r)));
      q += static_cast<double>(ld_dsmem_f32(map_to_rank(&s_sq[c], r)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      const double inv_n = 1.0 / (static_cast<double>(p.P) * cg);
      const double mean = s * inv_n;
      double var = q * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_own[2 * gi] = static_cast<float>(mean);
      s_own[2 * gi + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
    }
  }
  cluster_sync_all();

  // ---- phase 3: gather every group's (mean, rstd) from its owner
  for (int i = threadIdx.x; i < 2 * p.G; i += blockDim.x) {
    const int g = i >> 1;
    s_grp[i] = ld_dsmem_f32(map_to_rank(&s_own[2 * (g % gper) + (i & 1)], static_cast<uint32_t>(g / gper)));
  }
  cluster_sync_all();          // no CTA leaves (or reuses its statistics) while peers still read its shared memory

  // ---- phase 4: apply on the slab (second read hits L2)
  if (!active) return;
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const int g = c / cg;
    const float ga = p.gamma ? __ldg(p.gamma + c) : 1.f;
    const float be = p.beta ? __ldg(p.beta + c) : 0.f;
    sc[j] = s_grp[2 * g + 1] * ga;
    sf[j] = be - s_grp[2 * g] * sc[j];
  }
  bf16* dst = p.out + b * p.iso + c0;
  for (int px = p0 + pl; px < p1; px += 4 * PL) {
    uint4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (px + i * PL < p1) v[i] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(px + i * PL) * lds));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (px + i * PL >= p1) break;
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a, c;
        unpack_bf16(u[j], a, c);
        a = a * sc[2 * j] + sf[2 * j];
        c = c * sc[2 * j + 1] + sf[2 * j + 1];
        if (p.silu) {
          a = silu_f(a);
          c = silu_f(c);
        }
        o[j] = pack_bf16(a, c);
      }
      *reinterpret_cast<uint4*>(dst + static_cast<long long>(px + i * PL) * p.ldo) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// largest usable cluster size (16 needs the non-portable opt-in and enough SMs per GPC), probed once
static int g_gn_cluster = 0;
static int probe_cluster(size_t smem_max) {
  if (g_gn_cluster) return g_gn_cluster;
  cudaFuncSetAttribute(group_norm_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(group_norm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max));
  int best = 8;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(16, 1, 1);
  cfg.blockDim = dim3(kGnThreads);
  cfg.dynamicSmemBytes = smem_max;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 16;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  // measured on B200 (tools/bench_norm.py): 16-CTA clusters are co-scheduled too sparsely (8 is 25-30 % faster), so the
  // non-portable size is only used on request (ur_debug_set_group_norm_cluster)
  if (cudaOccupancyMaxActiveClusters(&n, group_norm_cluster_kernel, &cfg) == cudaSuccess && n >= 8 &&
      getenv("UR_GN_CLUSTER16"))
    best = 16;
  cudaGetLastError();
  g_gn_cluster = best;
  return best;
}

}  // namespace ur

using namespace ur;

extern "C" int ur_group_norm_cluster_size(void) { return probe_cluster(160 * 1024); }
// development: override the cluster size (0 = probe again)
extern "C" int ur_debug_set_group_norm_cluster(int n) {
  probe_cluster(160 * 1024);
  const int old = g_gn_cluster;
  g_gn_cluster = n > 0 ? n : 0;
  return old;
}

extern "C" int ur_group_norm(const void* x1, int64_t ld1, int64_t is1, int c1, const void* x2, int64_t ld2, int64_t is2,
                             int c2, int groups, int batch, int pixels, const float* gamma, const float* beta, float eps,
                             int silu, void* out, int64_t ldo, int64_t iso, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int C = c1 + c2;
  if (!x1 || !out || c1 <= 0 || c1 % 8 || c2 % 8 || (c2 && !x2) || groups <= 0 || C % groups || C > 8 * kGnThreads ||
      ld1 % 8 || ldo % 8 || (c2 && ld2 % 8) || batch <= 0 || pixels <= 0)
    return set_error(UR_ERR_ARG, "ur_group_norm: bad arguments (C=%d+%d groups=%d)", c1, c2, groups);
  int CL = probe_cluster(160 * 1024);
  while (CL > 1 && (groups % CL || pixels < CL)) CL >>= 1;
  GnParams p;
  p.x1 = static_cast<const bf16*>(x1);
  p.ld1 = ld1;
  p.is1 = is1;
  p.C1 = c1;
  p.x2 = static_cast<const bf16*>(x2);
  p.ld2 = ld2;
  p.is2 = is2;
  p.C2 = c2;
  p.G = groups;
  p.P = pixels;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  p.silu = silu;
  p.out = static_cast<bf16*>(out);
  p.ldo = ldo;
  p.iso = iso;
  const int PL = kGnThreads / (C >> 3);
  const size_t smem = sizeof(float) * (2 * static_cast<size_t>(C) + 2 * (groups / CL) + 2 * groups + 8 +
                                      2 * static_cast<size_t>(C) * PL);
  if (smem > 160 * 1024) return set_error(UR_ERR_ARG, "ur_group_norm: too many channels / groups (%d / %d)", C, groups);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, batch, 1);
  cfg.blockDim = dim3(kGnThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, group_norm_cluster_kernel, p);
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "ur_group_norm launch");
}
