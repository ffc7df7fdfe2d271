// HBM-bound normalisation kernels on bf16 channels-last activations (128-bit vectorised):
//   ur_chan_stats    per-(image, channel) sum / sum-of-squares in fp64   (GroupNorm / InstanceNorm / GAP)
//   ur_norm_apply    GroupNorm / InstanceNorm apply (+affine, +SiLU), concatenating two sources
//   ur_layernorm     per-token LayerNorm over C (nn.LayerNorm in BasicTransformerBlock, timm LayerNorm2d)
//   ur_scale_channels  x[b, p, c] *= s[b, c]   (NAFBlock SCA nafnet_arch.py:122, AdaNAFV2 gates cfrm.py:49-52)
#include "ur_common.cuh"
#include "ur_host.h"

#include <stdlib.h>

namespace ur {

// ------------------------------------------------------------------------------------ chan_stats
// grid (chunks, B); block = CV * PL threads: thread (cv, pl) owns channel vector cv (8 channels)
// and pixels p0+pl, p0+pl+PL, ...  fp32 per-thread partials -> shared fp32 atomics -> fp64 global atomics.
__global__ void chan_stats_kernel(const bf16* __restrict__ x, long long ld, long long img_stride, int P, int C, int CV,
                                  int PL, int chunk, double* __restrict__ stats, int stats_ld, int stats_off) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float sh[];  // [PL][2][C] per-pixel-lane partials (no shared-memory float atomics:
                                               // they compile to CAS loops that serialise PL-fold per channel)
  const int b = blockIdx.y;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(P, p0 + chunk);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const bf16* base = x + b * img_stride + cv * 8;
  for (int p = p0 + pl; p < p1; p += 4 * PL) {
    uint4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      v[i] = (p + i * PL < p1) ? __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(p + i * PL) * ld))
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
  float* row = sh + static_cast<size_t>(pl) * 2 * C + cv * 8;
  *reinterpret_cast<float4*>(row) = make_float4(s[0], s[1], s[2], s[3]);
  *reinterpret_cast<float4*>(row + 4) = make_float4(s[4], s[5], s[6], s[7]);
  *reinterpret_cast<float4*>(row + C) = make_float4(q[0], q[1], q[2], q[3]);
  *reinterpret_cast<float4*>(row + C + 4) = make_float4(q[4], q[5], q[6], q[7]);
  __syncthreads();
  double* out = stats + (static_cast<long long>(b) * stats_ld + stats_off) * 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {      // i < C: sum of channel i, else sumsq of channel i - C
    float t = 0.f;
    for (int l = 0; l < PL; ++l) t += sh[static_cast<size_t>(l) * 2 * C + i];
    const int c = i < C ? i : i - C;
    atomicAdd(out + 2 * c + (i < C ? 0 : 1), static_cast<double>(t));
  }
}

// ------------------------------------------------------------------------------------ norm_apply
__global__ void norm_apply_kernel(const bf16* __restrict__ x1, long long ld1, long long is1, int C1,
                                  const bf16* __restrict__ x2, long long ld2, long long is2, int C2,
                                  const double* __restrict__ stats, const double* __restrict__ stats2, int G, int P,
                                  int CV, int PL, int chunk,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                                  bf16* __restrict__ out, long long ldo, long long iso) {
  // The small tensors of the low-resolution levels (8 x 8 / 16 x 16 pixels, 1280..2560 channels: ~800 launches per
  // forward) are a pure latency chain -- statistics -> group reduction -> affine parameters -> data -> store took 12 us
  // for 10 MB (r2 chain benchmark).  So: the affine parameters (weights, not written by the predecessor) are fetched
  // BEFORE the programmatic-dependency wait; the first batch of activations and the statistics leave together right
  // after it (one L2 round trip instead of three); the fp64 group reduction runs on several threads per group.
  pdl_launch_dependents();
  extern __shared__ __align__(16) double shd[];  // per-channel (sum, sumsq) [C][2] as doubles, then mean[G], rstd[G] floats
  float* sh = reinterpret_cast<float*>(shd + 2 * (C1 + C2));
  const int C = C1 + C2;
  const int cg = C / G;
  const int b = blockIdx.y;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int c0 = cv * 8;
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = gamma ? __ldg(gamma + c0 + j) : 1.f;
    be[j] = beta ? __ldg(beta + c0 + j) : 0.f;
  }
  pdl_wait();
  const bf16* src = (c0 < C1) ? (x1 + b * is1 + c0) : (x2 + b * is2 + (c0 - C1));
  const long long lds = (c0 < C1) ? ld1 : ld2;
  bf16* dst = out + b * iso + c0;
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(P, p0 + chunk);
  // first batch of this thread's pixels: four independent 16-byte loads in flight while the statistics arrive
  uint4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (p0 + pl + i * PL < p1) v[i] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(p0 + pl + i * PL) * lds));
  // statistics of channel c: one array over cat(x1, x2), or one array per source.  Every thread fetches whole
  // (sum, sumsq) pairs in ONE round trip to L2 (16-byte loads), then TPG threads per group reduce from shared memory
  // (a serial walk over the cg channels of a group by one thread was cg dependent fp64 adds at the head of every block)
  const double* st1 = stats + static_cast<long long>(b) * (stats2 ? C1 : C) * 2;
  const double* st2 = stats2 ? stats2 + static_cast<long long>(b) * C2 * 2 - 2 * C1 : st1;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double2 sv = *reinterpret_cast<const double2*>((c < C1 ? st1 : st2) + 2 * c);
    shd[2 * c] = sv.x;
    shd[2 * c + 1] = sv.y;
  }
  __syncthreads();
  {
    // TPG = 1, 2, 4 or 8 threads per group (a power of two: the partial sums meet by warp shuffles inside aligned lane
    // clusters), as many as the block has: blockDim.x >= 32 always, G <= blockDim.x * ... handled by the outer loop
    int tpg = 8;
    while (tpg > 1 && G * tpg > static_cast<int>(blockDim.x & ~31u)) tpg >>= 1;
    const double inv_n = 1.0 / (static_cast<double>(P) * cg);
    const int slots = static_cast<int>(blockDim.x & ~31u) / tpg;          // groups handled per pass (whole warps only)
    for (int g0 = 0; g0 < G; g0 += slots) {
      const int g = g0 + static_cast<int>(threadIdx.x) / tpg, sub = threadIdx.x % tpg;
      const bool active = threadIdx.x < (blockDim.x & ~31u) && g < G;
      double s = 0.0, q = 0.0;
      if (active) {
        for (int c = g * cg + sub; c < (g + 1) * cg; c += tpg) {
          s += shd[2 * c];
          q += shd[2 * c + 1];
        }
      }
      if (threadIdx.x < (blockDim.x & ~31u)) {         // whole warps: the shuffles below name every lane
        for (int o = tpg >> 1; o > 0; o >>= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
      }
      if (active && sub == 0) {
        const double mean = s * inv_n;
        double var = q * inv_n - mean * mean;
        if (var < 0.0) var = 0.0;
        sh[g] = static_cast<float>(mean);
        sh[G + g] = rsqrtf(static_cast<float>(var) + eps);
      }
    }
  }
  __syncthreads();
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c0 + j) / cg;
    sc[j] = sh[G + g] * ga[j];
    sf[j] = be[j] - sh[g] * sc[j];
  }
  for (int p = p0 + pl; p < p1; p += 4 * PL) {
    if (p != p0 + pl) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (p + i * PL < p1) v[i] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(p + i * PL) * lds));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (p + i * PL >= p1) break;
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // packed f32x2 arithmetic: one FMA for the pair, and the SiLU's multiplies / add around its two MUFU per element
        float a, c;
        unpack_bf16(u[j], a, c);
        f2 y = fma2(f2{a, c}, f2{sc[2 * j], sc[2 * j + 1]}, f2{sf[2 * j], sf[2 * j + 1]});
        if (silu) {
          const f2 t = mul2(y, splat2(-1.4426950408889634f));                   // -x log2(e)
          const f2 d = fma2(f2{exp2_approx(t.x), exp2_approx(t.y)}, splat2(1.0f), splat2(1.0f));   // 1 + e^-x
          y = mul2(y, f2{rcp_approx(d.x), rcp_approx(d.y)});
        }
        o[j] = pack_bf16(y.x, y.y);
      }
      *reinterpret_cast<uint4*>(dst + static_cast<long long>(p + i * PL) * ldo) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------ layernorm
// One warp per R consecutive tokens; the rows live in registers (C <= 8*32*NV), exact two-pass mean / variance.
// All R * NV 16-byte loads of a warp are issued before the first reduction (the kernel is latency-bound otherwise).
// The rows stay PACKED (bf16, as loaded) and are unpacked in each of the three passes: held as fp32 they took twice
// the registers, 4 blocks per SM fitted and the ~40 KB per SM in flight bounded the kernel at 3.2 TB/s (r2 chain
// benchmark); the unpacking is two ALU instructions per pair.
template <int NV, int R>
__global__ void __launch_bounds__(128, NV <= 2 ? 8 : 1) layernorm_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ out, long long ldo,
                                 int M, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  if (row0 >= M) return;
  const int nvec = C >> 3;
  uint4 raw[R][NV];
  float sum[R], sq[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool row_ok = row0 + r < M;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      raw[r][i] = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok && vi < nvec) raw[r][i] = __ldg(reinterpret_cast<const uint4*>(x + (row0 + r) * ldx + vi * 8));
    }
  }
  // (volatile asm: the compiler would otherwise unpack once and keep -- or spill -- the fp32 copies); the arithmetic
  // of all three passes runs on packed f32x2 instructions, one per PAIR of elements (the kernel is issue-bound next to
  // its memory time: ~11 scalar instructions per element before, ~6 now)
  auto unpack2 = [](uint32_t u) {
    uint32_t lo, hi;
    asm volatile("shl.b32 %0, %2, 16;\n\tand.b32 %1, %2, 0xffff0000;" : "=r"(lo), "=r"(hi) : "r"(u));
    return f2{__uint_as_float(lo), __uint_as_float(hi)};
  };
  const f2 one = splat2(1.0f);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    f2 acc = f2{0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t u[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) acc = fma2(unpack2(u[j]), one, acc);          // padded vectors are zero
    }
    sum[r] = acc.x + acc.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < R; ++r) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float mean = sum[r] / C;
    sum[r] = mean;
    const f2 nm = splat2(-mean);
    f2 acc = f2{0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
        const uint32_t u[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const f2 d = fma2(unpack2(u[j]), one, nm);
          acc = fma2(d, d, acc);
        }
      }
    }
    sq[r] = acc.x + acc.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < R; ++r) sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], o);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
      const f2 ga[4] = {f2{g0.x, g0.y}, f2{g0.z, g0.w}, f2{g1.x, g1.y}, f2{g1.z, g1.w}};
      const f2 be[4] = {f2{b0.x, b0.y}, f2{b0.z, b0.w}, f2{b1.x, b1.y}, f2{b1.z, b1.w}};
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (row0 + r < M) {
          const f2 nm = splat2(-sum[r]);
          const f2 rs = splat2(rsqrtf(sq[r] / C + eps));
          const uint32_t u[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const f2 t = mul2(fma2(unpack2(u[j]), one, nm), rs);       // (x - mean) * rstd
            const f2 y = fma2(t, ga[j], be[j]);
            o[j] = pack_bf16(y.x, y.y);
          }
          *reinterpret_cast<uint4*>(out + (row0 + r) * ldo + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------ scale_channels
__global__ void scale_channels_kernel(bf16* __restrict__ x, long long ld, long long img_stride, int P, int CV,
                                      const float* __restrict__ s, int s_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const long long total = static_cast<long long>(P) * CV;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % CV);
    const long long p = i / CV;
    bf16* ptr = x + b * img_stride + p * ld + cv * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(ptr);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    const float* sp = s + static_cast<long long>(b) * s_ld + cv * 8;
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a, c;
      unpack_bf16(u[j], a, c);
      o[j] = pack_bf16(a * __ldg(sp + 2 * j), c * __ldg(sp + 2 * j + 1));
    }
    *reinterpret_cast<uint4*>(ptr) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

static void pick_block(int C, int P, int& CV, int& PL, int& chunk, int& nchunks, int B) {
  CV = C / 8;
  PL = CV >= 256 ? 1 : 256 / CV;
  if (PL < 1) PL = 1;
  // aim for ~8 waves of blocks over the GPU (same-box A/B in the full forward: 8 > 4 > 2), at least 8 pixels per thread
  static const int waves = getenv("UR_NORM_WAVES") ? atoi(getenv("UR_NORM_WAVES")) : 8;
  const int target = max(1, (waves * num_sms()) / max(1, B));
  chunk = (P + target - 1) / target;
  static const int min_iters = getenv("UR_NORM_MIN_ITERS") ? atoi(getenv("UR_NORM_MIN_ITERS")) : 2;   // x 4 PL pixels per block at least
  const int min_chunk = PL * 4 * min_iters;
  if (chunk < min_chunk) chunk = min_chunk;
  nchunks = (P + chunk - 1) / chunk;
}

}  // namespace ur

using namespace ur;

extern "C" int ur_chan_stats(const void* x, int64_t ld, int64_t img_stride, int batch, int pixels, int channels,
                             double* stats, int stats_ld, int stats_off, int zero_first, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!x || !stats || channels % 8 || channels > 8192 || ld % 8 || img_stride % 8 || batch <= 0 || pixels <= 0)
    return set_error(UR_ERR_ARG, "ur_chan_stats: bad arguments (C=%d ld=%lld)", channels, (long long)ld);
  if (zero_first) {
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * static_cast<size_t>(batch) * stats_ld, stream);
    if (e != cudaSuccess) return set_cuda_error(e, "ur_chan_stats memset");
  }
  int CV, PL, chunk, nchunks;
  pick_block(channels, pixels, CV, PL, chunk, nchunks, batch);
  dim3 grid(nchunks, batch);
  launch_kernel(chan_stats_kernel, dim3(grid), dim3(CV * PL), 2 * static_cast<size_t>(channels) * PL * sizeof(float), stream, 
      static_cast<const bf16*>(x), ld, img_stride, pixels, channels, CV, PL, chunk, stats, stats_ld, stats_off);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "ur_chan_stats launch");
}

extern "C" int ur_norm_apply(const void* x1, int64_t ld1, int64_t is1, int c1, const void* x2, int64_t ld2,
                             int64_t is2, int c2, const double* stats, const double* stats2, int groups, int batch,
                             int pixels,
                             const float* gamma, const float* beta, float eps, int silu, void* out, int64_t ldo,
                             int64_t iso, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int C = c1 + c2;
  if (!x1 || !stats || !out || c1 % 8 || c2 % 8 || (c2 && !x2) || groups <= 0 || C % groups || C > 8192 || ld1 % 8 ||
      ldo % 8 || (c2 && ld2 % 8))
    return set_error(UR_ERR_ARG, "ur_norm_apply: bad arguments (C=%d+%d groups=%d)", c1, c2, groups);
  int CV, PL, chunk, nchunks;
  pick_block(C, pixels, CV, PL, chunk, nchunks, batch);
  static bool configured = false;
  if (!configured) {          // up to C = 8192 channels of fp64 statistics staged in shared memory
    cudaError_t e = cudaFuncSetAttribute(norm_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(norm_apply)");
    configured = true;
  }
  const size_t smem = 2 * groups * sizeof(float) + 2 * static_cast<size_t>(C) * sizeof(double);
  // Grid = ONE full wave of resident blocks when the tensor is big enough for it (round 2, chain benchmark): with the
  // "8 waves" rule the 64 x 64 level launched 688 blocks where 444 are resident (72 registers, 240 threads: 3 per SM) --
  // 1.55 waves, the second one half empty: 14.1 -> 12.6 us (C = 320), 26.7 -> 21.4 (640), 24.6 -> 18.6 (1920 @ 32 x 32).
  {
    static int occ_cache[33][2];                  // [warps per block][smem class] -> resident blocks per SM (0 = not asked yet)
    const int warps = (CV * PL + 31) / 32;
    const int cls = smem > 24 * 1024 ? 1 : 0;
    int bps = warps <= 32 ? occ_cache[warps][cls] : 0;
    if (!bps) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, norm_apply_kernel, CV * PL, smem) != cudaSuccess || bps < 1) bps = 1;
      if (warps <= 32 && !cls) occ_cache[warps][cls] = bps;      // (the large-smem class depends on C: not cached)
    }
    const int per_image = (bps * num_sms()) / (batch > 0 ? batch : 1);
    if (per_image >= 1) {
      const int one_wave = (pixels + per_image - 1) / per_image;          // pixels per block for exactly one wave
      if (one_wave > chunk) {
        chunk = one_wave;
        nchunks = (pixels + chunk - 1) / chunk;
      }
    }
  }
  dim3 grid(nchunks, batch);
  launch_kernel(norm_apply_kernel, dim3(grid), dim3(CV * PL), smem, stream, 
      static_cast<const bf16*>(x1), ld1, is1, c1, static_cast<const bf16*>(x2), ld2, is2, c2, stats, stats2, groups, pixels, CV,
      PL, chunk, gamma, beta, eps, silu, static_cast<bf16*>(out), ldo, iso);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "ur_norm_apply launch");
}

extern "C" int ur_layernorm(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int channels,
                            const float* gamma, const float* beta, float eps, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!x || !out || !gamma || !beta || channels % 8 || channels > 2048 || ldx % 8 || ldo % 8 || rows <= 0)
    return set_error(UR_ERR_ARG, "ur_layernorm: bad arguments (C=%d)", channels);
  if ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15)
    return set_error(UR_ERR_ARG, "ur_layernorm: gamma / beta must be 16-byte aligned");
  const int warps = 4;
  const bf16* xp = static_cast<const bf16*>(x);
  bf16* op = static_cast<bf16*>(out);
  const int M = static_cast<int>(rows);
  auto blocks = [&](int r) { return static_cast<unsigned>((rows + static_cast<int64_t>(warps) * r - 1) / (warps * r)); };
  if (channels <= 256)
    launch_kernel(layernorm_kernel<1, 4>, dim3(blocks(4)), dim3(warps * 32), 0, stream, xp, ldx, op, ldo, M, channels, gamma, beta, eps);
  else if (channels <= 512)
    launch_kernel(layernorm_kernel<2, 4>, dim3(blocks(4)), dim3(warps * 32), 0, stream, xp, ldx, op, ldo, M, channels, gamma, beta, eps);
  else if (channels <= 1280)
    launch_kernel(layernorm_kernel<5, 2>, dim3(blocks(2)), dim3(warps * 32), 0, stream, xp, ldx, op, ldo, M, channels, gamma, beta, eps);
  else
    launch_kernel(layernorm_kernel<8, 1>, dim3(blocks(1)), dim3(warps * 32), 0, stream, xp, ldx, op, ldo, M, channels, gamma, beta, eps);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "ur_layernorm launch");
}

extern "C" int ur_scale_channels(void* x, int64_t ld, int64_t img_stride, int batch, int pixels, int channels,
                                 const float* scale, int scale_ld, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!x || !scale || channels % 8 || ld % 8)
    return set_error(UR_ERR_ARG, "ur_scale_channels: bad arguments");
  const int CV = channels / 8;
  const long long total = static_cast<long long>(pixels) * CV;
  long long gxl = (total + 255) / 256;
  if (gxl > 8LL * num_sms()) gxl = 8LL * num_sms();
  int gx = static_cast<int>(gxl);
  if (gx < 1) gx = 1;
  launch_kernel(scale_channels_kernel, dim3(dim3(gx, batch)), dim3(256), 0, stream, static_cast<bf16*>(x), ld, img_stride, pixels, CV, scale,
                                                           scale_ld);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "ur_scale_channels launch");
}
