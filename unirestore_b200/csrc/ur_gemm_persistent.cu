// Persistent implicit-GEMM convolution / linear kernel (the fast path of ur_conv_gemm).
//
//   one CTA per SM, static round-robin over the (m_tile, n_tile) list (n fastest, so CTAs running side by side
//   share the activation tile in L2);  192 threads:
//     warp 0      TMA producer        smem ring of STAGES x (A 128x64 bf16 + W BNx64 bf16), 128-byte swizzle
//     warp 1      tcgen05.mma issuer  accumulators double-buffered in TMEM (2 x BN fp32 columns): the epilogue of
//                                     tile i overlaps the main loop of tile i+1
//     warps 2..5  epilogue            tcgen05.ld (32 columns per load) -> +bias/+temb (staged in smem) -> activation
//                                     -> channel scale -> +residual -> bf16 -> smem staging (64-byte swizzle)
//                                     -> TMA store (coalesced, clipped at the tensor edge by the hardware)
//
// Same arithmetic and operand layout as ur_gemm.cu (see there / include/unirestore_b200.h for the reference
// call sites); this variant requires a bf16 output whose pitches are multiples of 8 elements.
#include "ur_gemm.h"

namespace ur {

constexpr int kStagePitch = 144;                 // 64 bf16 + 16 B pad: conflict-free row-wise 16-byte accesses
constexpr int kStagingBytes = kBlockM * kStagePitch;

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

template <int BN>
__host__ __device__ constexpr int pstages() {
  return BN == 256 ? 4 : (BN == 160 ? 5 : (BN == 128 ? 6 : 8));
}

template <int BN>
__global__ void __launch_bounds__(320, 1)
conv_gemm_persistent_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ CUtensorMap mapA1,
                            const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW,
                            const __grid_constant__ CUtensorMap mapOut, int total_tiles, int n_tiles) {
  constexpr int STAGES = pstages<BN>();
  constexpr int kWBytes = BN * kBlockK * 2;
  constexpr int kStageBytes = kABytes + kWBytes;
  constexpr uint32_t kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* staging = smem + STAGES * kStageBytes;                       // 128 x 144 B
  float* s_add = reinterpret_cast<float*>(staging + kStagingBytes);      // [2][BN]
  float* s_mul = s_add + 2 * BN;                                         // [2][BN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_mul + 2 * BN);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Wt = 1 << p.wt_log2, Ht = 1 << p.ht_log2;
  const int Bt = kBlockM >> (p.wt_log2 + p.ht_log2);
  const int nkb = p.ntaps * p.cblocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int kiter = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int nt = t % n_tiles;
        int mt = t / n_tiles;
        const int n0 = nt * BN;
        const int tx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tb = mt / p.tiles_y;
        const int x0 = tx * Wt, y0 = ty * Ht, b0 = tb * Bt;
        const int cbase = p.group_kc ? (n0 / p.group_nc) * p.group_kc : 0;
        const int wb = p.w_batched ? b0 : 0;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < nkb; ++kb, ++kiter) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], kStageBytes);
          uint8_t* sa = smem + s * kStageBytes;
          const int dy = static_cast<int>((p.dy_pack >> (4 * tap)) & 15) - 8;
          const int dx = static_cast<int>((p.dx_pack >> (4 * tap)) & 15) - 8;
          const int c = cbase + cb * kBlockK;
          const int xi = x0 * p.stride + dx, yi = y0 * p.stride + dy;
          if (c < p.c1)
            tma_load_4d(sa, &mapA1, &full_bar[s], c, xi, yi, b0);
          else
            tma_load_4d(sa, &mapA2, &full_bar[s], c - p.c1, xi, yi, b0);
          tma_load_3d(sa + kABytes, &mapW, &full_bar[s], tap * p.kc + cb * kBlockK, n0, wb);
          if (++cb == p.cblocks) {
            cb = 0;
            ++tap;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN);
      int kiter = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < nkb; ++kb, ++kiter) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * kStageBytes);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) tc_mma_bf16(tacc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[s]);
        }
        tc_commit(&tfull_bar[as]);
      }
    }
  } else {
    // =============================== epilogue ===============================
    // 8 warps: two per TMEM lane quarter; warp group g = (warp - 2) / 4 handles the 32-column sub-blocks
    // g, g + 2, g + 4, ...  Each thread finishes 32 columns of ITS row (TMEM lane) entirely in registers and
    // writes 64 contiguous bytes; the residual of the next sub-block is prefetched while the current one is
    // being processed.  No shared-memory staging, no intra-tile barriers.
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;             // 0 or 1
    const int r = q * 32 + lane;                 // tile row = TMEM lane
    const int et = threadIdx.x - 64;             // 0..255
    const bool gated = p.act == UR_ACT_GEGLU || p.act == UR_ACT_GATE;
    const int n_out = gated ? (p.N >> 1) : p.N;
    const int ncols = gated ? (BN >> 1) : BN;
    const int xl = r & (Wt - 1);
    const int yl = (r >> p.wt_log2) & (Ht - 1);
    const int bl = r >> (p.wt_log2 + p.ht_log2);
    const bool has_mul = p.chscale != nullptr;
    const bool has_alpha = p.alpha != 1.0f;
    bf16* outp = reinterpret_cast<bf16*>(p.out);
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int nt = t % n_tiles;
      int mt = t / n_tiles;
      const int n0 = nt * BN;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int tb = mt / p.tiles_y;
      const int x0 = tx * Wt, y0 = ty * Ht, b0 = tb * Bt;
      const int nout0 = gated ? (n0 >> 1) : n0;
      const int as = it & 1;
      // ---- stage the per-column add / mul vectors of this tile (double-buffered by `as`)
      float* add = s_add + as * BN;
      float* mul = s_mul + as * BN;
      if (et < BN) {
        const int c = et;
        const int n = n0 + c;
        float a = 0.f;
        if (n < p.N) {
          if (p.bias) a += __ldg(p.bias + n);
          if (p.rowvec) a += __ldg(p.rowvec + b0 * p.rowvec_sb + n);
        }
        add[c] = a;
        float m = 1.f;
        if (has_mul && c < ncols && nout0 + c < n_out) m = __ldg(p.chscale + b0 * p.chscale_sb + nout0 + c);
        mul[c] = m;
      }
      const int x = x0 + xl, y = y0 + yl, b = b0 + bl;
      const bool row_ok = (x < p.Wo) && (y < p.Ho) && (b < p.B);
      bf16* orow = outp + b * p.out_sb + y * p.out_sy + x * p.out_sx + nout0;
      const bf16* rrow = (p.residual && row_ok) ? p.residual + b * p.res_sb + y * p.res_sy + x * p.res_sx + nout0 : nullptr;
      // residual prefetch of this warp group's first sub-block
      uint4 rv[4];
      int c = grp * 32;
      if (rrow != nullptr && c < ncols && nout0 + c + 32 <= n_out) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rv[j] = __ldg(reinterpret_cast<const uint4*>(rrow + c) + j);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull_bar[as], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);

      for (; c < ncols; c += 64) {
        uint32_t va[32], vg[32];
        tmem_ld32(trow + c, va);
        if (gated) tmem_ld32(trow + (BN >> 1) + c, vg);
        // prefetch the residual of the next sub-block of this warp group
        uint4 rn[4];
        const int cn = c + 64;
        const bool pre = rrow != nullptr && cn < ncols && nout0 + cn + 32 <= n_out;
        if (pre) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rn[j] = __ldg(reinterpret_cast<const uint4*>(rrow + cn) + j);
        }
        tmem_ld_wait();
        const bool full = nout0 + c + 32 <= n_out;
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 ad = *reinterpret_cast<const float4*>(add + c + 4 * j);
          float f[4] = {__uint_as_float(va[4 * j]), __uint_as_float(va[4 * j + 1]), __uint_as_float(va[4 * j + 2]),
                        __uint_as_float(va[4 * j + 3])};
          if (has_alpha) {
#pragma unroll
            for (int k = 0; k < 4; ++k) f[k] *= p.alpha;
          }
          f[0] += ad.x;
          f[1] += ad.y;
          f[2] += ad.z;
          f[3] += ad.w;
          if (gated) {
            const float4 ag = *reinterpret_cast<const float4*>(add + (BN >> 1) + c + 4 * j);
            float g[4] = {__uint_as_float(vg[4 * j]), __uint_as_float(vg[4 * j + 1]), __uint_as_float(vg[4 * j + 2]),
                          __uint_as_float(vg[4 * j + 3])};
            if (has_alpha) {
#pragma unroll
              for (int k = 0; k < 4; ++k) g[k] *= p.alpha;
            }
            g[0] += ag.x;
            g[1] += ag.y;
            g[2] += ag.z;
            g[3] += ag.w;
            if (p.act == UR_ACT_GEGLU) {
#pragma unroll
              for (int k = 0; k < 4; ++k) f[k] *= gelu_erf_f(g[k]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) f[k] *= g[k];
            }
          } else if (p.act == UR_ACT_SILU) {
#pragma unroll
            for (int k = 0; k < 4; ++k) f[k] = silu_f(f[k]);
          } else if (p.act == UR_ACT_GELU) {
#pragma unroll
            for (int k = 0; k < 4; ++k) f[k] = gelu_erf_f(f[k]);
          }
          if (has_mul) {
            const float4 mu = *reinterpret_cast<const float4*>(mul + c + 4 * j);
            f[0] *= mu.x;
            f[1] *= mu.y;
            f[2] *= mu.z;
            f[3] *= mu.w;
          }
          if (rrow != nullptr && full) {
            const uint4 rq = rv[j >> 1];
            const uint32_t u0 = (j & 1) ? rq.z : rq.x, u1 = (j & 1) ? rq.w : rq.y;
            float a0, a1;
            unpack_bf16(u0, a0, a1);
            f[0] += a0;
            f[1] += a1;
            unpack_bf16(u1, a0, a1);
            f[2] += a0;
            f[3] += a1;
          }
          o[2 * j] = pack_bf16(f[0], f[1]);
          o[2 * j + 1] = pack_bf16(f[2], f[3]);
        }
        if (row_ok) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              reinterpret_cast<uint4*>(orow + c)[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          } else {                                  // ragged tail of the channel dimension (n_out % 32 != 0)
            for (int j = 0; j < 32; ++j) {
              if (nout0 + c + j < n_out) {
                float lo, hi;
                unpack_bf16(o[j >> 1], lo, hi);
                float v = (j & 1) ? hi : lo;
                if (rrow != nullptr) v += __bfloat162float(rrow[c + j]);
                orow[c + j] = __float2bfloat16(v);
              }
            }
          }
        }
        if (pre) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rv[j] = rn[j];
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN>
static int launch_p(const GemmParams& p, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w,
                    const CUtensorMap& out, int total_tiles, int n_tiles, cudaStream_t stream) {
  constexpr int smem = pstages<BN>() * (kABytes + BN * kBlockK * 2) + kStagingBytes + 4 * BN * 4 + 512;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_persistent_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_gemm_persistent)");
    configured = true;
  }
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  conv_gemm_persistent_kernel<BN><<<grid, 320, smem, stream>>>(p, a1, a2, w, out, total_tiles, n_tiles);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "conv_gemm_persistent launch");
}

int launch_conv_gemm_persistent(const GemmParams& p, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w,
                                const CUtensorMap& out, int bn, int total_tiles, int n_tiles, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_p<64>(p, a1, a2, w, out, total_tiles, n_tiles, stream);
    case 128: return launch_p<128>(p, a1, a2, w, out, total_tiles, n_tiles, stream);
    case 160: return launch_p<160>(p, a1, a2, w, out, total_tiles, n_tiles, stream);
    default: return launch_p<256>(p, a1, a2, w, out, total_tiles, n_tiles, stream);
  }
}

}  // namespace ur
