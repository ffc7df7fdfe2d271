// ur_conv_gemm: implicit-GEMM convolution / linear layer on the 5th-gen tensor cores.
//
//   D[128 pixels, BN channels] (fp32, TMEM) += A[128, 64] (bf16 smem, TMA) * W[BN, 64]^T (bf16 smem, TMA)
//
// * The activation operand is NEVER im2col-materialised: for k-block (tap, c0) the producer issues one
//   4-D tiled TMA load of the box (64 ch, Wt, Ht, Bt) at coordinates (c0, x0*s+dx, y0*s+dy, b0) of the
//   NHWC tensor.  Out-of-image coordinates are zero-filled by the TMA unit (= conv zero padding); stride-2
//   convolutions use the tensor map's elementStrides.  The box lands in shared memory as 128 rows of
//   128 B with the 128-byte swizzle, which is exactly the K-major SW128 UMMA operand layout.
// * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
//   tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> registers -> fused epilogue -> HBM).
// * smem ring of STAGES x (A 16 KB + W BN*128 B), full/empty mbarriers, tcgen05.commit frees slots.
//
// Reference arithmetic implemented: every nn.Conv2d / nn.Linear on the UniRestore hot path
// (see include/unirestore_b200.h for the call-site list).
#include "ur_gemm.h"

#include <math.h>
#include <stdlib.h>

namespace ur {

template <int BN>
__host__ __device__ constexpr int tmem_cols() {
  return BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
}

template <int BN, int STAGES, int MINB>
__global__ void __launch_bounds__(192, MINB)
conv_gemm_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW) {
  constexpr int kWBytes = BN * kBlockK * 2;
  constexpr int kStageBytes = kABytes + kWBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = warp_uniform(static_cast<int>(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int n0 = blockIdx.x * BN;
  int mt = blockIdx.y;
  const int tx = mt % p.tiles_x;
  mt /= p.tiles_x;
  const int ty = mt % p.tiles_y;
  const int tb = mt / p.tiles_y;
  const int Wt = 1 << p.wt_log2, Ht = 1 << p.ht_log2;
  const int Bt = kBlockM >> (p.wt_log2 + p.ht_log2);
  const int x0 = tx * Wt, y0 = ty * Ht, b0 = tb * Bt;
  const int nkb = p.ntaps * p.cblocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols<BN>());
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // =============================== TMA producer (whole warp; the elected lane issues) ===============================
    const int cbase = p.group_kc ? (n0 / p.group_nc) * p.group_kc : 0;
    const int wb = p.w_batched ? b0 : 0;
    int tap = 0, cb = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sa = smem + s * kStageBytes;
      const int dy = static_cast<int>((p.dy_pack >> (4 * tap)) & 15) - 8;
      const int dx = static_cast<int>((p.dx_pack >> (4 * tap)) & 15) - 8;
      const int c = cbase + cb * kBlockK;
      const int xi = x0 * p.stride + dx, yi = y0 * p.stride + dy;
      const bool src1 = c < p.c1;
      const CUtensorMap* mp = src1 ? &mapA1 : &mapA2;
      const int cc = src1 ? c : c - p.c1;
      if (elect_one_sync()) {
        mbar_expect_tx(&full_bar[s], kStageBytes);
        tma_load_4d(sa, mp, &full_bar[s], cc, xi, yi, b0);
        tma_load_3d(sa + kABytes, &mapW, &full_bar[s], tap * p.kc + cb * kBlockK, n0, wb);
      }
      __syncwarp();
      if (++cb == p.cblocks) {
        cb = 0;
        ++tap;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp; the elected lane issues) ===============================
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN);
    const uint32_t smem_base = smem_u32(smem);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t sa = smem_base + s * kStageBytes;
      const uint64_t da = umma_desc_k_sw128(sa);
      const uint64_t db = umma_desc_k_sw128(sa + kABytes);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
          tc_mma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) tc_commit(accum_bar);
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // row of the tile = TMEM lane
    const int xl = r & (Wt - 1);
    const int yl = (r >> p.wt_log2) & (Ht - 1);
    const int bl = r >> (p.wt_log2 + p.ht_log2);
    const int x = x0 + xl, y = y0 + yl, b = b0 + bl;
    const bool row_ok = (x < p.Wo) && (y < p.Ho) && (b < p.B);
    const bool gated = p.act == UR_ACT_GEGLU || p.act == UR_ACT_GATE;
    const int n_out = gated ? (p.N >> 1) : p.N;
    const int ncols = gated ? (BN >> 1) : BN;
    const int nout0 = gated ? (n0 >> 1) : n0;
    const long long ooff = b * p.out_sb + y * p.out_sy + x * p.out_sx;
    const long long roff = b * p.res_sb + y * p.res_sy + x * p.res_sx;
    const float* rowvec = p.rowvec ? p.rowvec + b * p.rowvec_sb : nullptr;
    const float* chscale = p.chscale ? p.chscale + b * p.chscale_sb : nullptr;

    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    for (int c = 0; c < ncols; c += 16) {
      uint32_t va[16], vg[16];
      tmem_ld16(trow + c, va);
      if (gated) tmem_ld16(trow + (BN >> 1) + c, vg);
      tmem_ld_wait();
      const int no = nout0 + c;   // first output column of this chunk
      if (row_ok && no < n_out) {
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int na = n0 + c + j;  // GEMM column of the `a` value (bias / rowvec index)
        float v = __uint_as_float(va[j]) * p.alpha;
        const bool col_ok = (no + j) < n_out;
        if (col_ok) {
          if (p.bias) v += __ldg(p.bias + na);
          if (rowvec) v += __ldg(rowvec + na);
          if (gated) {
            const int ng = na + (BN >> 1);
            float g = __uint_as_float(vg[j]) * p.alpha;
            if (p.bias) g += __ldg(p.bias + ng);
            if (rowvec) g += __ldg(rowvec + ng);
            v = (p.act == UR_ACT_GEGLU) ? v * gelu_erf_f(g) : v * g;
          } else if (p.act == UR_ACT_SILU) {
            v = silu_f(v);
          } else if (p.act == UR_ACT_GELU) {
            v = gelu_erf_f(v);
          }
          if (chscale) v *= __ldg(chscale + no + j);
        }
        f[j] = v;
      }
      const bool full = (no + 16) <= n_out;
      if (p.residual) {
        const bf16* rp = p.residual + roff + no;
        if (full && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
          const uint4 r0 = *reinterpret_cast<const uint4*>(rp);
          const uint4 r1 = *reinterpret_cast<const uint4*>(rp + 8);
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a, bb;
            unpack_bf16(rr[j], a, bb);
            f[2 * j] += a;
            f[2 * j + 1] += bb;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (no + j < n_out) f[j] += __bfloat162float(rp[j]);
        }
      }
      if (p.out_f32) {
        float* op = reinterpret_cast<float*>(p.out) + ooff + no;
        if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            reinterpret_cast<float4*>(op)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (no + j < n_out) op[j] = f[j];
        }
      } else {
        bf16* op = reinterpret_cast<bf16*>(p.out) + ooff + no;
        if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
          uint4 o0, o1;
          o0.x = pack_bf16(f[0], f[1]);
          o0.y = pack_bf16(f[2], f[3]);
          o0.z = pack_bf16(f[4], f[5]);
          o0.w = pack_bf16(f[6], f[7]);
          o1.x = pack_bf16(f[8], f[9]);
          o1.y = pack_bf16(f[10], f[11]);
          o1.z = pack_bf16(f[12], f[13]);
          o1.w = pack_bf16(f[14], f[15]);
          reinterpret_cast<uint4*>(op)[0] = o0;
          reinterpret_cast<uint4*>(op)[1] = o1;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (no + j < n_out) op[j] = __float2bfloat16(f[j]);
        }
      }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols<BN>());
  }
}

// ------------------------------------------------------------------------------------------ host
template <int BN, int STAGES, int MINB>
static int launch_conv_gemm(const GemmParams& p, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w,
                            dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (kABytes + BN * kBlockK * 2) + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_gemm)");
    configured = true;
  }
  cudaError_t e = launch_kernel(conv_gemm_kernel<BN, STAGES, MINB>, grid, dim3(192), smem, stream, p, a1, a2, w);
  if (e != cudaSuccess) return set_cuda_error(e, "conv_gemm launch");
  return UR_OK;
}


}  // namespace ur

using namespace ur;

static int g_force_v1 = getenv("UR_GEMM_V1") != nullptr;
static long long* g_trace = nullptr;
// development: device buffer of 128 int64 receiving per-role clock64 timestamps of CTA 0 (nullptr = off)
extern "C" int ur_debug_set_gemm_trace(void* buf) {
  g_trace = static_cast<long long*>(buf);
  return 0;
}
// development switch: 1 = route every ur_conv_gemm call to the one-tile-per-CTA kernel
extern "C" int ur_debug_force_gemm_v1(int on) {
  const int old = g_force_v1;
  g_force_v1 = on;
  return old;
}

extern "C" int ur_conv_gemm_pick_bn(int n, int gated) {
  if (gated) {   // M-independent (weights are packed for it); half tiles must be multiples of 32 columns
    if (n % 256 == 0) return 256;
    if (n % 128 == 0) return 128;
    return n % 64 == 0 ? 64 : 0;
  }
  if (n % 160 == 0) return 160;
  if (n % 256 == 0 && n >= 1024) return 256;
  if (n % 128 == 0) return 128;
  if (n <= 64) return 64;
  if (n % 64 == 0 && n < 256) return 64;
  return 128;
}

// Tile configuration (N tile, single CTA vs CTA pair) minimising (rounds over the SMs) x (time per k-block).
// Measured model of the main loop (cycles per 64-deep k-block): the UMMA itself 2*bn (128 rows per SM), the
// shared-memory traffic of TMA write + UMMA read, (16 KB + W bytes) * 2 / 128 B per cycle, and ~75 cycles of issue
// per tcgen05.mma.  A pair stages only bn/2 weight rows per CTA.
// epilogue stores: 0 = st.global, 1 = one TMA store per 128-row sub-block (group barrier), 2 = one per warp (default)
static int g_lean_epilogue = getenv("UR_GEMM_LEAN") ? atoi(getenv("UR_GEMM_LEAN")) : 1;   // development: 0 = general kernel always
static int g_tma_store = getenv("UR_GEMM_TMA_STORE") ? atoi(getenv("UR_GEMM_TMA_STORE")) : 2;
extern "C" int ur_debug_set_gemm_tma_store(int on) {
  const int old = g_tma_store;
  g_tma_store = on;
  return old;
}
static int g_split_mode = getenv("UR_GEMM_SPLITK") ? atoi(getenv("UR_GEMM_SPLITK")) : 1;   // 0: never split K
extern "C" int ur_debug_set_gemm_splitk(int on) {
  const int old = g_split_mode;
  g_split_mode = on;
  return old;
}
static int g_pair_mode = getenv("UR_GEMM_PAIR") ? atoi(getenv("UR_GEMM_PAIR")) : -1;   // -1 auto, 0 never, 1 whenever legal
extern "C" int ur_debug_set_gemm_pair_mode(int mode) {
  const int old = g_pair_mode;
  g_pair_mode = mode;
  return old;
}

static void pick_tile_config(int n, long long m_tiles, int nkb, int fixed_bn, bool pair_legal, bool short_k_linear,
                             int* bn_out, bool* pair_out) {
  const int cands[4] = {256, 160, 128, 64};
  const int sms = num_sms();
  double best_cost = -1.0;
  int best_bn = fixed_bn ? fixed_bn : 64;
  bool best_pair = false;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (fixed_bn && bn != fixed_bn) continue;
    // Round 2 (tools/bench_chain.py, profiles/epilogue_store_experiments_r2.txt): short-K linears whose N is a multiple of
    // 160 run faster on 160-wide tiles than the rounds x k-block-time model predicts for 256 (QKV projections N = 960 /
    // 1920 / 3840: 27.5 vs 30.0, 22.9 vs 28.3, 17.9 vs 18.6 us) -- a 256-wide stage is 48 KB, only three fit, and with
    // 5..20 k-blocks per tile the ring never reaches a steady state
    if (!fixed_bn && short_k_linear && n % 160 == 0 && bn != 160) continue;
    const long long n_tiles = (n + bn - 1) / bn;
    for (int pair = 0; pair < 2; ++pair) {
      if (pair && (!pair_legal || g_pair_mode == 0)) continue;
      if (!pair && pair_legal && g_pair_mode == 1) continue;
      // Measured on B200 (tools/exp_gemm_modes.py, round 1, after the warp-uniform issue path): CTA pairs win on every
      // large-K problem with >= 160-wide N tiles (UNet 3x3 convs at 64^2..16^2: +9..16 %, VAE 256/512-channel convs:
      // +10 %) because they halve the weight traffic L2 -> shared memory; they lose on small-K linears (epilogue
      // bound, the cross-CTA accumulator hand-shake costs more than it saves) and with <= 128-wide tiles.
      // Round 2 (profiles/bench_gemm_r2.txt): the VAE's 128 -> 128 3x3 convolutions at 512^2 (N = one 128-wide tile, the
      // weight tile as large as the activation tile and re-read for each of 16 384 M tiles) gain 9.5 % from pairs as well.
      if (pair && g_pair_mode < 0 && !((bn >= 160 && nkb >= 18) || (bn == 128 && n == 128 && nkb >= 18))) continue;
      double t, rounds;
      if (pair) {
        t = fmax(fmax(2.0 * bn, 256.0 + bn), 300.0);
        rounds = static_cast<double>(((m_tiles + 1) / 2 * n_tiles + sms / 2 - 1) / (sms / 2));
      } else {
        t = fmax(256.0 + 2.0 * bn, 300.0);
        rounds = static_cast<double>((m_tiles * n_tiles + sms - 1) / sms);
      }
      const double cost = rounds * t;
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_bn = bn;
        best_pair = pair != 0;
      }
    }
  }
  *bn_out = best_bn;
  *pair_out = best_pair;
}

// statistics of the finished bf16 output through the stand-alone kernel (tiles spanning images, split-K)
static int stats_fallback(const ur_conv_desc* d, int n_out, cudaStream_t stream) {
  if (d->out_sy != static_cast<int64_t>(d->wout) * d->out_sx && d->hout > 1)
    return set_error(UR_ERR_ARG, "ur_conv_gemm: stats fallback needs a dense pixel pitch");
  return ur_chan_stats(d->out, d->out_sx, d->out_sb, d->batch, d->hout * d->wout, n_out, d->stats, d->stats_ld,
                       d->stats_off, 0, stream);
}

extern "C" int ur_conv_gemm(const ur_conv_desc* d, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!d || !d->x1 || !d->w || !d->out) return set_error(UR_ERR_ARG, "ur_conv_gemm: null pointer");
  const int ctot = d->c1 + d->c2;
  if (d->c1 <= 0 || d->c1 % 8 || d->c2 % 8 || d->ld1 % 8 || (d->c2 && (d->ld2 % 8 || !d->x2)))
    return set_error(UR_ERR_ARG, "ur_conv_gemm: channel counts / pitches must be multiples of 8");
  if (d->c2 > 0 && d->c1 % 64) return set_error(UR_ERR_ARG, "ur_conv_gemm: c1 %% 64 != 0 with two sources");
  if (d->ntaps < 1 || d->ntaps > 9 || (d->stride != 1 && d->stride != 2))
    return set_error(UR_ERR_ARG, "ur_conv_gemm: bad ntaps / stride");
  if (d->n <= 0) return set_error(UR_ERR_ARG, "ur_conv_gemm: n must be positive");
  const bool gated = d->act == UR_ACT_GEGLU || d->act == UR_ACT_GATE;
  if ((reinterpret_cast<uintptr_t>(d->x1) | reinterpret_cast<uintptr_t>(d->x2) | reinterpret_cast<uintptr_t>(d->w)) & 15)
    return set_error(UR_ERR_ARG, "ur_conv_gemm: pointers must be 16-byte aligned");

  // ---- M tile shape: minimise padded work; among equal tile counts take the fewest images per tile (an image's rows
  //      stay together: the epilogue statistics need >= 32 rows of a tile per image, and the TMA boxes are denser),
  //      then the widest tile
  int best_w = 0, best_h = 0, best_bt = 0;
  long long best_cost = -1;
  for (int wl = 7; wl >= 0; --wl) {
    for (int hl = 0; wl + hl <= 7; ++hl) {
      const int Wt = 1 << wl, Ht = 1 << hl, Bt = 128 >> (wl + hl);
      if (d->w_batched && Bt != 1) continue;
      const long long tiles = 1LL * ((d->wout + Wt - 1) / Wt) * ((d->hout + Ht - 1) / Ht) * ((d->batch + Bt - 1) / Bt);
      if (best_cost < 0 || tiles < best_cost || (tiles == best_cost && Bt < best_bt)) {
        best_cost = tiles;
        best_w = wl;
        best_h = hl;
        best_bt = Bt;
      }
    }
  }
  if (best_cost < 0) return set_error(UR_ERR_ARG, "ur_conv_gemm: no tile shape");
  const int Wt = 1 << best_w, Ht = 1 << best_h, Bt = 128 >> (best_w + best_h);

  // ---- split-K candidates: few output tiles and a long K (the UNet / Controller 8x8 level: 4 M tiles, up to 360
  //      k-blocks).  Wide pair tiles keep the per-k-block efficiency; the K range is then cut so that the units fill the
  //      GPU, partial sums meet in the caller's fp32 workspace (red.global.add.v4.f32) and splitk_finish applies the
  //      epilogue.  Needs: plain epilogue (bias / temb row vector / residual), bf16 output, workspace large enough.
  const int nkb_total = d->ntaps * (((d->group_kc ? d->group_kc : ctot) + 63) / 64);
  const long long m_rows = static_cast<long long>(d->batch) * d->hout * d->wout;
  const bool split_ok = g_split_mode != 0 && d->workspace && !(reinterpret_cast<uintptr_t>(d->workspace) & 15) && !gated &&
                        d->act == UR_ACT_NONE && d->alpha == 1.0f && !d->chscale && !d->w_batched && !d->group_kc &&
                        d->out_dtype == UR_DT_BF16 && d->n % 8 == 0 && !d->bn && best_cost <= 8 && nkb_total >= 64 &&
                        2 * 4LL * m_rows * d->n <= d->workspace_bytes;
  int bn = 0;
  bool pair = false;
  pick_tile_config(d->n, best_cost, d->ntaps * (((d->group_kc ? d->group_kc : ctot) + 63) / 64), d->bn ? d->bn : (gated ? ur_conv_gemm_pick_bn(d->n, 1) : 0),
                   !d->w_batched && best_cost >= 2, d->ntaps == 1 && !d->group_kc && (ctot + 63) / 64 <= 20, &bn, &pair);
  if (split_ok) {
    bn = d->n % 160 == 0 ? 160 : 128;
    pair = best_cost >= 2 && g_pair_mode != 0;
  }
  if (bn != 64 && bn != 128 && bn != 160 && bn != 256) return set_error(UR_ERR_ARG, "ur_conv_gemm: bad N tile");
  if (gated && (d->n % bn)) return set_error(UR_ERR_ARG, "ur_conv_gemm: gated act needs n %% bn == 0");
  int kc = ctot;
  if (d->group_kc) {
    if (d->group_kc % 64 || d->group_nc % bn || d->c2)
      return set_error(UR_ERR_ARG, "ur_conv_gemm: grouped conv needs group_kc %% 64 == 0, group_nc %% bn == 0");
    kc = d->group_kc;
  }

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->batch;
  p.Ho = d->hout;
  p.Wo = d->wout;
  p.N = d->n;
  p.cblocks = (kc + 63) / 64;
  p.c1 = d->group_kc ? (1 << 30) : d->c1;
  p.kc = kc;
  p.ntaps = d->ntaps;
  p.stride = d->stride;
  for (int t = 0; t < d->ntaps; ++t) {
    if (d->tap_dy[t] < -8 || d->tap_dy[t] > 7 || d->tap_dx[t] < -8 || d->tap_dx[t] > 7)
      return set_error(UR_ERR_ARG, "ur_conv_gemm: tap offset out of range");
    p.dy_pack |= static_cast<unsigned long long>(d->tap_dy[t] + 8) << (4 * t);
    p.dx_pack |= static_cast<unsigned long long>(d->tap_dx[t] + 8) << (4 * t);
  }
  p.wt_log2 = best_w;
  p.ht_log2 = best_h;
  p.tiles_x = (d->wout + Wt - 1) / Wt;
  p.tiles_y = (d->hout + Ht - 1) / Ht;
  const int tiles_b = (d->batch + Bt - 1) / Bt;
  p.group_kc = d->group_kc;
  p.group_nc = d->group_nc;
  p.w_batched = d->w_batched;
  p.out = d->out;
  p.out_f32 = d->out_dtype == UR_DT_F32;
  p.out_sb = d->out_sb;
  p.out_sy = d->out_sy;
  p.out_sx = d->out_sx;
  p.alpha = d->alpha;
  p.bias = d->bias;
  p.rowvec = d->rowvec;
  p.rowvec_sb = d->rowvec_sb;
  p.chscale = d->chscale;
  p.chscale_sb = d->chscale_sb;
  p.residual = static_cast<const bf16*>(d->residual);
  p.res_sb = d->res_sb;
  p.res_sy = d->res_sy;
  p.res_sx = d->res_sx;
  p.act = d->act;
  p.trace = g_trace;
  p.ksplit = 1;
  p.ws = nullptr;
  p.tma_store = 0;
  p.stats = nullptr;
  p.stats_ld = 0;

  // ---- fast path (persistent kernel): bf16 output with 16-byte aligned pitches
  const int n_out = gated ? d->n / 2 : d->n;
  const bool force_v1 = g_force_v1 != 0;
  const bool out_ok = d->out_dtype == UR_DT_BF16 && !(reinterpret_cast<uintptr_t>(d->out) & 15) && d->out_sx % 8 == 0 &&
                      d->out_sy % 8 == 0 && d->out_sb % 8 == 0 && n_out % 8 == 0;
  const bool res_ok = !d->residual || (!(reinterpret_cast<uintptr_t>(d->residual) & 15) && d->res_sx % 8 == 0 &&
                                       d->res_sy % 8 == 0 && d->res_sb % 8 == 0);
  const bool vec_ok = (!d->rowvec || d->rowvec_sb == 0 || Bt == 1) && (!d->chscale || d->chscale_sb == 0 || Bt == 1);
  const bool fast_path = !force_v1 && out_ok && res_ok && vec_ok && (!gated || bn % 64 == 0);
  const bool pair_path = fast_path && pair;

  // ---- tensor maps
  CUtensorMap mA1, mA2, mW;
  const uint32_t s = static_cast<uint32_t>(d->stride);
  const uint32_t boxA[4] = {64u, static_cast<uint32_t>(Wt) * s, static_cast<uint32_t>(Ht) * s, static_cast<uint32_t>(Bt)};
  const uint32_t estrA[4] = {1u, s, s, 1u};
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->c1), static_cast<uint64_t>(d->win), static_cast<uint64_t>(d->hin),
                              static_cast<uint64_t>(d->batch)};
    const uint64_t str[3] = {static_cast<uint64_t>(d->ld1) * 2, static_cast<uint64_t>(d->ld1) * 2 * d->win,
                             static_cast<uint64_t>(d->ld1) * 2 * d->win * d->hin};
    int rc = encode_tensor_map(&mA1, const_cast<void*>(d->x1), 4, dims, str, boxA, estrA);
    if (rc) return rc;
  }
  if (d->c2 > 0) {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->c2), static_cast<uint64_t>(d->win), static_cast<uint64_t>(d->hin),
                              static_cast<uint64_t>(d->batch)};
    const uint64_t str[3] = {static_cast<uint64_t>(d->ld2) * 2, static_cast<uint64_t>(d->ld2) * 2 * d->win,
                             static_cast<uint64_t>(d->ld2) * 2 * d->win * d->hin};
    int rc = encode_tensor_map(&mA2, const_cast<void*>(d->x2), 4, dims, str, boxA, estrA);
    if (rc) return rc;
  } else {
    mA2 = mA1;
  }
  {
    const uint64_t ktot = static_cast<uint64_t>(d->ntaps) * kc;
    const uint64_t dims[3] = {ktot, static_cast<uint64_t>(d->n), static_cast<uint64_t>(d->w_batched ? d->batch : 1)};
    const uint64_t wld = d->w_ld ? static_cast<uint64_t>(d->w_ld) : ktot;
    const uint64_t wbs = d->w_bs ? static_cast<uint64_t>(d->w_bs) : wld * d->n;
    if (wld % 8 || wbs % 8) return set_error(UR_ERR_ARG, "ur_conv_gemm: weight pitches must be multiples of 8");
    const uint64_t str[2] = {wld * 2, wbs * 2};
    const uint32_t box[3] = {64u, static_cast<uint32_t>(pair_path ? bn / 2 : bn), 1u};
    const uint32_t estr[3] = {1u, 1u, 1u};
    int rc = encode_tensor_map(&mW, const_cast<void*>(d->w), 3, dims, str, box, estr);
    if (rc) return rc;
  }

  if (fast_path) {
    // output tensor map for the TMA-store epilogue: (n_out, Wo, Ho, B) with the caller's strides (channel-slice and
    // sub-pixel phase views included), box (32 ch, Wt, Ht, Bt) = one 128 x 64 B staging buffer, 64-byte swizzle
    CUtensorMap mO = mA1;
    p.tma_store = 0;
    if (g_tma_store) {
      const uint64_t dims[4] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(d->wout), static_cast<uint64_t>(d->hout),
                                static_cast<uint64_t>(d->batch)};
      const uint64_t str[3] = {static_cast<uint64_t>(d->out_sx) * 2, static_cast<uint64_t>(d->out_sy) * 2,
                               static_cast<uint64_t>(d->out_sb) * 2};
      // warp mode (tma_store = 2): a 32-row box = the rows of ONE epilogue warp (32 consecutive tile rows are 32 pixels of
      // a row, or 32 / Wt rows of Wt pixels, inside one image as long as an image contributes >= 32 rows to the tile);
      // otherwise one box per 128-row sub-block issued after a group barrier (tma_store = 1)
      // the lean kernel (bias / residual / statistics only; ur_gemm_persistent.cu) stores through its store warps: 128-row
      // boxes like mode 1, tma_store = 4
      const bool lean = g_lean_epilogue && g_tma_store == 2 && d->act == UR_ACT_NONE && !d->chscale && d->alpha == 1.0f && !split_ok;
      const bool warp_box = !lean && g_tma_store != 1 && Wt * Ht >= 32;
      const uint32_t bwx = static_cast<uint32_t>(Wt < 32 ? Wt : 32);
      const uint32_t box_g[4] = {32u, static_cast<uint32_t>(Wt), static_cast<uint32_t>(Ht), static_cast<uint32_t>(Bt)};
      const uint32_t box_w[4] = {32u, bwx, 32u / bwx, 1u};
      const uint32_t es[4] = {1u, 1u, 1u, 1u};
      const bool str_ok = d->out_sx > 0 && (d->hout == 1 || d->out_sy > 0) && (d->batch == 1 || d->out_sb > 0);
      if (str_ok && encode_tensor_map(&mO, d->out, 4, dims, str, warp_box ? box_w : box_g, es, 64) == UR_OK)
        p.tma_store = lean ? 4 : (warp_box ? 2 : 1);
    }
    // GroupNorm statistics of the output: fused into the epilogue when an M tile never spans two images and K is not
    // split; otherwise a ur_chan_stats pass over the finished output (dense pixel pitch required) does the same
    // (an M tile may hold Bt = 1, 2 or 4 whole images: every 32-row quarter of the tile then lies inside one image)
    const bool stats_fused = d->stats && Bt <= 4;
    if (stats_fused) {
      p.stats = d->stats + 2LL * d->stats_off;
      p.stats_ld = d->stats_ld;
    }
    const int n_tiles = (d->n + bn - 1) / bn;
    const long long m_tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * tiles_b;
    const long long total = static_cast<long long>(n_tiles) * (pair_path ? (m_tiles + 1) / 2 : m_tiles);
    if (total > 0x3fffffffLL) return set_error(UR_ERR_ARG, "ur_conv_gemm: too many tiles");
    p.ksplit = 1;
    p.fd_ntiles = make_fastdiv(static_cast<uint32_t>(n_tiles));
    p.fd_ksplit = make_fastdiv(1);
    p.fd_tx = make_fastdiv(static_cast<uint32_t>(p.tiles_x));
    p.fd_ty = make_fastdiv(static_cast<uint32_t>(p.tiles_y));
    if (split_ok) {
      const long long ctas = pair_path ? 2 * total : total;
      int s = static_cast<int>(num_sms() / (ctas > 0 ? ctas : 1));
      if (s > nkb_total / 16) s = nkb_total / 16;
      if (s > 8) s = 8;
      while (s >= 2 && 4LL * s * m_rows * d->n > d->workspace_bytes) --s;      // one fp32 slab per K slice
      const bool finish_stats_ok = !d->stats || 2ULL * d->n * ((d->n >> 3) >= 256 ? 1 : 256 / (d->n >> 3)) * 4 <= 48 * 1024;
      if (s >= 2 && finish_stats_ok) {
        p.ksplit = s;
        p.fd_ksplit = make_fastdiv(static_cast<uint32_t>(s));
        p.ws = static_cast<float*>(d->workspace);
        p.ws_slab = m_rows * d->n;
        GemmParams pm = p;
        pm.stats = nullptr;                       // the main kernel only writes partial tiles
        int rc = launch_conv_gemm_persistent(pm, mA1, mA2, mW, mO, pair_path, bn, static_cast<int>(total * s), n_tiles, stream);
        if (rc) return rc;
        if (d->stats) {                           // statistics of the finished bf16 output, fused into the finish pass
          p.stats = d->stats + 2LL * d->stats_off;
          p.stats_ld = d->stats_ld;
        }
        return launch_splitk_finish(p, stream);
      }
    }
    int rc = launch_conv_gemm_persistent(p, mA1, mA2, mW, mO, pair_path, bn, static_cast<int>(total), n_tiles, stream);
    if (rc || !d->stats || stats_fused) return rc;
    return stats_fallback(d, n_out, stream);
  }

  if (d->stats && d->out_dtype != UR_DT_BF16) return set_error(UR_ERR_ARG, "ur_conv_gemm: stats needs a bf16 output");
  dim3 grid((d->n + bn - 1) / bn, p.tiles_x * p.tiles_y * tiles_b, 1);
  if (grid.y > 65535) return set_error(UR_ERR_ARG, "ur_conv_gemm: too many M tiles");
  int rc;
  switch (bn) {
    case 64: rc = launch_conv_gemm<64, 4, 2>(p, mA1, mA2, mW, grid, stream); break;
    case 128: rc = launch_conv_gemm<128, 3, 2>(p, mA1, mA2, mW, grid, stream); break;
    case 160: rc = launch_conv_gemm<160, 3, 2>(p, mA1, mA2, mW, grid, stream); break;
    default: rc = launch_conv_gemm<256, 4, 1>(p, mA1, mA2, mW, grid, stream); break;
  }
  if (rc || !d->stats) return rc;
  return stats_fallback(d, n_out, stream);          // non-persistent path: statistics by a pass over the output
}
