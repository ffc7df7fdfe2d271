// Shared between the two implicit-GEMM kernels (ur_gemm.cu: one tile per CTA, direct-store epilogue, any output
// type / alignment; ur_gemm_persistent.cu: persistent, TMEM double-buffered, TMA-store epilogue).
#pragma once
#include "ur_common.cuh"
#include "ur_host.h"

namespace ur {

// Division by a run-time constant without the ~25-instruction IDIV sequence: q = umulhi(x, mul) >> shr (x < 2^31).
struct FastDiv {
  uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  if (d <= 1) {
    f.mul = 0;
    f.shr = 0;
    return f;
  }
  uint32_t lg = 0;
  while ((1u << lg) < d) ++lg;                       // ceil(log2(d))
  const uint32_t p = 31 + lg;
  f.mul = static_cast<uint32_t>(((1ull << p) + d - 1) / d);
  f.shr = p - 32;
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ int fast_div(int x, const FastDiv& f) {
  return f.d <= 1 ? x : static_cast<int>(__umulhi(static_cast<uint32_t>(x), f.mul) >> f.shr);
}
#endif

struct GemmParams {
  int B, Ho, Wo, N;
  int cblocks;         // 64-channel blocks per tap
  int c1;              // channels of source 1 (k-blocks with c >= c1 read source 2)
  int kc;              // K extent per tap in the packed weights
  int ntaps, stride;
  unsigned long long dy_pack, dx_pack;  // 4 bits per tap, value+8
  int wt_log2, ht_log2;                 // M tile = Wt x Ht x Bt pixels, Wt*Ht*Bt = 128
  int tiles_x, tiles_y;
  FastDiv fd_ntiles, fd_ksplit, fd_tx, fd_ty;   // fast division by n_tiles / ksplit / tiles_x / tiles_y (persistent kernel)
  int group_kc, group_nc;
  int w_batched;
  void* out;
  int out_f32;
  long long out_sb, out_sy, out_sx;
  float alpha;
  const float* bias;
  const float* rowvec;
  long long rowvec_sb;
  const float* chscale;
  long long chscale_sb;
  const bf16* residual;
  long long res_sb, res_sy, res_sx;
  int act;
  double* stats;      // persistent kernel: (sum, sumsq) of the bf16 output per (image, channel) accumulated here (GroupNorm
  int stats_ld;       //   statistics fused into the epilogue); channel n of image b at stats[(b * stats_ld + n) * 2]
  int tma_store;      // persistent kernel: 0 = coalesced st.global, 1 = one TMA store per 128-row sub-block, 2 = one per warp,
                      //   4 = lean kernel: one per sub-block issued by the group's store warp
  int ksplit;         // split-K factor (1 = off); work unit u -> (tile u / ksplit, K slice u % ksplit)
  float* ws;          // split-K: fp32 [ksplit][B*Ho*Wo, N] partial-sum slabs (plain stores), epilogue deferred to splitk_finish
  long long ws_slab;  // elements per slab = B*Ho*Wo*N
  long long* trace;   // development: per-role clock64 timestamps of CTA 0 (nullptr = off)
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB

// persistent kernel entry (ur_gemm_persistent.cu).  pair = true: CTA pairs with tcgen05.mma.cta_group::2 on 256 x bn
// tiles (w: tensor map with bn/2-row boxes, total_units = pair tiles); else one CTA per 128 x bn tile.
// split-K epilogue: out = bf16(ws + bias + rowvec + residual) (ur_gemm_persistent.cu)
int launch_splitk_finish(const GemmParams& p, cudaStream_t stream);
int launch_conv_gemm_persistent(const GemmParams& p, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w,
                                const CUtensorMap& mo, bool pair, int bn, int total_units, int n_tiles,
                                cudaStream_t stream);

}  // namespace ur
