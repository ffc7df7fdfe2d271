// ur_attention: fused scaled-dot-product attention (flash-style, online softmax) on tcgen05 / TMEM.
//
//   O[b, q, h*D:(h+1)*D] = softmax(Q K^T / sqrt(D)) V      per (image b, head h), D in {64, 128}
//
// One CTA per (128-query tile, head, image); KV is streamed in tiles of 128 keys:
//   warp 0      TMA producer: Q once, then a 2-stage ring of (K_j, V_j) tiles (128-byte swizzled boxes)
//   warp 1      single-thread tcgen05.mma issuer:  S = Q K_j^T   (128x128 fp32 in TMEM columns [0,128))
//                                                   T = P_j V_j   (128xD   fp32 in TMEM columns [128,128+D))
//   warps 2..9  softmax: TWO THREADS PER QUERY ROW (TMEM lane; 64 score columns each) -> no shuffles.  Two TMEM
//               passes over S (max, then exp2 + bf16 pack), P_j written to shared memory in the K-major SW128 UMMA
//               layout.  O ACCUMULATES IN TMEM across KV tiles (P V issued with accumulate); the softmax reference
//               maximum is only moved - and O / l rescaled in TMEM by exp2(m_ref - m_new) - when the running row
//               maximum has grown by more than 2^8 since the reference was set (lazy rescaling: P <= 256 stays
//               exact in bf16 / fp32; softmax is shift-invariant, so the result is unchanged).
// P V^T uses V exactly as TMA lands it ([keys, D] rows of 128 B) through an MN-major UMMA descriptor.
// With D = 64 a CTA needs 112 KB smem and 256 TMEM columns, so two CTAs share an SM and one CTA's softmax
// overlaps the other's MMAs.
//
// Reference arithmetic: F.scaled_dot_product_attention inside diffusers Attention (BasicTransformerBlock
// attn1/attn2 at base_model.py:138,159,191; Controller AttnDownBlock2D / mid block controller.py:101-141).
#include "ur_common.cuh"
#include "ur_host.h"

namespace ur {

struct AttnParams {
  int Tq, Tk, heads, batch;
  int kv_shared;          // K/V have no batch dimension (constant prompt)
  float scale_log2;       // softmax scale * log2(e)
  bf16* out;
  long long ldo, out_bs;
  long long* trace;       // development: clock64 timestamps of CTA (0,0,0), softmax thread 0 and the MMA thread
};

constexpr int kTileQ = 128;
constexpr int kTileK = 128;

template <int D>
__global__ void __launch_bounds__(320, D == 64 ? 2 : 1)
attention_kernel(const __grid_constant__ AttnParams p, const __grid_constant__ CUtensorMap mapQ,
                 const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV) {
  constexpr int ND = D / 64;                       // 64-column blocks per head
  constexpr int kQBytes = kTileQ * D * 2;
  constexpr int kKBytes = kTileK * D * 2;
  constexpr int kPBytes = kTileQ * kTileK * 2;      // 32 KB: two K-major blocks of [128 x 64]
  constexpr int kStages = 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + kQBytes;                      // stage s: K at s*2*kKBytes, V right after
  uint8_t* sP = sKV + kStages * 2 * kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                     // [kStages]
  uint64_t* kv_empty = kv_full + kStages;           // [kStages]
  uint64_t* s_full = kv_empty + kStages;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_empty + 1;
  uint64_t* o_full = p_full + 1;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);
  int* xch = reinterpret_cast<int*>(tmem_slot + 2);       // [128] running row max (order-preserving int key) shared by the two half-row threads

  const int warp = warp_uniform(static_cast<int>(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTileQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nkv = (p.Tk + kTileK - 1) / kTileK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 256);
    mbar_init(p_full, 256);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_T = tmem_base + 128;

  if (warp == 0) {
    // =============================== TMA producer (whole warp; the elected lane issues) ===============================
    if (elect_one_sync()) {
      mbar_expect_tx(q_full, kQBytes);
#pragma unroll
      for (int nb = 0; nb < ND; ++nb) tma_load_3d(sQ + nb * (kTileQ * 128), &mapQ, q_full, h * D + nb * 64, q0, b);
    }
    __syncwarp();
    const int bk = p.kv_shared ? 0 : b;
    for (int j = 0; j < nkv; ++j) {
      const int s = j % kStages;
      const uint32_t ph = (j / kStages) & 1;
      mbar_wait(&kv_empty[s], ph ^ 1);
      uint8_t* sk = sKV + s * 2 * kKBytes;
      if (elect_one_sync()) {
        mbar_expect_tx(&kv_full[s], 2 * kKBytes);
#pragma unroll
        for (int nb = 0; nb < ND; ++nb) {
          tma_load_3d(sk + nb * (kTileK * 128), &mapK, &kv_full[s], h * D + nb * 64, j * kTileK, bk);
          tma_load_3d(sk + kKBytes + nb * (kTileK * 128), &mapV, &kv_full[s], h * D + nb * 64, j * kTileK, bk);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp; the elected lane issues) ===============================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 1);     // B (= V) is MN-major
    const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP), aKV = smem_u32(sKV);
    mbar_wait(q_full, 0);
    // S_j = Q K_j^T into TMEM columns [0,128); K_j lives in stage j % kStages
    auto issue_s = [&](int j) {
      const int s = j % kStages;
      mbar_wait(&kv_full[s], (j / kStages) & 1);
      tc_fence_after();
      const uint32_t aK = aKV + s * 2 * kKBytes;
      if (elect_one_sync()) {
#pragma unroll
        for (int nb = 0; nb < ND; ++nb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc_mma_bf16(tmem_S, umma_desc_k_sw128(aQ + nb * (kTileQ * 128)) + 2 * k,
                        umma_desc_k_sw128(aK + nb * (kTileK * 128)) + 2 * k, idesc_s, (nb | k) != 0 ? 1u : 0u);
          }
        }
        tc_commit(s_full);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      const int s = j % kStages;
      const uint32_t aV = aKV + s * 2 * kKBytes + kKBytes;
      const bool trm = p.trace && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j >= 4 && j < 8;
      if (trm) p.trace[32 + (j - 4) * 8 + 0] = clock64();
      // S_{j+1} goes first (as soon as S_j has been read) so that the softmax warps find it ready when they finish
      // tile j; P_j V_j follows once P_j is in shared memory
      mbar_wait(s_empty, j & 1);          // arrives as soon as the softmax threads hold the last S_j chunk in registers
      if (trm) p.trace[32 + (j - 4) * 8 + 1] = clock64();
      if (j + 1 < nkv) issue_s(j + 1);    // overlaps the second half of the exp pass of tile j
      if (trm) p.trace[32 + (j - 4) * 8 + 2] = clock64();
      mbar_wait(p_full, j & 1);
      // ---- O += P_j V_j (any rescaling of O for tile j happened before the p_full arrivals)
      if (trm) p.trace[32 + (j - 4) * 8 + 3] = clock64();
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int nb = 0; nb < ND; ++nb) {
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k) {
            // A = P: K-major, keys [16k, 16k+16) live in block k/4 (16 KB each), 32 B per 16 keys inside the atom
            const uint64_t da = umma_desc_k_sw128(aP + (k >> 2) * (kTileQ * 128)) + 2 * (k & 3);
            // B = V: MN-major, 16 keys = 16 rows of 128 B
            const uint64_t db = umma_desc_mn_sw128(aV + nb * (kTileK * 128) + k * 16 * 128, kTileK * 128);
            tc_mma_bf16(tmem_T + nb * 64, da, db, idesc_pv, (j | k) != 0 ? 1u : 0u);
          }
        }
        tc_commit(o_full);
        tc_commit(&kv_empty[s]);
      }
      __syncwarp();
      if (trm) p.trace[32 + (j - 4) * 8 + 4] = clock64();
    }
  } else {
    // =============================== softmax / output: TWO threads per query row ===============================
    // Warps 2..9: warp w and w+4 share a TMEM lane quarter; thread (row r, half hf) owns score columns
    // [64 hf, 64 hf + 64) of its row and output columns [D/2 hf, D/2 hf + D/2).  The row maximum is exchanged through
    // shared memory once per KV tile; the row sum is kept as two partial sums and combined at the end.
    constexpr int DH = D / 2;
    const int qd = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tS = tmem_S + lane_off + hf * 64;
    const uint32_t tT = tmem_T + lane_off + hf * DH;
    float m = -INFINITY, l = 0.f;             // m: reference maximum (log2 domain) the exponentials are taken against
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    uint8_t* prow = sP + hf * (kTileQ * 128) + r * 128;        // key block hf (64 keys) of the P tile, row r
    int* my_x = xch + r;
    // order-preserving float <-> int key so that one shared atomicMax keeps the RUNNING row maximum of both halves
    auto f2key = [](float f) { const int b = __float_as_int(f); return b >= 0 ? b : (b ^ 0x7fffffff); };
    auto key2f = [](int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); };
    if (hf == 0) *my_x = f2key(-INFINITY);
    asm volatile("bar.sync 1, 256;" ::: "memory");

    for (int j = 0; j < nkv; ++j) {
      const bool trs = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && r == 0 && hf == 0 && j >= 4 && j < 8;
      if (trs) p.trace[(j - 4) * 8 + 0] = clock64();
      mbar_wait(s_full, j & 1);
      if (trs) p.trace[(j - 4) * 8 + 1] = clock64();
      tc_fence_after();
      const int kvalid = p.Tk - j * kTileK - hf * 64;   // keys of this thread's 64-column half that exist
      const bool full_tile = kvalid >= 64;               // warp-uniform: only the last tile needs key masking
      // ---- pass 1: partial row max of the raw scores (scale > 0 is applied afterwards)
      float raw = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + c * 32, v);
        tmem_ld_wait();
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 32; ++i) raw = fmaxf(raw, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < kvalid) raw = fmaxf(raw, __uint_as_float(v[i]));
        }
      }
      atomicMax(my_x, f2key(raw));
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float mx = key2f(*my_x) * p.scale_log2;    // running max over all tiles so far, log2 domain
      // ---- lazy rescaling: move the reference only when the running max outgrew it by more than 2^8 (always on
      //      the first tile).  tcgen05.ld/st are warp-collective, so the warp goes through the O update when ANY of
      //      its rows needs it; rows that do not get alpha = 1.  Both threads of a row see the same history of
      //      *my_x, hence take identical decisions.
      const bool need = (j == 0) || (mx - m > 8.0f);
      if (trs) p.trace[(j - 4) * 8 + 2] = clock64();
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? exp2_approx(m - mx) : 1.0f;       // j == 0: exp2(-inf) = 0, but O is not read then
        if (j > 0) {
          mbar_wait(o_full, (j - 1) & 1);                             // P_{j-1} V_{j-1} has landed in TMEM
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < DH / 32; ++c) {
            uint32_t t0[32];
            tmem_ld32(tT + c * 32, t0);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t0[i] = __float_as_uint(__uint_as_float(t0[i]) * alpha);
            tmem_st32(tT + c * 32, t0);
          }
          tmem_st_wait();
          l *= alpha;
        }
        if (need) m = mx;
      } else if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);      // P_{j-1} V_{j-1} has consumed the P buffer this tile overwrites
      }
      if (trs) p.trace[(j - 4) * 8 + 4] = clock64();
      // ---- pass 2: P = exp2(s * scale - max) -> bf16 -> shared memory (K-major SW128), partial row sum
      const float nmx = -m;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + c * 32, v);
        tmem_ld_wait();
        if (c == 1) {                     // S_j is fully consumed: the MMA warp may overwrite it with S_{j+1}
          tc_fence_before();
          mbar_arrive(s_empty);
        }
        uint32_t pk[16];
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = exp2_approx(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, nmx));
            const float p1 = exp2_approx(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, nmx));
            l0 += p0;
            l1 += p1;
            pk[i] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k0 = c * 32 + 2 * i;
            const float p0 = (k0 < kvalid) ? exp2_approx(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, nmx)) : 0.f;
            const float p1 = (k0 + 1 < kvalid) ? exp2_approx(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, nmx)) : 0.f;
            l0 += p0;
            l1 += p1;
            pk[i] = pack_bf16(p0, p1);
          }
        }
        // 32 keys = 4 chunks of 16 B at chunk index c * 4 + q of the 128-byte row (128-byte swizzle)
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const uint32_t chunk = static_cast<uint32_t>(c * 4 + qq) ^ swz;
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
        }
      }
      l += l0 + l1;
      if (trs) p.trace[(j - 4) * 8 + 5] = clock64();
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      if (trs) p.trace[(j - 4) * 8 + 6] = clock64();
    }
    // ---- last tile's P V, combine the two partial row sums, normalise, store
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    float O[DH];
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t t0[32];
      tmem_ld32(tT + c * 32, t0);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) O[c * 32 + i] = __uint_as_float(t0[i]);
    }
    // combine the two partial row sums through the same slot (two rounds: hf 0 publishes, hf 1 adds and publishes)
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (hf == 0) *my_x = __float_as_int(l);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (hf == 1) {
      l += __int_as_float(*my_x);
      *my_x = __float_as_int(l);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (hf == 0) l = __int_as_float(*my_x);
    const int q = q0 + r;
    if (q < p.Tq) {
      const float inv = 1.f / l;
      bf16* op = p.out + b * p.out_bs + static_cast<long long>(q) * p.ldo + h * D + hf * DH;
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        uint4 o;
        o.x = pack_bf16(O[8 * i] * inv, O[8 * i + 1] * inv);
        o.y = pack_bf16(O[8 * i + 2] * inv, O[8 * i + 3] * inv);
        o.z = pack_bf16(O[8 * i + 4] * inv, O[8 * i + 5] * inv);
        o.w = pack_bf16(O[8 * i + 6] * inv, O[8 * i + 7] * inv);
        reinterpret_cast<uint4*>(op)[i] = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// =====================================================================================================================
// attention2_kernel: second-generation flash attention (default).
//
//   * ONE softmax thread per query row (TMEM lane) and NQ = 2 query tiles of 128 rows per CTA (D = 64): the two softmax
//     warp groups ping-pong -- while group A runs the softmax of KV tile j, the tensor core computes S_B / P_B V of
//     group B -- and share every K/V tile (half the K/V traffic of one-tile CTAs).  A thread owns its whole score row:
//     no cross-thread maximum exchange, no named barriers, no atomics in the loop.
//   * ONE TMEM pass over S per KV tile: the 128 scores of the row are loaded once (4 x tcgen05.ld.x32, 128 registers);
//     s_empty is released as soon as they are in registers, so S(j+1) = Q K_{j+1}^T runs under the whole softmax of
//     tile j.  Row maximum by 3-input max (FMNMX3: 64 instead of 128 ALU operations).
//   * The softmax arithmetic is issue- and MUFU-bound (16 384 exponentials per 128 x 128 tile = 1 024 MUFU cycles per
//     SM against 512 tensor cycles): scale / subtract and the row sum use packed fma.rn.f32x2 / add.rn.f32x2 (half the
//     FMA-pipe instructions) and kPolyOf8 of every 8 element pairs take a degree-3 Cody-Waite polynomial exp2 on the
//     FMA / ALU pipes instead of MUFU.EX2 (max relative error 7.5e-5, far below the bf16 rounding of P).
//   * P never touches shared memory: the softmax threads write bf16 P straight into tensor memory (tcgen05.st) and
//     P V runs with its A operand read from TMEM (tcgen05.mma [d], [a_tmem], b_desc).  With P in shared memory the kernel
//     was bound by shared-memory bandwidth (measured slower than attention_kernel, gpurun_out/r2c2_bench_attn.log):
//     32 KB of P stores + 32 KB of UMMA reads per 128 x 128 tile on top of 64 KB of Q / K / V operand reads at
//     128 B/clk/SM.  TMEM: S 2 x 128 + O 2 x 64 + P 2 x 64 = 512 columns.
//   * O accumulates in TMEM with the same exact lazy rescaling as attention_kernel (reference maximum only moves when
//     the running maximum outgrew it by more than 2^8); the rescale touches the thread's own row only.
//   warp 0 TMA (Q tiles once, 3-stage K/V ring), warp 1 tcgen05 issuer, warps 4.. softmax (one warp group per query
//   tile).  The softmax threads hold a whole score row (128 registers): warp group 0 gives its registers away
//   (setmaxnreg.dec 56) and the softmax warp groups grow to 224 (setmaxnreg.inc).
// D = 128 (Controller low-resolution levels) runs with NQ = 1 and a 2-stage ring (shared-memory budget).
// =====================================================================================================================
constexpr int kPolyOf8Default = 3;      // element pairs (of every 8) whose exp2 runs on the FMA pipe (template parameter POLY)

__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\n"
      "mov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %4};\nmov.b64 rc, {%5, %5};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\n"
      "mov.b64 {%0, %1}, rd;\n}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void fma2v(float& d0, float& d1, float a0, float a1, float b0, float b1, float c) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\n"
      "mov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %6};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\n"
      "mov.b64 {%0, %1}, rd;\n}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n.reg .b64 ra, rb, rd;\n"
      "mov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "add.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// 2^x for a pair on the FMA / ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax
// polynomial (relative error <= 7.5e-5); 2^n by adding n to the exponent field.  x is clamped to >= -125 so the
// exponent arithmetic cannot wrap (the result is then < 2^-124: zero for every purpose here).
__device__ __forceinline__ void exp2_poly2(float& y0, float& y1, float x0, float x1) {
  constexpr float kMagic = 12582912.0f;     // 1.5 * 2^23: adding it leaves round(x) in the low mantissa bits
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  float t0, t1, n0, n1, f0, f1, p0, p1;
  add2(t0, t1, x0, x1, kMagic, kMagic);
  add2(n0, n1, t0, t1, -kMagic, -kMagic);
  add2(f0, f1, x0, x1, -n0, -n1);
  fma2(p0, p1, f0, f1, 0.05517156941252964f, 0.24261111474211297f);
  fma2v(p0, p1, p0, p1, f0, f1, 0.6932610038691536f);
  fma2v(p0, p1, p0, p1, f0, f1, 0.9999280744577427f);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// One softmax thread's whole KV loop (attention2_kernel and attention512_kernel): tS / tO / tP = TMEM addresses of this
// thread's S row (128 fp32 columns), O row (OW fp32 columns) and P row (64 columns holding 128 bf16: P is the A operand
// of the P V tensor-core product and is read from tensor memory, so it never touches shared memory -- with P in shared
// memory the kernel was bound by shared-memory bandwidth: 32 KB of P stores + 32 KB of UMMA reads per 128 x 128 tile on
// top of the Q / K / V operand reads).
template <int OW, int POLY>
__device__ __forceinline__ void softmax_rows(uint32_t tS, uint32_t tO, uint32_t tP, uint64_t* s_full,
                                             uint64_t* s_empty, uint64_t* p_full, uint64_t* o_full, int nkv, int Tk,
                                             float sc, bf16* op, bool row_valid) {
  float m = -INFINITY, mrun = -INFINITY, l0 = 0.f, l1 = 0.f;
  for (int j = 0; j < nkv; ++j) {
    const int kvalid = Tk - j * kTileK;
    const int nchunk = kvalid >= kTileK ? 4 : ((kvalid + 31) >> 5);
    mbar_wait(s_full, j & 1);
    tc_fence_after();
    uint32_t v[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < nchunk) tmem_ld32(tS + c * 32, v[c]);
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(s_empty);
    if (kvalid < kTileK) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i >= kvalid && c < nchunk) v[c][i] = __float_as_uint(-INFINITY);
    }
    float mx0 = mrun, mx1 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < nchunk) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = max3(mx0, __uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1]));
          mx1 = max3(mx1, __uint_as_float(v[c][i + 2]), __uint_as_float(v[c][i + 3]));
        }
      }
    }
    mrun = fmaxf(mx0, mx1);
    const float mxl = mrun * sc;
    const bool need = (j == 0) || (mxl - m > 8.0f);
    if (__any_sync(0xffffffffu, need)) {
      const float alpha = need ? exp2_approx(m - mxl) : 1.0f;
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < OW / 32; ++c) {
          uint32_t t0[32];
          tmem_ld32(tO + c * 32, t0);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) t0[i] = __float_as_uint(__uint_as_float(t0[i]) * alpha);
          tmem_st32(tO + c * 32, t0);
        }
        tmem_st_wait();
        l0 *= alpha;
        l1 *= alpha;
      }
      if (need) m = mxl;
    } else if (j > 0) {
      mbar_wait(o_full, (j - 1) & 1);
    }
    const float nm = -m;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < nchunk) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0, a1, e0, e1;
          fma2(a0, a1, __uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1]), sc, nm);
          if ((i & 7) < POLY) {
            exp2_poly2(e0, e1, a0, a1);
          } else {
            e0 = exp2_approx(a0);
            e1 = exp2_approx(a1);
          }
          add2(l0, l1, l0, l1, e0, e1);
          pk[i] = pack_bf16(e0, e1);
        }
        tmem_st16(tP + c * 16, pk);          // keys [32 c, 32 c + 32) of this row: 16 packed bf16 pairs
      }
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(p_full);
  }
  mbar_wait(o_full, (nkv - 1) & 1);
  tc_fence_after();
  const float inv = 1.f / (l0 + l1);
#pragma unroll
  for (int c = 0; c < OW / 32; ++c) {
    uint32_t t0[32];
    tmem_ld32(tO + c * 32, t0);
    tmem_ld_wait();
    if (row_valid) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(t0[8 * i]) * inv, __uint_as_float(t0[8 * i + 1]) * inv);
        o.y = pack_bf16(__uint_as_float(t0[8 * i + 2]) * inv, __uint_as_float(t0[8 * i + 3]) * inv);
        o.z = pack_bf16(__uint_as_float(t0[8 * i + 4]) * inv, __uint_as_float(t0[8 * i + 5]) * inv);
        o.w = pack_bf16(__uint_as_float(t0[8 * i + 6]) * inv, __uint_as_float(t0[8 * i + 7]) * inv);
        reinterpret_cast<uint4*>(op + c * 32)[i] = o;
      }
    }
  }
}

template <int D, int NQ>
struct Attn2Cfg {
  static constexpr int kStages = D == 64 ? 4 : 3;
  static constexpr int kQBytes = kTileQ * D * 2;
  static constexpr int kKBytes = kTileK * D * 2;
  static constexpr int kSmem = NQ * kQBytes + kStages * 2 * kKBytes + 1024;
  static constexpr int kThreads = (4 + 4 * NQ) * 32;      // warp group 0: TMA, MMA, 2 idle warps; then 4 warps per query tile
  // TMEM columns: S [0, 128 NQ), O [128 NQ, 128 NQ + D NQ), P (bf16 pairs) [128 NQ + D NQ, + 64 NQ)
  static constexpr int kColO = NQ * 128, kColP = NQ * 128 + NQ * D;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kColP + NQ * 64 <= 512, "tensor memory budget");
};

template <int D, int NQ, int POLY>
__global__ void __launch_bounds__(Attn2Cfg<D, NQ>::kThreads, 1)
attention2_kernel(const __grid_constant__ AttnParams p, const __grid_constant__ CUtensorMap mapQ,
                  const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV) {
  using Cfg = Attn2Cfg<D, NQ>;
  constexpr int ND = D / 64;
  constexpr int kStages = Cfg::kStages;
  constexpr int kQBytes = Cfg::kQBytes, kKBytes = Cfg::kKBytes;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                   // [NQ] x [ND blocks of 128 rows x 128 B]
  uint8_t* sKV = sQ + NQ * kQBytes;                     // stage s: K at s * 2 * kKBytes, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kStages * 2 * kKBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = q_full + 1;                       // [kStages]
  uint64_t* kv_empty = kv_full + kStages;               // [kStages]
  uint64_t* s_full = kv_empty + kStages;                // [NQ]
  uint64_t* s_empty = s_full + NQ;                      // [NQ]
  uint64_t* p_full = s_empty + NQ;                      // [NQ]
  uint64_t* o_full = p_full + NQ;                       // [NQ]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + NQ);

  const int warp = warp_uniform(static_cast<int>(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const int q_base = blockIdx.x * (kTileQ * NQ);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nkv = (p.Tk + kTileK - 1) / kTileK;
  const int nq_act = (NQ == 2 && q_base + kTileQ < p.Tq) ? 2 : 1;     // query tiles of this CTA that hold rows

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int x = 0; x < NQ; ++x) {
      mbar_init(&s_full[x], 1);
      mbar_init(&s_empty[x], 128);
      mbar_init(&p_full[x], 128);
      mbar_init(&o_full[x], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);
  pdl_launch_dependents();
  pdl_wait();
  // (each setmaxnreg sits inside its role's branch: after a control-flow merge ptxas would apply the smaller budget
  //  to every role)
  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one_sync()) {
      mbar_expect_tx(q_full, nq_act * kQBytes);
      for (int x = 0; x < nq_act; ++x)
#pragma unroll
        for (int nb = 0; nb < ND; ++nb)
          tma_load_3d(sQ + x * kQBytes + nb * (kTileQ * 128), &mapQ, q_full, h * D + nb * 64, q_base + x * kTileQ, b);
    }
    __syncwarp();
    const int bk = p.kv_shared ? 0 : b;
    for (int j = 0; j < nkv; ++j) {
      const int s = j % kStages;
      mbar_wait(&kv_empty[s], ((j / kStages) & 1) ^ 1);
      uint8_t* sk = sKV + s * 2 * kKBytes;
      if (elect_one_sync()) {
        mbar_expect_tx(&kv_full[s], 2 * kKBytes);
#pragma unroll
        for (int nb = 0; nb < ND; ++nb) {
          tma_load_3d(sk + nb * (kTileK * 128), &mapK, &kv_full[s], h * D + nb * 64, j * kTileK, bk);
          tma_load_3d(sk + kKBytes + nb * (kTileK * 128), &mapV, &kv_full[s], h * D + nb * 64, j * kTileK, bk);
        }
      }
      __syncwarp();
    }
   } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 1);     // B (= V) is MN-major
    const uint32_t aQ = smem_u32(sQ), aKV = smem_u32(sKV);
    mbar_wait(q_full, 0);
    auto issue_s = [&](int x, int j) {          // S_x(j) = Q_x K_j^T -> TMEM columns [128 x, 128 x + 128)
      const int s = j % kStages;
      if (x == 0) mbar_wait(&kv_full[s], (j / kStages) & 1);       // (x = 1 follows x = 0 of the same j)
      tc_fence_after();
      const uint32_t aK = aKV + s * 2 * kKBytes;
      if (elect_one_sync()) {
#pragma unroll
        for (int nb = 0; nb < ND; ++nb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_base + x * 128, umma_desc_k_sw128(aQ + x * kQBytes + nb * (kTileQ * 128)) + 2 * k,
                        umma_desc_k_sw128(aK + nb * (kTileK * 128)) + 2 * k, idesc_s, (nb | k) != 0 ? 1u : 0u);
        tc_commit(&s_full[x]);
      }
      __syncwarp();
    };
    for (int x = 0; x < nq_act; ++x) issue_s(x, 0);
    for (int j = 0; j < nkv; ++j) {
      const int s = j % kStages;
      const uint32_t aV = aKV + s * 2 * kKBytes + kKBytes;
      // k-steps of 16 keys that hold valid keys (the last KV tile may be partial; P columns past them are not written)
      const int kvalid = min(kTileK, p.Tk - j * kTileK);
      const int nk16 = ((kvalid + 31) >> 5) << 1;
      for (int x = 0; x < nq_act; ++x) {
        if (j + 1 < nkv) {
          mbar_wait(&s_empty[x], j & 1);        // group x holds S_x(j) in registers
          issue_s(x, j + 1);                    // runs under the softmax of tile j
        }
        mbar_wait(&p_full[x], j & 1);           // P_x(j) in tensor memory, O_x rescaled if it had to be
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int nb = 0; nb < ND; ++nb) {
            for (int k = 0; k < nk16; ++k) {     // A = P from TMEM: 16 keys = 8 columns of bf16 pairs per k-step
              const uint64_t db = umma_desc_mn_sw128(aV + nb * (kTileK * 128) + k * 16 * 128, kTileK * 128);
              tc_mma_bf16_ts(tmem_base + Cfg::kColO + x * D + nb * 64, tmem_base + Cfg::kColP + x * 64 + 8 * k, db,
                             idesc_pv, (j | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&o_full[x]);
          if (x == nq_act - 1) tc_commit(&kv_empty[s]);
        }
        __syncwarp();
      }
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // =============================== softmax / output: one thread per query row ===============================
    const int x = (warp - 4) >> 2;                       // query tile (softmax group)
    const int qd = warp & 3;                             // TMEM lane quarter this warp may access
    const int r = qd * 32 + lane;
    if (x < nq_act) {
      const uint32_t tl = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
      const int q = q_base + x * kTileQ + r;
      softmax_rows<D, POLY>(tl + x * 128, tl + Cfg::kColO + x * D, tl + Cfg::kColP + x * 64, &s_full[x], &s_empty[x], &p_full[x],
                      &o_full[x], nkv, p.Tk, p.scale_log2,
                      p.out + b * p.out_bs + static_cast<long long>(q) * p.ldo + h * D, q < p.Tq);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int D, int NQ, int POLY>
static int launch_attention2(const AttnParams& p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                             cudaStream_t stream) {
  using Cfg = Attn2Cfg<D, NQ>;
  static_assert(Cfg::kSmem <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention2_kernel<D, NQ, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attention2)");
    configured = true;
  }
  dim3 grid((p.Tq + kTileQ * NQ - 1) / (kTileQ * NQ), p.heads, p.batch);
  cudaError_t e = launch_kernel(attention2_kernel<D, NQ, POLY>, grid, dim3(Cfg::kThreads), Cfg::kSmem, stream, p, mq, mk, mv);
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "attention2 launch");
}

// =====================================================================================================================
// attention512_kernel: flash attention for ONE 512-wide head (the VAE mid-block Attention of autoencoder.py:32,44:
// diffusers Attention(512, heads=1, dim_head=512) over all 4 096 / 16 384 latent pixels).  Replaces the GEMM -> fp32
// scores -> softmax -> GEMM path: no score tensor in HBM (4 096^2 fp32 = 64 MB per image, 16 384^2 = 1 GB).
//
//   O[128 x 512] fp32 would need all 512 TMEM columns, leaving none for S: the output columns are split in two halves
//   and a CTA owns (128-query tile, column half, image); both halves compute S = Q K^T (1.5x the minimal FLOPs, the
//   price of keeping S and O on chip).  TMEM: S 128 + O 256 + P 64 columns (P is the TMEM A operand of P V).
//   Q (128 x 512, 128 KB) stays resident in shared memory; K_j streams through a 6-stage ring as eight 64-channel
//   chunks (S accumulates over them), V_j as four 64-column chunks of this CTA's half, all 16 KB boxes:
//     warp 0 TMA producer   K(0); then per KV tile j:  K(j+1) chunks, V(j) chunks      (the MMA warp's consumption order)
//     warp 1 tcgen05 issuer S(0); then per j: [s_empty(j)] S(j+1);  [p_full(j)] O += P(j) V(j)
//     warps 4..7 softmax, one thread per query row -- same single-pass arithmetic as attention2_kernel.
//   The main loop is tensor-bound (3 072 tensor cycles per KV tile against ~700 softmax cycles).
// =====================================================================================================================
constexpr int kD5 = 512, kOW5 = 256, kRing5 = 6, kChunk5 = kTileK * 64 * 2;     // 16 KB ring stages

__global__ void __launch_bounds__(256, 1)
attention512_kernel(const __grid_constant__ AttnParams p, const __grid_constant__ CUtensorMap mapQ,
                    const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV) {
  constexpr int kQBytes = kTileQ * kD5 * 2;             // 128 KB: 8 blocks of [128 rows x 128 B]
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sR = sQ + kQBytes;                           // ring: kRing5 x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + kRing5 * kChunk5);
  uint64_t* q_full = bars;
  uint64_t* r_full = q_full + 1;                        // [kRing5]
  uint64_t* r_empty = r_full + kRing5;                  // [kRing5]
  uint64_t* s_full = r_empty + kRing5;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_empty + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = warp_uniform(static_cast<int>(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTileQ;
  const int h = blockIdx.y >> 1, half = blockIdx.y & 1;
  const int b = blockIdx.z;
  const int nkv = (p.Tk + kTileK - 1) / kTileK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kRing5; ++s) {
      mbar_init(&r_full[s], 1);
      mbar_init(&r_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);
  pdl_launch_dependents();
  pdl_wait();

  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one_sync()) {
      mbar_expect_tx(q_full, kQBytes);
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) tma_load_3d(sQ + nb * (kTileQ * 128), &mapQ, q_full, h * kD5 + nb * 64, q0, b);
    }
    __syncwarp();
    const int bk = p.kv_shared ? 0 : b;
    int it = 0;
    auto load = [&](const CUtensorMap* m, int ch, int j) {
      const int s = it % kRing5;
      mbar_wait(&r_empty[s], ((it / kRing5) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&r_full[s], kChunk5);
        tma_load_3d(sR + s * kChunk5, m, &r_full[s], ch, j * kTileK, bk);
      }
      __syncwarp();
      ++it;
    };
    for (int c = 0; c < 8; ++c) load(&mapK, h * kD5 + c * 64, 0);
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv)
        for (int c = 0; c < 8; ++c) load(&mapK, h * kD5 + c * 64, j + 1);
      for (int c = 0; c < 4; ++c) load(&mapV, h * kD5 + half * kOW5 + c * 64, j);
    }
   } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 1);     // B (= V chunk) is MN-major
    const uint32_t aQ = smem_u32(sQ), aR = smem_u32(sR);
    const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 128 + kOW5;
    int it = 0;
    mbar_wait(q_full, 0);
    auto issue_s = [&]() {                    // S = sum over eight 64-channel chunks of Q_c K_c^T
      for (int c = 0; c < 8; ++c, ++it) {
        const int s = it % kRing5;
        mbar_wait(&r_full[s], (it / kRing5) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tS, umma_desc_k_sw128(aQ + c * (kTileQ * 128)) + 2 * k, umma_desc_k_sw128(aR + s * kChunk5) + 2 * k,
                        idesc_s, (c | k) != 0 ? 1u : 0u);
          tc_commit(&r_empty[s]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) tc_commit(s_full);
      __syncwarp();
    };
    issue_s();
    for (int j = 0; j < nkv; ++j) {
      const int kvalid = min(kTileK, p.Tk - j * kTileK);
      const int nk16 = ((kvalid + 31) >> 5) << 1;
      mbar_wait(s_empty, j & 1);              // the softmax threads hold S(j) in registers
      if (j + 1 < nkv) issue_s();             // S(j+1) runs under the softmax of tile j
      mbar_wait(p_full, j & 1);
      for (int c = 0; c < 4; ++c, ++it) {     // O[:, 64 c .. 64 c + 64) += P(j) V_c(j)
        const int s = it % kRing5;
        mbar_wait(&r_full[s], (it / kRing5) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          for (int k = 0; k < nk16; ++k) {
            const uint64_t db = umma_desc_mn_sw128(aR + s * kChunk5 + k * 16 * 128, kTileK * 128);
            tc_mma_bf16_ts(tO + c * 64, tP + 8 * k, db, idesc_pv, (j | k) != 0 ? 1u : 0u);
          }
          tc_commit(&r_empty[s]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) tc_commit(o_full);
      __syncwarp();
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // =============================== softmax / output: one thread per query row ===============================
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int q = q0 + r;
    softmax_rows<kOW5, kPolyOf8Default>(tmem_base + lane_off, tmem_base + lane_off + 128, tmem_base + lane_off + 128 + kOW5,
                       s_full, s_empty, p_full, o_full, nkv, p.Tk, p.scale_log2,
                       p.out + b * p.out_bs + static_cast<long long>(q) * p.ldo + h * kD5 + half * kOW5, q < p.Tq);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int launch_attention512(const AttnParams& p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                               cudaStream_t stream) {
  constexpr int smem = kTileQ * kD5 * 2 + kRing5 * kChunk5 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attention512)");
    configured = true;
  }
  dim3 grid((p.Tq + kTileQ - 1) / kTileQ, 2 * p.heads, p.batch);
  cudaError_t e = launch_kernel(attention512_kernel, grid, dim3(256), smem, stream, p, mq, mk, mv);
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "attention512 launch");
}

// =====================================================================================================================
// attention_kv1_kernel: head_dim 64, ALL keys in one 128-key tile (cross-attention on the 77-token prompt:
// base_model.py:159 attn2 -- 300 launches per forward).  The general kernel spends one CTA per 128 queries, and with a
// single KV tile a CTA is nothing but prologue + three dependent latencies (26 us for 4096 x 77, B = 8, 5 heads).  Here
// a CTA keeps K / V resident and LOOPS over `nq` consecutive query tiles of its (image, head): Q tiles stream through a
// 2-stage ring, S_{i+1} = Q_{i+1} K^T runs on the tensor core while the softmax threads finish tile i, and there is no
// running maximum / rescaling at all.  One softmax thread per query row (4 warps), two TMEM passes over its <= 128
// scores; the thread that wrote P_i also reads O_i = P_i V back, scales by 1 / l and stores its 128-byte output row,
// so S, O and the P tile need no double buffering (a thread starts tile i+1 only after it has seen O_i).
// smem: K 16 KB + V 16 KB + Q 2 x 16 KB + P 32 KB; TMEM: S 128 + O 64 columns -> two CTAs per SM.
__global__ void __launch_bounds__(192, 2)
attention_kv1_kernel(const __grid_constant__ AttnParams p, const __grid_constant__ CUtensorMap mapQ,
                     const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV, int nq) {
  constexpr int D = 64;
  constexpr int kQBytes = kTileQ * D * 2, kKBytes = kTileK * D * 2, kPBytes = kTileQ * kTileK * 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + kKBytes;
  uint8_t* sQ = sV + kKBytes;                       // [2] stages
  uint8_t* sP = sQ + 2 * kQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;                      // [2]
  uint64_t* q_empty = bars + 3;                     // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = warp_uniform(static_cast<int>(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int tiles = (p.Tq + kTileQ - 1) / kTileQ;
  const int t0 = blockIdx.x * nq;                                  // first query tile of this CTA
  const int n_my = min(nq, tiles - t0);                            // (host: t0 < tiles for every block)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    const int bk = p.kv_shared ? 0 : b;
    if (elect_one_sync()) {
      mbar_expect_tx(kv_full, 2 * kKBytes);
      tma_load_3d(sK, &mapK, kv_full, h * D, 0, bk);
      tma_load_3d(sV, &mapV, kv_full, h * D, 0, bk);
    }
    __syncwarp();
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1;
      mbar_wait(&q_empty[s], ((i >> 1) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&q_full[s], kQBytes);
        tma_load_3d(sQ + s * kQBytes, &mapQ, &q_full[s], h * D, (t0 + i) * kTileQ, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 1);     // B (= V) is MN-major
    const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP), aK = smem_u32(sK), aV = smem_u32(sV);
    const int ksteps = (min(p.Tk, kTileK) + 15) >> 4;              // 16-key steps of P V that hold real keys
    auto issue_s = [&](int i) {
      const int s = i & 1;
      mbar_wait(&q_full[s], (i >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(tmem_S, umma_desc_k_sw128(aQ + s * kQBytes) + 2 * k, umma_desc_k_sw128(aK) + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        tc_commit(s_full);
        tc_commit(&q_empty[s]);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    if (n_my > 0) issue_s(0);
    for (int i = 0; i < n_my; ++i) {
      if (i + 1 < n_my) {
        mbar_wait(s_empty, i & 1);                  // every softmax thread holds its last S_i chunk in registers
        issue_s(i + 1);
      }
      mbar_wait(p_full, i & 1);                     // P_i is in shared memory (and O_{i-1} has been read by its rows' threads)
      tc_fence_after();
      if (elect_one_sync()) {
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t da = umma_desc_k_sw128(aP + (k >> 2) * (kTileQ * 128)) + 2 * (k & 3);
          const uint64_t db = umma_desc_mn_sw128(aV + k * 16 * 128, kTileK * 128);
          tc_mma_bf16(tmem_O, da, db, idesc_pv, k != 0 ? 1u : 0u);
        }
        tc_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // =============================== softmax + output: one thread per query row ===============================
    const int qd = warp & 3;                          // TMEM lane quarter of this warp (warps 2..5 -> 2, 3, 0, 1)
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tS = tmem_S + lane_off, tO = tmem_O + lane_off;
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    const int tk = min(p.Tk, kTileK);
    const int nchunk = (tk + 31) >> 5;                // 32-column score chunks that hold real keys
    for (int i = 0; i < n_my; ++i) {
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      float l0 = 0.f, l1 = 0.f;
      if (nchunk <= 3) {
        // ---- the prompt case (<= 96 keys): ONE TMEM pass, the whole score row in registers; S is released to the MMA
        //      warp before any arithmetic.  Chunks below nchunk - 1 are full (no key masking).
        uint32_t v[3][32];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < nchunk) tmem_ld32(tS + c * 32, v[c]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_empty);
        float raw = -INFINITY;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c < nchunk - 1) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) raw = max3(raw, __uint_as_float(v[c][j]), __uint_as_float(v[c][j + 1]));
          } else if (c == nchunk - 1) {
            const int kv = tk - c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < kv) raw = fmaxf(raw, __uint_as_float(v[c][j]));
          }
        }
        const float nmx = -raw * p.scale_log2;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (c < nchunk) {
            const int kv = tk - c * 32;
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float a0, a1;
              fma2(a0, a1, __uint_as_float(v[c][2 * j]), __uint_as_float(v[c][2 * j + 1]), p.scale_log2, nmx);
              float p0 = exp2_approx(a0), p1 = exp2_approx(a1);
              if (c == nchunk - 1) {                 // (warp-uniform) only the last chunk can hold keys that do not exist
                p0 = (2 * j < kv) ? p0 : 0.f;
                p1 = (2 * j + 1 < kv) ? p1 : 0.f;
              }
              add2(l0, l1, l0, l1, p0, p1);
              pk[j] = pack_bf16(p0, p1);
            }
            uint8_t* prow = sP + (c >> 1) * (kTileQ * 128) + r * 128;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + qq) ^ swz;
              *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
            }
          }
        }
      } else {
        // ---- pass 1: row maximum of the raw scores
        float raw = -INFINITY;
        for (int c = 0; c < nchunk; ++c) {
          uint32_t v[32];
          tmem_ld32(tS + c * 32, v);
          tmem_ld_wait();
          const int kv = tk - c * 32;
  #pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < kv) raw = fmaxf(raw, __uint_as_float(v[j]));
        }
        const float nmx = -raw * p.scale_log2;
        // ---- pass 2: P = exp2(s * scale - max) -> bf16 -> shared memory (K-major SW128), row sum
        for (int c = 0; c < nchunk; ++c) {
          uint32_t v[32];
          tmem_ld32(tS + c * 32, v);
          tmem_ld_wait();
          if (c == nchunk - 1) {                        // S_i is fully consumed: the MMA warp may overwrite it with S_{i+1}
            tc_fence_before();
            mbar_arrive(s_empty);
          }
          const int kv = tk - c * 32;
          uint32_t pk[16];
  #pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = (2 * j < kv) ? exp2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, nmx)) : 0.f;
            const float p1 = (2 * j + 1 < kv) ? exp2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, nmx)) : 0.f;
            l0 += p0;
            l1 += p1;
            pk[j] = pack_bf16(p0, p1);
          }
          // keys [32 c, 32 c + 32) = 4 chunks of 16 B at chunk index (c & 1) * 4 + q of row r in key block c >> 1
          uint8_t* prow = sP + (c >> 1) * (kTileQ * 128) + r * 128;
  #pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + qq) ^ swz;
            *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
          }
        }

      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      // ---- O_i = P_i V: normalise and store this thread's row (64 bf16 = 128 contiguous bytes)
      const float inv = 1.f / (l0 + l1);
      mbar_wait(o_full, i & 1);
      tc_fence_after();
      const int q = (t0 + i) * kTileQ + r;
      bf16* op = p.out + b * p.out_bs + static_cast<long long>(q) * p.ldo + h * D;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + c * 32, o);
        tmem_ld_wait();
        if (q < p.Tq) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
            w.y = pack_bf16(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
            w.z = pack_bf16(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
            w.w = pack_bf16(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
            reinterpret_cast<uint4*>(op)[c * 4 + j] = w;
          }
        }
      }
      tc_fence_before();       // the O reads above are ordered before the next p_full arrive (-> before P_{i+1} V overwrites O)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static int launch_attention_kv1(const AttnParams& p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                                cudaStream_t stream) {
  constexpr int smem = 2 * kTileK * 64 * 2 + 2 * kTileQ * 64 * 2 + kTileQ * kTileK * 2 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attention_kv1)");
    configured = true;
  }
  // query tiles per CTA: one wave of 2 CTAs per SM over all (image, head, tile) units
  const int tiles = (p.Tq + kTileQ - 1) / kTileQ;
  const long long total = static_cast<long long>(tiles) * p.heads * p.batch;
  int nq = static_cast<int>((total + 2LL * num_sms() - 1) / (2LL * num_sms()));
  static const int force_nq = getenv("UR_ATTN_NQ") ? atoi(getenv("UR_ATTN_NQ")) : 0;     // development
  if (force_nq > 0) nq = force_nq;
  if (nq < 1) nq = 1;
  if (nq > tiles) nq = tiles;
  dim3 grid((tiles + nq - 1) / nq, p.heads, p.batch);
  cudaError_t e = launch_kernel(attention_kv1_kernel, grid, dim3(192), smem, stream, p, mq, mk, mv, nq);
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "attention_kv1 launch");
}

template <int D>
static int launch_attention(const AttnParams& p, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                            dim3 grid, cudaStream_t stream) {
  constexpr int smem = kTileQ * D * 2 + 2 * 2 * kTileK * D * 2 + kTileQ * kTileK * 2 + 640;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attention)");
    configured = true;
  }
  cudaError_t e = launch_kernel(attention_kernel<D>, grid, dim3(320), smem, stream, p, mq, mk, mv);
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "attention launch");
}

static int make_qkv_map(CUtensorMap* m, const void* ptr, int width, int64_t ld, int64_t bs, int tokens, int batch) {
  const uint64_t dims[3] = {static_cast<uint64_t>(width), static_cast<uint64_t>(tokens), static_cast<uint64_t>(batch)};
  const uint64_t str[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(batch > 1 ? bs : ld * tokens) * 2};
  const uint32_t box[3] = {64u, 128u, 1u};
  const uint32_t es[3] = {1u, 1u, 1u};
  return encode_tensor_map(m, const_cast<void*>(ptr), 3, dims, str, box, es);
}

}  // namespace ur

using namespace ur;

static long long* g_attn_trace = nullptr;
static int g_attn_kv1 = 1;       // 1: head_dim 64 with all keys in one tile takes attention_kv1_kernel (query-tile loop)
extern "C" int ur_debug_set_attention_kv1(int on) {
  const int old = g_attn_kv1;
  g_attn_kv1 = on;
  return old;
}
static int g_attn_impl = 1;      // 1: attention_kernel (default: measured faster, r2c2); 2: attention2_kernel
static int g_attn_poly = kPolyOf8Default;
// development: exponential pairs (of every 8) evaluated by the FMA-pipe polynomial in attention2_kernel (0..3)
extern "C" int ur_debug_set_attention_poly(int n) {
  const int old = g_attn_poly;
  if (n >= 0 && n <= 3) g_attn_poly = n;
  return old;
}
// development: select the attention kernel generation (returns the previous one)
extern "C" int ur_debug_set_attention_impl(int impl) {
  const int old = g_attn_impl;
  if (impl == 1 || impl == 2) g_attn_impl = impl;
  return old;
}
// development: device buffer of 64 int64 receiving clock64 timestamps (nullptr = off)
extern "C" int ur_debug_set_attention_trace(void* buf) {
  g_attn_trace = static_cast<long long*>(buf);
  return 0;
}

extern "C" int ur_attention(const void* q, int64_t ldq, int64_t q_bs, const void* k, int64_t ldk, int64_t k_bs,
                            const void* v, int64_t ldv, int64_t v_bs, void* out, int64_t ldo, int64_t out_bs, int batch,
                            int heads, int head_dim, int tq, int tk, int kv_shared, float scale, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!q || !k || !v || !out) return set_error(UR_ERR_ARG, "ur_attention: null pointer");
  if (head_dim != 64 && head_dim != 128 && head_dim != 512)
    return set_error(UR_ERR_ARG, "ur_attention: head_dim must be 64, 128 or 512");
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || q_bs % 8 || k_bs % 8 || v_bs % 8 || out_bs % 8 ||
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(out)) & 15))
    return set_error(UR_ERR_ARG, "ur_attention: pitches must be multiples of 8 elements and pointers 16-byte aligned");
  if (batch <= 0 || heads <= 0 || tq <= 0 || tk <= 0) return set_error(UR_ERR_ARG, "ur_attention: bad sizes");
  const int width = heads * head_dim;
  CUtensorMap mq, mk, mv;
  int rc = make_qkv_map(&mq, q, width, ldq, q_bs, tq, batch);
  if (rc) return rc;
  const int kvb = kv_shared ? 1 : batch;
  rc = make_qkv_map(&mk, k, width, ldk, k_bs, tk, kvb);
  if (rc) return rc;
  rc = make_qkv_map(&mv, v, width, ldv, v_bs, tk, kvb);
  if (rc) return rc;
  AttnParams p;
  p.Tq = tq;
  p.Tk = tk;
  p.heads = heads;
  p.batch = batch;
  p.kv_shared = kv_shared;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = static_cast<bf16*>(out);
  p.ldo = ldo;
  p.out_bs = out_bs;
  p.trace = g_attn_trace;
  if (head_dim == 512) return launch_attention512(p, mq, mk, mv, stream);
  if (g_attn_impl == 2) {
    if (head_dim == 128) return launch_attention2<128, 1, kPolyOf8Default>(p, mq, mk, mv, stream);
    switch (g_attn_poly) {
      case 0: return launch_attention2<64, 2, 0>(p, mq, mk, mv, stream);
      case 1: return launch_attention2<64, 2, 1>(p, mq, mk, mv, stream);
      case 2: return launch_attention2<64, 2, 2>(p, mq, mk, mv, stream);
      default: return launch_attention2<64, 2, 3>(p, mq, mk, mv, stream);
    }
  }
  if (head_dim == 64 && tk <= kTileK && g_attn_kv1) return launch_attention_kv1(p, mq, mk, mv, stream);
  dim3 grid((tq + kTileQ - 1) / kTileQ, heads, batch);
  return head_dim == 64 ? launch_attention<64>(p, mq, mk, mv, grid, stream)
                        : launch_attention<128>(p, mq, mk, mv, grid, stream);
}
