// Persistent implicit-GEMM convolution / linear kernel (the fast path of ur_conv_gemm).
//
//   one CTA per SM, static round-robin over the (m_tile, n_tile) list (n fastest, so CTAs running side by side
//   share the activation tile in L2);  14 warps:
//     warps 0..2  activation (A) TMA producers, k-block kb is issued by warp kb % 3   } a single thread issuing every
//     warps 3..4  weight (W) TMA producers,     k-block kb is issued by warp 3 + kb%2 } TMA costs ~700 cycles per
//                 k-block (the 4-D tiled load alone ~390) and starves the tensor core, hence the fan-out
//     warp  5     TMEM allocator + single-thread tcgen05.mma issuer; accumulators double-buffered in TMEM
//                 (2 x BN fp32 columns): the epilogue of tile i overlaps the main loop of tile i+1
//     warps 6..13 epilogue, two warp groups of 128 threads; group g owns the 32-column sub-blocks g, g+2, ...
//                 tcgen05.ld -> +bias/+temb (staged in smem) -> activation / GEGLU / gate -> channel scale
//                 -> +residual -> bf16 -> padded smem tile -> coalesced 16-byte global stores (8 rows x 64 B per
//                 warp instruction instead of 32 scattered rows)
//     warps 14..15 (lean kernel) TMA-store issuers, one per epilogue warp group: issuing a 4-D tiled TMA store costs its
//                 thread ~340 cycles, a third of a 32-column sub-block when lane 0 of an epilogue warp does it while the
//                 other 31 lanes wait; the epilogue warps only fence + arrive on an mbarrier and go on
//   smem ring of STAGES x (A 128x64 bf16 + W BNx64 bf16), 128-byte swizzle, full/empty mbarriers.
//
// Same arithmetic and operand layout as ur_gemm.cu (see there / include/unirestore_b200.h for the reference
// call sites); this variant requires a bf16 output whose pitches are multiples of 8 elements.
#include "ur_gemm.h"

namespace ur {

// Epilogue staging: per warp group TWO buffers of 128 rows x 64 B (32 bf16) in the TMA 64-byte-swizzle layout
// (16-byte chunk index ^= (row >> 1) & 3): conflict-free row-wise 16-byte writes AND the box layout of the TMA store.
constexpr int kEpiBufBytes = kBlockM * 64;           // 8 KB
constexpr int kEpiStageBytes = 2 * kEpiBufBytes;     // per epilogue warp group
__device__ __forceinline__ int stg_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }
constexpr int kNumAProd = 3, kNumWProd = 2;
constexpr int kMmaWarp = kNumAProd + kNumWProd;      // 5
constexpr int kEpiWarp0 = kMmaWarp + 1;              // 6
constexpr int kStoreWarp0 = kEpiWarp0 + 8;           // 14: one TMA-store issuer per epilogue warp group (lean kernel)
constexpr int kThreads = (kStoreWarp0 + 2) * 32;     // 512

// PAIR = true: CTA pairs (cluster of 2) run tcgen05.mma.cta_group::2 on a 256 x BN tile; each CTA stages its own
// 128 activation rows and HALF of the weight rows, which halves the weight traffic into shared memory (the main
// loop of the single-CTA kernel is bound by shared-memory bandwidth: TMA write + UMMA read of A and W).
template <int BN, bool PAIR>
struct PCfg {
  static constexpr int kWRows = PAIR ? BN / 2 : BN;
  static constexpr int kWBytes = kWRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kFixed = 2 * kEpiStageBytes + 8 * BN * 4 + 4 * 128 * 8 + 512;
  static constexpr int kFit = (227 * 1024 - kFixed) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr int kSmem = kStages * kStageBytes + kFixed;
};

struct TileCoord {
  int n0, x0, y0, b0;
};
// work unit u -> tile of this CTA: n tile = u % n_tiles (fastest), m tile = u / n_tiles (PAIR: 2 * that + cta rank)
__device__ __forceinline__ TileCoord tile_coord(const GemmParams& p, int u, int n_tiles, int BN, int Wt, int Ht, int Bt,
                                                int pair_rank) {
  TileCoord c;
  int mt = fast_div(u, p.fd_ntiles);
  const int nt = u - mt * n_tiles;
  if (pair_rank >= 0) mt = 2 * mt + pair_rank;
  c.n0 = nt * BN;
  const int q1 = fast_div(mt, p.fd_tx);
  const int tx = mt - q1 * p.tiles_x;
  const int tb = fast_div(q1, p.fd_ty);
  const int ty = q1 - tb * p.tiles_y;
  c.x0 = tx * Wt;
  c.y0 = ty * Ht;
  c.b0 = tb * Bt;
  return c;
}

__device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// LEAN = true: the epilogue of the common case only -- bias / temb row vector, optional residual and fused statistics,
// per-warp TMA stores, no activation / gate / channel scale / alpha, no split-K -- so that the instruction footprint of
// the ~80 % of launches that need nothing else is a third of the general kernel's (the 14 warps run four different
// roles through one instruction cache; r2 ncu: 25 % of the epilogue warps' issue stalls were stall_no_inst).
template <int BN, bool PAIR, bool LEAN>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_persistent_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ CUtensorMap mapA1,
                            const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW,
                            const __grid_constant__ CUtensorMap mapOut, int total_tiles, int n_tiles) {
  using Cfg = PCfg<BN, PAIR>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int kStageBytes = Cfg::kStageBytes;
  constexpr uint32_t kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* staging = smem + STAGES * kStageBytes;                          // [2 groups][128 x 80 B]
  float* s_add = reinterpret_cast<float*>(staging + 2 * kEpiStageBytes);    // [2 groups][2][BN] (a private copy per
  float* s_mul = s_add + 4 * BN;                                            //  warp group: the groups never sync)
  float* s_stat = s_mul + 4 * BN;                                           // [2 groups][2 buffers][128] float2 (epilogue statistics)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stat + 4 * 128 * 2);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint64_t* store_req = tempty_bar + 2;       // [2 groups][2 staging buffers]: the group's 4 warps have staged a sub-block
  uint64_t* store_done = store_req + 4;       // [2][2]: the TMA store issued from the buffer has finished reading it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(store_done + 4);

  const int warp = warp_uniform(static_cast<int>(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  const int Wt = 1 << p.wt_log2, Ht = 1 << p.ht_log2;
  const int Bt = kBlockM >> (p.wt_log2 + p.ht_log2);
  const int nkb = p.ntaps * p.cblocks;
  const int ksplit = p.ksplit;
  // scheduling: work units (tiles, or 256-row pair tiles) are dealt round-robin to CTAs (or CTA pairs)
  const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : -1;
  const int u_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int u_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const bool tracing = !LEAN && p.trace != nullptr && blockIdx.x == 0;     // (the lean kernel carries no probes)

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
      mbar_init(&tempty_bar[s], PAIR ? 512 : 256);      // PAIR: the peer's epilogue threads arrive remotely on rank 0
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&store_req[s], 4);
      mbar_init(&store_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if constexpr (PAIR) {
      tmem_alloc_2sm(tmem_slot, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();               // peer barriers initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);
  pdl_launch_dependents();
  pdl_wait();                      // prologue done; from here on global memory written by the predecessor is read
  if (tracing && threadIdx.x == 0) p.trace[7 * 16] = p.trace[7 * 16 + 1] = clock64();

  if (warp < kNumAProd) {
    // =============================== activation TMA producers (whole warp, elected lane issues) ===============
    int kiter = 0, it = 0;
    for (int t = u_first; t < total_tiles; t += u_stride, ++it) {
      const int tile = fast_div(t, p.fd_ksplit), ks = t - tile * ksplit;
      const int kb0 = fast_div(ks * nkb, p.fd_ksplit), kb1 = fast_div((ks + 1) * nkb, p.fd_ksplit);          // this unit's K slice (split-K)
      const TileCoord tc = tile_coord(p, tile, n_tiles, BN, Wt, Ht, Bt, rank);
      const int cbase = p.group_kc ? (tc.n0 / p.group_nc) * p.group_kc : 0;
      if (tracing && lane == 0 && warp == 0 && it < 8) p.trace[0 * 16 + it] = clock64();
      int tap = 0, cb = 0;
      if (kb0) {                                   // split-K slices start mid-way
        tap = kb0 / p.cblocks;
        cb = kb0 - tap * p.cblocks;
      }
      for (int kb = kb0; kb < kb1; ++kb, ++kiter) {
        if (kiter % kNumAProd == warp) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          const bool trk = tracing && lane == 0 && warp == 0 && it == 1 && kb >= 6 && kb < 6 + 3 * 16;
          if (trk) p.trace[192 + 4 * ((kb - 6) / 3)] = clock64();
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (trk) p.trace[192 + 4 * ((kb - 6) / 3) + 1] = clock64();
          const int dy = static_cast<int>((p.dy_pack >> (4 * tap)) & 15) - 8;
          const int dx = static_cast<int>((p.dx_pack >> (4 * tap)) & 15) - 8;
          const int c = cbase + cb * kBlockK;
          const int xi = tc.x0 * p.stride + dx, yi = tc.y0 * p.stride + dy;
          uint8_t* sa = smem + s * kStageBytes;
          const bool src1 = c < p.c1;
          const CUtensorMap* mp = src1 ? &mapA1 : &mapA2;
          const int cc = src1 ? c : c - p.c1;
          if (elect_one_sync()) {
            // one arming arrive per stage (rank 0 only in PAIR mode), covering A and W bytes of both CTAs
            if (!PAIR || rank == 0) mbar_expect_tx(&full_bar[s], PAIR ? 2 * kStageBytes : kStageBytes);
            if constexpr (PAIR)
              tma_load_4d_2sm(sa, mp, &full_bar[s], cc, xi, yi, tc.b0);
            else
              tma_load_4d(sa, mp, &full_bar[s], cc, xi, yi, tc.b0);
          }
          __syncwarp();
          if (trk) p.trace[192 + 4 * ((kb - 6) / 3) + 2] = clock64();
        }
        if (++cb == p.cblocks) {
          cb = 0;
          ++tap;
        }
      }
    }
  } else if (warp < kMmaWarp) {
    // =============================== weight TMA producers ===============================
    const int me = warp - kNumAProd;
    int kiter = 0, it = 0;
    for (int t = u_first; t < total_tiles; t += u_stride, ++it) {
      const int tile = fast_div(t, p.fd_ksplit), ks = t - tile * ksplit;
      const int kb0 = fast_div(ks * nkb, p.fd_ksplit), kb1 = fast_div((ks + 1) * nkb, p.fd_ksplit);
      const TileCoord tc = tile_coord(p, tile, n_tiles, BN, Wt, Ht, Bt, rank);
      const int wb = p.w_batched ? tc.b0 : 0;
      const int wrow = PAIR ? tc.n0 + rank * (BN / 2) : tc.n0;     // PAIR: this CTA stages half of the N rows
      int tap = 0, cb = 0;
      if (kb0) {                                   // split-K slices start mid-way
        tap = kb0 / p.cblocks;
        cb = kb0 - tap * p.cblocks;
      }
      for (int kb = kb0; kb < kb1; ++kb, ++kiter) {
        if (kiter % kNumWProd == me) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          const bool trk = tracing && lane == 0 && me == 0 && it == 1 && kb >= 6 && kb < 6 + 2 * 16;
          if (trk) p.trace[256 + 4 * ((kb - 6) / 2)] = clock64();
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (trk) p.trace[256 + 4 * ((kb - 6) / 2) + 1] = clock64();
          if (elect_one_sync()) {
            if constexpr (PAIR)
              tma_load_3d_2sm(smem + s * kStageBytes + kABytes, &mapW, &full_bar[s], tap * p.kc + cb * kBlockK, wrow, wb);
            else
              tma_load_3d(smem + s * kStageBytes + kABytes, &mapW, &full_bar[s], tap * p.kc + cb * kBlockK, wrow, wb);
          }
          __syncwarp();
          if (trk) p.trace[256 + 4 * ((kb - 6) / 2) + 2] = clock64();
        }
        if (++cb == p.cblocks) {
          cb = 0;
          ++tap;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer ===============================
    // whole warp, warp-uniform loop; only the tcgen05 instructions are predicated on the elected lane
    if (!PAIR || rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : kBlockM, BN);
      const uint32_t a_lo0 = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (1u << 16);   // descriptor low word of stage 0 (LBO = 1)
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);       // SBO 1024 B, version 1, SW128
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int t = u_first; t < total_tiles; t += u_stride, ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        if (tracing && lane == 0 && it < 8) p.trace[1 * 16 + it] = clock64();
        // Two k-blocks per iteration: the fixed cost of one trip through the issue path (barrier test, fence, elect,
        // commit: ~250 cycles measured) is paid once per 8 MMAs.  The full-barrier tests of the NEXT two k-blocks are
        // issued before the MMAs of the current ones so that their ~50-100 cycle latency is off the issue path.
        const int ks = t - fast_div(t, p.fd_ksplit) * ksplit;
        const int nk = fast_div((ks + 1) * nkb, p.fd_ksplit) - fast_div(ks * nkb, p.fd_ksplit);      // k-blocks of this unit
        bool ready0 = mbar_test_wait(&full_bar[stage], phase);
        bool ready1 = false;
        for (int kb = 0; kb < nk; kb += 2) {
          const bool two = kb + 1 < nk;
          const bool trk = tracing && lane == 0 && it == 1 && kb >= 8 && kb < 40;
          if (trk) p.trace[128 + 4 * ((kb - 8) >> 1)] = clock64();
          const uint32_t s0 = stage, ph0 = phase;
          uint32_t s1 = s0 + 1, ph1 = ph0;
          if (s1 == STAGES) {
            s1 = 0;
            ph1 ^= 1;
          }
          if (!ready0) mbar_wait(&full_bar[s0], ph0);
          if (two && !ready1) mbar_wait(&full_bar[s1], ph1);
          if (trk) p.trace[128 + 4 * ((kb - 8) >> 1) + 1] = clock64();
          if (tracing && lane == 0 && it < 8 && kb == 0) p.trace[2 * 16 + it] = clock64();
          tc_fence_after();
          // advance the ring past the k-blocks consumed here and pre-test the next two
          stage = s1;
          phase = ph1;
          if (two && ++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
          {
            uint32_t n1 = stage + 1, nph1 = phase;
            if (n1 == STAGES) {
              n1 = 0;
              nph1 ^= 1;
            }
            ready0 = (kb + 2 < nk) ? mbar_test_wait(&full_bar[stage], phase) : false;
            ready1 = (kb + 3 < nk) ? mbar_test_wait(&full_bar[n1], nph1) : false;
          }
          const uint32_t alo0 = a_lo0 + s0 * (kStageBytes >> 4);           // (smem address >> 4) of the A tiles
          const uint32_t alo1 = a_lo0 + s1 * (kStageBytes >> 4);
          if (elect_one_sync()) {
            tc_mma_bf16_x4<PAIR>(tacc, kDescHi, alo0, alo0 + (kABytes >> 4), idesc, kb != 0 ? 1u : 0u);
            if constexpr (PAIR) tc_commit_2sm_mc(&empty_bar[s0]); else tc_commit(&empty_bar[s0]);
            if (two) {
              tc_mma_bf16_x4<PAIR>(tacc, kDescHi, alo1, alo1 + (kABytes >> 4), idesc, 1u);
              if constexpr (PAIR) tc_commit_2sm_mc(&empty_bar[s1]); else tc_commit(&empty_bar[s1]);
            }
          }
          if (trk) p.trace[128 + 4 * ((kb - 8) >> 1) + 2] = clock64();
        }
        if (elect_one_sync()) {
          if constexpr (PAIR) tc_commit_2sm_mc(&tfull_bar[as]); else tc_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (tracing && lane == 0 && it < 8) p.trace[3 * 16 + it] = clock64();
      }
    }
  } else if (warp < kStoreWarp0) {
    // =============================== epilogue ===============================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int grp = (warp - kEpiWarp0) >> 2;      // warp group 0 / 1
    const int r = q * 32 + lane;                  // tile row = TMEM lane
    const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..255
    const int gt = et & 127;                      // thread index inside the warp group
    const bool gated = !LEAN && (p.act == UR_ACT_GEGLU || p.act == UR_ACT_GATE);
    const int n_out = gated ? (p.N >> 1) : p.N;
    const int ncols = gated ? (BN >> 1) : BN;
    const bool has_mul = !LEAN && p.chscale != nullptr;
    const bool plain = LEAN || (!gated && p.act == UR_ACT_NONE && !has_mul && p.alpha == 1.0f);
    // cooperative phase: thread gt moves 16-byte chunk (gt & 3) of rows (gt >> 2) + 32 i, i = 0..3
    const int cchunk = gt & 3;
    const int crow0 = gt >> 2;
    uint8_t* const stg_base = staging + grp * kEpiStageBytes;
    int sbuf = 0;                                 // staging buffer of the next sub-block (alternates)
    const bool use_tma = LEAN || p.tma_store != 0;
    // lean kernel: "store-warp mode" -- the four warps of a group stage their rows, fence, arrive on store_req, and warp
    // 14 + grp issues ONE TMA store per sub-block (128-row box) and reports on store_done when the buffer is free again
    constexpr bool sw_mode = LEAN;
    uint32_t nsub = 0;                            // sub-blocks this group has staged so far (buffer nsub & 1)
    // tma_store == 2 ("warp mode"): the output tensor map has a 32-row box and every epilogue WARP stores its own 32
    // rows (= its TMEM lane quarter) as soon as it has written them: no barrier between the four warps of a group on
    // the store path (the two 128-thread barriers per 32-column sub-block cost ~300 of its ~1 200 cycles, and the K <= 640
    // linears / GEGLU layers are epilogue-bound).  The residual is then staged by the warp for its own rows, too.
    const bool warp_mode = !LEAN && p.tma_store == 2;
    const bool warp_rows = LEAN || warp_mode;     // every warp moves / stages only its own 32 rows (no group barriers)
    const int wrow0 = q * 32;                     // first tile row of this warp
    const int mrow0 = warp_rows ? wrow0 + (lane >> 2) : crow0;       // rows this thread moves: mrow0 + mstep * i
    const int mstep = warp_rows ? 8 : 32;
    const int mchunk = warp_rows ? (lane & 3) : cchunk;
    bf16* outp = reinterpret_cast<bf16*>(p.out);
    // this thread's two columns (gt, gt + 128) of the NEXT tile's add / mul vectors (every warp group stages its own copy)
    // (bias and row vector are kept as loaded and only added when they are staged, one tile later: `a += __ldg(...)`
    //  made the add wait for the load right here, 500-1 000 cycles of exposed L2 latency per tile -- clock64 probes r2)
    float a_nx[2] = {0.f, 0.f}, r_nx[2] = {0.f, 0.f}, m_nx[2] = {1.f, 1.f};
    auto fetch_vec = [&](const TileCoord& tn) {
      const int no0 = gated ? (tn.n0 >> 1) : tn.n0;
      const int bb = tn.b0 < p.B ? tn.b0 : p.B - 1;      // (a PAIR ghost tile lies past the last image)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = gt + 128 * h;
        const int n = tn.n0 + col;
        const bool okn = col < BN && n < p.N;
        a_nx[h] = (okn && p.bias) ? __ldg(p.bias + n) : 0.f;
        r_nx[h] = (okn && p.rowvec) ? __ldg(p.rowvec + bb * p.rowvec_sb + n) : 0.f;
        m_nx[h] = (has_mul && col < ncols && no0 + col < n_out) ? __ldg(p.chscale + bb * p.chscale_sb + no0 + col) : 1.f;
      }
    };
    // (the coordinates of the next tile are computed once, for its vector fetch, and carried into the next iteration:
    //  the epilogue is bound by its instruction count, r2 profiles/epilogue_store_experiments_r2.txt)
    TileCoord tc = tile_coord(p, fast_div(u_first < total_tiles ? u_first : 0, p.fd_ksplit), n_tiles, BN, Wt, Ht, Bt, rank);
    TileCoord tn = tc;
    if (u_first < total_tiles) fetch_vec(tc);
    int it = 0;
    for (int t = u_first; t < total_tiles; t += u_stride, ++it, tc = tn) {
      const bool has_next = t + u_stride < total_tiles;
      if (has_next) tn = tile_coord(p, fast_div(t + u_stride, p.fd_ksplit), n_tiles, BN, Wt, Ht, Bt, rank);
      const int nout0 = gated ? (tc.n0 >> 1) : tc.n0;
      const int as = it & 1;
      if (!LEAN && p.ws) {
        // ---- split-K: this unit's fp32 partial tile goes to slab `ks` of the workspace with plain 16-byte stores
        //      (every element of every slab is written exactly once: no zero fill, no atomics, and the sum order in
        //      splitk_finish_kernel is fixed -> bit-reproducible); bias / residual / bf16 conversion / GroupNorm
        //      statistics happen once in splitk_finish_kernel
        mbar_wait(&tfull_bar[as], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t trow = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
        const int x = tc.x0 + (r & (Wt - 1)), y = tc.y0 + ((r >> p.wt_log2) & (Ht - 1));
        const int b = tc.b0 + (r >> (p.wt_log2 + p.ht_log2));
        const bool ok = (x < p.Wo) && (y < p.Ho) && (b < p.B);
        const int ksl = t - fast_div(t, p.fd_ksplit) * ksplit;
        float* wrow = p.ws + ksl * p.ws_slab + (static_cast<long long>(b * p.Ho + y) * p.Wo + x) * p.N + tc.n0;
        for (int c = grp * 32; c < BN; c += 64) {
          uint32_t va[32];
          tmem_ld32(trow + c, va);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (tc.n0 + c + 4 * j + 4 <= p.N)
                *reinterpret_cast<float4*>(wrow + c + 4 * j) =
                    make_float4(__uint_as_float(va[4 * j]), __uint_as_float(va[4 * j + 1]),
                                __uint_as_float(va[4 * j + 2]), __uint_as_float(va[4 * j + 3]));
            }
          }
        }
        tc_fence_before();
        if constexpr (PAIR) mbar_arrive_cluster(&tempty_bar[as], 0); else mbar_arrive(&tempty_bar[as]);
        continue;
      }
      // ---- stage the per-column add / mul vectors of this tile (double-buffered by `as`); the values were
      //      fetched one tile ahead so their global-load latency hides behind the previous tile's epilogue
      float* add = s_add + (grp * 2 + as) * BN;
      float* mul = s_mul + (grp * 2 + as) * BN;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (gt + 128 * h < BN) {
          add[gt + 128 * h] = a_nx[h] + r_nx[h];
          if (!LEAN) mul[gt + 128 * h] = m_nx[h];
        }
      }
      if (has_next) fetch_vec(tn);
      // global element offsets of the 4 rows this thread moves in the cooperative phases (-1: outside the tensor)
      long long roff[4] = {-1, -1, -1, -1};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!p.residual && (use_tma || LEAN)) break;        // only the residual fetch and the st.global path use them
        const int rr = mrow0 + mstep * i;
        const int x = tc.x0 + (rr & (Wt - 1)), y = tc.y0 + ((rr >> p.wt_log2) & (Ht - 1));
        const int b = tc.b0 + (rr >> (p.wt_log2 + p.ht_log2));
        const bool ok = (x < p.Wo) && (y < p.Ho) && (b < p.B);
        roff[i] = ok ? (b * p.res_sb + y * p.res_sy + x * p.res_sx + nout0) : -1;
      }
      // warp mode: output coordinates of this warp's first row (its 32 rows form one box of the output tensor map)
      const int wx = tc.x0 + (wrow0 & (Wt - 1)), wy = tc.y0 + ((wrow0 >> p.wt_log2) & (Ht - 1));
      const int wb = tc.b0 + (wrow0 >> (p.wt_log2 + p.ht_log2));
      auto ooff = [&](int i) {                    // output offset of cooperative row i (non-TMA store path only)
        const int rr = crow0 + 32 * i;
        const int x = tc.x0 + (rr & (Wt - 1)), y = tc.y0 + ((rr >> p.wt_log2) & (Ht - 1));
        const int b = tc.b0 + (rr >> (p.wt_log2 + p.ht_log2));
        return b * p.out_sb + y * p.out_sy + x * p.out_sx + nout0;
      };
      // residual of this group's first sub-block, fetched (coalesced) while the main loop still runs
      uint4 rres[4];
      int c = ((grp + it) & 1) * 32;            // the groups alternate which one takes the odd sub-block out (balance)
      if (p.residual) {
        const int col = c + mchunk * 8;
        const bool okc = c < ncols && nout0 + col + 8 <= n_out;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          rres[i] = (okc && roff[i] >= 0) ? __ldg(reinterpret_cast<const uint4*>(p.residual + roff[i] + col))
                                           : make_uint4(0u, 0u, 0u, 0u);
      }
      group_barrier(10 + grp);                  // this group's add / mul copy is staged
      if (tracing && it < 8 && et == 0) p.trace[4 * 16 + it] = clock64();
      mbar_wait(&tfull_bar[as], (it >> 1) & 1);
      if (tracing && it < 8 && et == 0) p.trace[5 * 16 + it] = clock64();
      tc_fence_after();
      const uint32_t trow = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);

      for (; c < ncols; c += 64, sbuf ^= 1) {
        const bool tre = tracing && it == 1 && et == 0 && c < 192;
        if (tre) p.trace[320 + (c >> 6) * 8] = clock64();
        uint8_t* stg = stg_base + sbuf * kEpiBufBytes;
        uint32_t va[32];
        tmem_ld32(trow + c, va);
        // (A) this staging buffer is free again: the TMA store issued from it two sub-blocks ago has read it
        if (sw_mode) {
          mbar_wait(&store_done[grp * 2 + sbuf], ((nsub >> 1) & 1) ^ 1);
        } else if (warp_mode) {
          if (lane == 0) bulk_wait_group_read<1>();
          __syncwarp();
        } else {
          if (use_tma && gt == 0) bulk_wait_group_read<1>();
          group_barrier(2 + grp);
        }
        if (tre) p.trace[320 + (c >> 6) * 8 + 1] = clock64();
        if (p.residual) {
          // park the prefetched (coalesced) residual chunk in the staging buffer
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(stg + stg_off(mrow0 + mstep * i, mchunk)) = rres[i];
          if (warp_rows) __syncwarp(); else group_barrier(4 + grp);
          // prefetch the residual of this group's next sub-block (the proxy fence below = MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC
          // waits for these loads if they are still in flight; r2 ablation: most of the +5.6 us a residual costs on the
          // 64x64-level linears is its 21 MB of traffic, ~2 us is exposed latency)
          const int cn = c + 64;
          const int col = cn + mchunk * 8;
          const bool okc = cn < ncols && nout0 + col + 8 <= n_out;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            rres[i] = (okc && roff[i] >= 0) ? __ldg(reinterpret_cast<const uint4*>(p.residual + roff[i] + col))
                                             : make_uint4(0u, 0u, 0u, 0u);
        }
        // accumulators -> (+bias / temb) -> activation / gate -> (* channel scale) -> (+residual) -> bf16 -> staging,
        // EIGHT columns (one 16-byte staging chunk) at a time: the gated / GELU epilogues are issue- and latency-bound,
        // and with all 32 columns of a, g and the results live (96 registers) the scheduler had no registers left to
        // overlap the activation chains (GEGLU with K = 320 ran at 0.19 of the tensor roof, r2 roofline); a chunk keeps
        // ~24 values live and its 4 pairs run on packed f32x2 instructions
        uint32_t vg[32];
        if (gated) tmem_ld32(trow + (BN >> 1) + c, vg);
        tmem_ld_wait();
        if (tre) p.trace[320 + (c >> 6) * 8 + 2] = clock64();
        const f2 al2 = splat2(p.alpha);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f2 f[4];
          if (plain) {
            const float4 ad0 = *reinterpret_cast<const float4*>(add + c + 8 * j);
            const float4 ad1 = *reinterpret_cast<const float4*>(add + c + 8 * j + 4);
            f[0] = f2{__uint_as_float(va[8 * j]) + ad0.x, __uint_as_float(va[8 * j + 1]) + ad0.y};
            f[1] = f2{__uint_as_float(va[8 * j + 2]) + ad0.z, __uint_as_float(va[8 * j + 3]) + ad0.w};
            f[2] = f2{__uint_as_float(va[8 * j + 4]) + ad1.x, __uint_as_float(va[8 * j + 5]) + ad1.y};
            f[3] = f2{__uint_as_float(va[8 * j + 6]) + ad1.z, __uint_as_float(va[8 * j + 7]) + ad1.w};
          } else {
            const float4 ad0 = *reinterpret_cast<const float4*>(add + c + 8 * j);
            const float4 ad1 = *reinterpret_cast<const float4*>(add + c + 8 * j + 4);
            const f2 adp[4] = {f2{ad0.x, ad0.y}, f2{ad0.z, ad0.w}, f2{ad1.x, ad1.y}, f2{ad1.z, ad1.w}};
#pragma unroll
            for (int k = 0; k < 4; ++k)
              f[k] = fma2(f2{__uint_as_float(va[8 * j + 2 * k]), __uint_as_float(va[8 * j + 2 * k + 1])}, al2, adp[k]);
            if (gated) {
              const float4 ag0 = *reinterpret_cast<const float4*>(add + (BN >> 1) + c + 8 * j);
              const float4 ag1 = *reinterpret_cast<const float4*>(add + (BN >> 1) + c + 8 * j + 4);
              const f2 agp[4] = {f2{ag0.x, ag0.y}, f2{ag0.z, ag0.w}, f2{ag1.x, ag1.y}, f2{ag1.z, ag1.w}};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                f2 g = fma2(f2{__uint_as_float(vg[8 * j + 2 * k]), __uint_as_float(vg[8 * j + 2 * k + 1])}, al2, agp[k]);
                if (p.act == UR_ACT_GEGLU) g = gelu_erf2(g);
                f[k] = mul2(f[k], g);
              }
            } else if (p.act == UR_ACT_SILU) {
#pragma unroll
              for (int k = 0; k < 4; ++k) f[k] = f2{silu_f(f[k].x), silu_f(f[k].y)};
            } else if (p.act == UR_ACT_GELU) {
#pragma unroll
              for (int k = 0; k < 4; ++k) f[k] = gelu_erf2(f[k]);
            }
            if (has_mul) {
              const float4 m0 = *reinterpret_cast<const float4*>(mul + c + 8 * j);
              const float4 m1 = *reinterpret_cast<const float4*>(mul + c + 8 * j + 4);
              f[0] = mul2(f[0], f2{m0.x, m0.y});
              f[1] = mul2(f[1], f2{m0.z, m0.w});
              f[2] = mul2(f[2], f2{m1.x, m1.y});
              f[3] = mul2(f[3], f2{m1.z, m1.w});
            }
          }
          // own row: (+ residual) -> bf16 -> staging (swizzled)
          uint4* sp = reinterpret_cast<uint4*>(stg + stg_off(r, j));
          if (p.residual) {
            const uint4 rv = *sp;
            const uint32_t u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float a0, a1;
              unpack_bf16(u[k], a0, a1);
              f[k].x += a0;
              f[k].y += a1;
            }
          }
          *sp = make_uint4(pack_bf16(f[0].x, f[0].y), pack_bf16(f[1].x, f[1].y), pack_bf16(f[2].x, f[2].y),
                           pack_bf16(f[3].x, f[3].y));
        }
        if (tre) p.trace[320 + (c >> 6) * 8 + 3] = clock64();
        if (sw_mode) {
          // (C) hand the staged rows to the group's store warp (release: proxy fence, then one arrive per warp)
          fence_proxy_async_smem();                 // generic-proxy writes above -> visible to the async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(&store_req[grp * 2 + sbuf]);
          ++nsub;
        } else if (warp_mode) {
          // (C) one TMA store per warp and sub-block: box (32 ch, 32 rows of this warp), clipped against the tensor
          fence_proxy_async_smem();                 // generic-proxy writes above -> visible to the async proxy
          __syncwarp();
          if (tre) p.trace[320 + (c >> 6) * 8 + 4] = clock64();
          if (lane == 0) {
            tma_store_4d(&mapOut, stg + wrow0 * 64, nout0 + c, wx, wy, wb);
            bulk_commit_group();
          }
        } else if (use_tma) {
          // (C) one TMA store per sub-block: the box (32 ch, Wt, Ht, Bt) is clipped against the output tensor
          fence_proxy_async_smem();                 // generic-proxy writes above -> visible to the async proxy
          group_barrier(6 + grp);
          if (tre) p.trace[320 + (c >> 6) * 8 + 4] = clock64();
          if (gt == 0) {
            tma_store_4d(&mapOut, stg, nout0 + c, tc.x0, tc.y0, tc.b0);
            bulk_commit_group();
          }
        } else {
          group_barrier(6 + grp);
          if (tre) p.trace[320 + (c >> 6) * 8 + 4] = clock64();
          // (C) coalesced write-out: a warp stores 8 rows x 64 contiguous bytes per instruction
          const int col = c + cchunk * 8;
          if (nout0 + col + 8 <= n_out) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (roff[i] >= 0)
                *reinterpret_cast<uint4*>(outp + ooff(i) + col) =
                    *reinterpret_cast<const uint4*>(stg + stg_off(crow0 + 32 * i, cchunk));
            }
          } else if (nout0 + col < n_out) {               // ragged channel tail (n_out % 8 != 0 never reaches here)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (roff[i] >= 0) {
                const bf16* sv = reinterpret_cast<const bf16*>(stg + stg_off(crow0 + 32 * i, cchunk));
                for (int k = 0; k < 8; ++k)
                  if (nout0 + col + k < n_out) outp[ooff(i) + col + k] = sv[k];
              }
            }
          }
        }
        if (p.stats) {
          // ---- fused GroupNorm statistics: column sums / sums of squares of the bf16 tile just staged.  Thread gt
          //      takes column (gt & 31) over rows 32 * (gt >> 5) .. + 31 (one conflict-free 2-byte LDS per row; the
          //      swizzle only depends on (row >> 1) & 3, so four base pointers cover all rows); the four row quarters
          //      meet in shared memory and one warp adds them to the fp64 statistics of the image(s) of the tile
          // (the quarter is the warp's OWN TMEM lane quarter q -- the rows its lanes just wrote -- not its position in
          //  the group: group 0 is warps 6..9 = quarters 2, 3, 0, 1; reading another warp's rows needed the group barrier
          //  that warp mode removed: racecheck r2c11)
          const int scol = lane, sq4 = q;
          const uint8_t* sb0 = stg + sq4 * 32 * 64 + (scol & 7) * 2;
          const int ch = scol >> 3;
          const uint8_t* sbk[4] = {sb0 + ((ch ^ 0) << 4), sb0 + ((ch ^ 1) << 4), sb0 + ((ch ^ 2) << 4), sb0 + ((ch ^ 3) << 4)};
          float s1 = 0.f, s2 = 0.f;
          const bool tile_inside = tc.x0 + Wt <= p.Wo && tc.y0 + Ht <= p.Ho;
          if (tile_inside) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = __uint_as_float(
                  static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(sbk[(i >> 1) & 3] + i * 64)) << 16);
              s1 += v;
              s2 = fmaf(v, v, s2);
            }
          } else {
            for (int i = 0; i < 32; ++i) {
              const int rr = sq4 * 32 + i;
              const int x = tc.x0 + (rr & (Wt - 1)), y = tc.y0 + ((rr >> p.wt_log2) & (Ht - 1));
              float v = __uint_as_float(
                  static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(sbk[(i >> 1) & 3] + i * 64)) << 16);
              if (x >= p.Wo || y >= p.Ho) v = 0.f;
              s1 += v;
              s2 = fmaf(v, v, s2);
            }
          }
          // (double-buffered by sub-block: without the store-path barriers a fast warp may write the next sub-block's
          //  partials while warp 0 still reads these; the barrier of the sub-block in between orders the reuse)
          float2* sred = reinterpret_cast<float2*>(s_stat) + (grp * 2 + sbuf) * 128;
          sred[sq4 * 32 + scol] = make_float2(s1, s2);
          group_barrier(8 + grp);
          if (gt < 32) {
            // the four 32-row quarters belong to Bt = 1, 2 or 4 consecutive images (a quarter never spans two: the
            // host only fuses the statistics when an image contributes >= 32 rows to the tile)
            const int ncol = nout0 + c + gt;
            const int qpi = 4 / Bt;                 // quarters per image
            if (ncol < n_out) {
              for (int bi = 0; bi < Bt; ++bi) {
                float t1 = 0.f, t2 = 0.f;
                for (int qq = bi * qpi; qq < (bi + 1) * qpi; ++qq) {
                  const float2 a = sred[gt + 32 * qq];
                  t1 += a.x;
                  t2 += a.y;
                }
                if (tc.b0 + bi < p.B) {
                  double* sp = p.stats + (static_cast<long long>(tc.b0 + bi) * p.stats_ld + ncol) * 2;
                  atomicAdd(sp, static_cast<double>(t1));
                  atomicAdd(sp + 1, static_cast<double>(t2));
                }
              }
            }
          }
        }
        if (tre) p.trace[320 + (c >> 6) * 8 + 5] = clock64();
      }
      tc_fence_before();
      if constexpr (PAIR) mbar_arrive_cluster(&tempty_bar[as], 0); else mbar_arrive(&tempty_bar[as]);
      if (tracing && it < 8 && et == 0) p.trace[6 * 16 + it] = clock64();
    }
    if (!sw_mode && use_tma && (warp_mode ? lane == 0 : gt == 0)) bulk_wait_group_all();      // shared memory must outlive the last TMA stores
  } else if (LEAN) {
    // =============================== TMA-store issuer of epilogue group g (lean kernel) ===============================
    // walks the group's sub-block sequence (same tile order, same column order), one 128-row box per sub-block
    const int g = warp - kStoreWarp0;
    uint32_t nsub = 0;
    int it = 0;
    for (int t = u_first; t < total_tiles; t += u_stride, ++it) {
      const TileCoord tc = tile_coord(p, t, n_tiles, BN, Wt, Ht, Bt, rank);
      for (int c = ((g + it) & 1) * 32; c < BN; c += 64, ++nsub) {
        const int b = nsub & 1;
        mbar_wait(&store_req[g * 2 + b], (nsub >> 1) & 1);
        if (lane == 0) {
          tma_store_4d(&mapOut, staging + g * kEpiStageBytes + b * kEpiBufBytes, tc.n0 + c, tc.x0, tc.y0, tc.b0);
          bulk_commit_group();
          bulk_wait_group_read<0>();
          mbar_arrive(&store_done[g * 2 + b]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) bulk_wait_group_all();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();               // the peer's shared memory / TMEM stay alive until both are done
  if (warp == kMmaWarp) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// out[b, y, x, n] = bf16(sum_ks ws[ks][m, n] + bias[n] + rowvec[b, n] + residual[b, y, x, n]), slabs added in ks order.
// grid (chunks, B), block CV * PL threads: thread (cv, pl) owns 8 channels and pixels p0+pl, p0+pl+PL, ... of image b;
// with p.stats the per-(image, channel) sum / sum of squares of the bf16 output (GroupNorm statistics of the consumer)
// are reduced like ur_chan_stats (per-lane partial rows in shared memory, fixed order) and added in fp64.
__global__ void splitk_finish_kernel(const GemmParams p, int CV, int PL, int chunk) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float sh[];   // [PL][2][N] (statistics only)
  const int b = blockIdx.y;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int P = p.Ho * p.Wo;
  const int p0 = blockIdx.x * chunk, p1 = min(P, p0 + chunk);
  bf16* outp = reinterpret_cast<bf16*>(p.out);
  float add[8], s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = cv * 8 + j;
    add[j] = (p.bias ? __ldg(p.bias + n) : 0.f) + (p.rowvec ? __ldg(p.rowvec + b * p.rowvec_sb + n) : 0.f);
    s[j] = q[j] = 0.f;
  }
  for (int pix = p0 + pl; pix < p1; pix += PL) {
    const long long m = static_cast<long long>(b) * P + pix;
    const int y = pix / p.Wo, x = pix - y * p.Wo;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    const float* wp = p.ws + m * p.N + cv * 8;
    for (int ks = 0; ks < p.ksplit; ++ks, wp += p.ws_slab) {
      const float4 a0 = *reinterpret_cast<const float4*>(wp);
      const float4 a1 = *reinterpret_cast<const float4*>(wp + 4);
      f[0] += a0.x; f[1] += a0.y; f[2] += a0.z; f[3] += a0.w;
      f[4] += a1.x; f[5] += a1.y; f[6] += a1.z; f[7] += a1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += add[j];
    if (p.residual) {
      const uint4 rv = __ldg(reinterpret_cast<const uint4*>(p.residual + b * p.res_sb + y * p.res_sy + x * p.res_sx + cv * 8));
      const uint32_t u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float r0, r1;
        unpack_bf16(u[k], r0, r1);
        f[2 * k] += r0;
        f[2 * k + 1] += r1;
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[k] = pack_bf16(f[2 * k], f[2 * k + 1]);
      float r0, r1;
      unpack_bf16(o[k], r0, r1);                 // statistics of the ROUNDED output (what the consumer reads)
      s[2 * k] += r0;
      q[2 * k] = fmaf(r0, r0, q[2 * k]);
      s[2 * k + 1] += r1;
      q[2 * k + 1] = fmaf(r1, r1, q[2 * k + 1]);
    }
    *reinterpret_cast<uint4*>(outp + b * p.out_sb + y * p.out_sy + x * p.out_sx + cv * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (p.stats) {
    const int C = p.N;
    float* row = sh + static_cast<size_t>(pl) * 2 * C + cv * 8;
    *reinterpret_cast<float4*>(row) = make_float4(s[0], s[1], s[2], s[3]);
    *reinterpret_cast<float4*>(row + 4) = make_float4(s[4], s[5], s[6], s[7]);
    *reinterpret_cast<float4*>(row + C) = make_float4(q[0], q[1], q[2], q[3]);
    *reinterpret_cast<float4*>(row + C + 4) = make_float4(q[4], q[5], q[6], q[7]);
    __syncthreads();
    double* so = p.stats + static_cast<long long>(b) * p.stats_ld * 2;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      float t = 0.f;
      for (int l = 0; l < PL; ++l) t += sh[static_cast<size_t>(l) * 2 * C + i];
      const int c = i < C ? i : i - C;
      atomicAdd(so + 2 * c + (i < C ? 0 : 1), static_cast<double>(t));
    }
  }
}

int launch_splitk_finish(const GemmParams& p, cudaStream_t stream) {
  const int CV = p.N >> 3;
  if (CV > 1024) return set_error(UR_ERR_ARG, "splitk_finish: N too large");
  const int PL = CV >= 256 ? 1 : 256 / CV;
  const int P = p.Ho * p.Wo;
  const int target = max(1, (2 * num_sms()) / (p.B > 0 ? p.B : 1));
  int chunk = (P + target - 1) / target;
  if (chunk < 2 * PL) chunk = 2 * PL;
  const size_t smem = p.stats ? 2 * static_cast<size_t>(p.N) * PL * sizeof(float) : 0;
  if (smem > 48 * 1024) return set_error(UR_ERR_ARG, "splitk_finish: statistics need too much shared memory");
  cudaError_t e = launch_kernel(splitk_finish_kernel, dim3((P + chunk - 1) / chunk, p.B), dim3(CV * PL), smem, stream, p,
                                CV, PL, chunk);
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "splitk_finish launch");
}

template <int BN, bool PAIR, bool LEAN>
static int launch_p(const GemmParams& p, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w,
                    const CUtensorMap& mo, int total_units, int n_tiles, cudaStream_t stream) {
  using Cfg = PCfg<BN, PAIR>;
  constexpr int smem = Cfg::kSmem;
  static_assert(smem <= 227 * 1024 && Cfg::kStages >= 3, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_persistent_kernel<BN, PAIR, LEAN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_gemm_persistent)");
    configured = true;
  }
  const int slots = PAIR ? num_sms() / 2 : num_sms();
  const int units = total_units < slots ? total_units : slots;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PAIR ? 2 * units : units);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_gemm_persistent_kernel<BN, PAIR, LEAN>, p, a1, a2, w, mo, total_units, n_tiles);
  return e == cudaSuccess ? UR_OK : set_cuda_error(e, "conv_gemm_persistent launch");
}


int launch_conv_gemm_persistent(const GemmParams& p, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w,
                                const CUtensorMap& mo, bool pair, int bn, int total_units, int n_tiles,
                                cudaStream_t stream) {
  const bool lean = p.tma_store == 4;          // chosen by the host together with the 128-row output box (ur_gemm.cu)
  if (pair) {
    switch (bn) {
      case 64: return lean ? launch_p<64, true, true>(p, a1, a2, w, mo, total_units, n_tiles, stream)
                  : launch_p<64, true, false>(p, a1, a2, w, mo, total_units, n_tiles, stream);
      case 128: return lean ? launch_p<128, true, true>(p, a1, a2, w, mo, total_units, n_tiles, stream)
                  : launch_p<128, true, false>(p, a1, a2, w, mo, total_units, n_tiles, stream);
      case 160: return lean ? launch_p<160, true, true>(p, a1, a2, w, mo, total_units, n_tiles, stream)
                  : launch_p<160, true, false>(p, a1, a2, w, mo, total_units, n_tiles, stream);
      default: return lean ? launch_p<256, true, true>(p, a1, a2, w, mo, total_units, n_tiles, stream)
                  : launch_p<256, true, false>(p, a1, a2, w, mo, total_units, n_tiles, stream);
    }
  }
  switch (bn) {
    case 64: return lean ? launch_p<64, false, true>(p, a1, a2, w, mo, total_units, n_tiles, stream)
                  : launch_p<64, false, false>(p, a1, a2, w, mo, total_units, n_tiles, stream);
    case 128: return lean ? launch_p<128, false, true>(p, a1, a2, w, mo, total_units, n_tiles, stream)
                  : launch_p<128, false, false>(p, a1, a2, w, mo, total_units, n_tiles, stream);
    case 160: return lean ? launch_p<160, false, true>(p, a1, a2, w, mo, total_units, n_tiles, stream)
                  : launch_p<160, false, false>(p, a1, a2, w, mo, total_units, n_tiles, stream);
    default: return lean ? launch_p<256, false, true>(p, a1, a2, w, mo, total_units, n_tiles, stream)
                  : launch_p<256, false, false>(p, a1, a2, w, mo, total_units, n_tiles, stream);
  }
}

}  // namespace ur
