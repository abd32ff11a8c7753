// CTA-pair (cta_group::2) variant of the tcgen05 GEMM: two SMs of one TPC cooperate on a 256 x BN output tile.
//
// Why: with one CTA per 128 x BN tile every SM pulls (128 + BN) x 64 fp16 from L2 per 128 x BN x 64 MMA block; once the
// issue loop was tight (profiles/r01_gemm_trace_v5.txt) that feed bounded the 1-CTA kernel.  In a pair each CTA loads its
// own 128 rows of A and only HALF of the B tile; the MMA (M = 256, issued by the leader CTA) reads both halves of B from
// the two CTAs' shared memory, so the B bytes per flop halve and the main loop runs at the MMA floor (335-340 clk per
// 64-deep k-block of a 160-wide tile, floor 320).  Accumulators stay per-CTA: each CTA's TMEM holds its 128 rows x BN columns and each CTA runs
// its own epilogue warps.
//
// Protocol (leader = cluster rank 0):
//   full[s]   : leader's barrier, armed by the leader's producer with the bytes of BOTH CTAs; both producers' TMA
//               loads complete_tx on it (the peer addresses it through mapa).
//   empty[s]  : one per CTA; tcgen05.commit.multicast from the leader releases the slot in both CTAs.
//   tfull[a]  : one per CTA (multicast commit); tempty[a]: leader's, 16 arrivals (8 epilogue warps x 2 CTAs).
#pragma once
#include "gemm_tc.cuh"

namespace samrs {

// EW = epilogue warps per CTA: 8 (two column groups, 8 KiB of staging each) or 12 (three groups, 4 KiB each, fp16 output)
template <int BN, int EW = GEMM_EPI_WARPS>
struct Gemm2Cfg {
  static constexpr int kStageBytes = GEMM_BM * 128 + (BN / 2) * 128;     // per CTA
  static constexpr int kEpiWarpSmem = EW > 8 ? 4096 : GEMM_EPI_WARP_SMEM;
  static constexpr int kEpiSmem = EW * kEpiWarpSmem;
  static constexpr int kStages = (226 * 1024 - 1280 - kEpiSmem) / kStageBytes > 8 ? 8 : (226 * 1024 - 1280 - kEpiSmem) / kStageBytes;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kEpiSmem;
  static constexpr int kThreads = 128 + 32 * EW;
};

template <int BN, bool OUT_HALF, int ACT, int EW = GEMM_EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * EW, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC16 /*16-column tail box (BN % 32 == 16)*/,
                const GemmParams p) {
  using Cfg = Gemm2Cfg<BN, EW>;
  static_assert(EW == 8 || (EW == 12 && OUT_HALF), "twelve epilogue warps exist for the fp16 epilogue");
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 0) { gemm_dbg(p, 0); gemm_dbg_wall(p, false); }
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + S * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + Cfg::kEpiSmem);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = p.tiles_m * p.tiles_n;      // tiles of 256 x BN

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 2 * EW);
    mbar_init(&tempty[1], 2 * EW);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();                               // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) gemm_dbg(p, 1);
  pdl_trigger();
  pdl_wait();                                       // the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (both CTAs)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int m0 = (t % p.tiles_m) * 256 + int(rank) * GEMM_BM;
        const int n0 = (t / p.tiles_m) * BN + int(rank) * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + GEMM_BM * 128;
          const uint32_t lead_full = mapa_u32(&full[stage], 0);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
            tma_load_2d_pair(sa, &tmA, lead_full, ((p.a_wrap_kb > 0 && kb >= p.a_wrap_kb) ? kb - p.a_wrap_kb : kb) * GEMM_BK, m0);
            tma_load_2d_pair(sb, &tmB, lead_full, kb * GEMM_BK, n0);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(256, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0, ti = 0;
      uint32_t aphase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        gemm_dbg(p, 16 + ti * 64 + 0);
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        gemm_dbg(p, 16 + ti * 64 + 1);
        const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (kGemmExp && kb < 40) gemm_dbg(p, 16 + ti * 64 + 2 + kb);
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + GEMM_BM * 128;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sb);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              tc_mma_f16_pair(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0);
            tc_commit_pair(&empty[stage], 0x3);
            if (kb == num_kb - 1) tc_commit_pair(&tfull[as], 0x3);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        gemm_dbg(p, 16 + ti * 64 + 60);
        ++ti;
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;
    const float oscale = (p.out_scale != 0.f) ? p.out_scale : 1.0f;
    uint8_t* my_stage = epi_stage + (warp - 4) * Cfg::kEpiWarpSmem;
    uint32_t nstaged = 0;
    constexpr int NCH = (BN + 31) / 32, SPLIT = OUT_HALF ? ((NCH + 1) / 4) * 2 : (NCH + 1) / 2;   // fp16: whole chunk pairs per group
    int cb = (warp < 8) ? 0 : SPLIT, ce = (warp < 8) ? SPLIT : NCH;
    if (EW == 12) {                                    // three column groups, the larger shares first
      const int g = (warp - 4) >> 2;
      cb = g * (NCH / 3) + (g < NCH % 3 ? g : NCH % 3);
      ce = cb + NCH / 3 + (g < NCH % 3 ? 1 : 0);
    }
    int as = 0, eti = 0;
    uint32_t aphase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int m0 = (t % p.tiles_m) * 256 + int(rank) * GEMM_BM;
      const int n0 = (t / p.tiles_m) * BN;
      float bias_r[(BN + 31) / 32];
      gemm_load_bias<BN>(p, n0, lane, bias_r);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if (warp == 4 && lane == 0) gemm_dbg(p, 16 + eti * 64 + 61);
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      if constexpr (EW == 12)
        gemm_epilogue_warp_half1<BN, ACT>(p, &tmC, t_addr, m0 + q * 32, n0, oscale, my_stage, nstaged, lane, bias_r, cb, ce);
      else
        gemm_epilogue_warp<BN, OUT_HALF, ACT>(p, &tmC, &tmC16, t_addr, m0 + q * 32, n0, 0, oscale, my_stage, nstaged, lane, bias_r, cb, ce);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty[as], 0));
      if (warp == 4 && lane == 0) gemm_dbg(p, 16 + eti * 64 + 62);
      ++eti;
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (elect_one()) tma_store_wait_all();
  }

  tc_fence_before();
  if (threadIdx.x == 0) gemm_dbg(p, 2);
  cluster_sync_all();                               // nobody exits (or frees TMEM) while the pair still signals
  if (threadIdx.x == 0) { gemm_dbg(p, 3); gemm_dbg_wall(p, true); }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}


#ifdef SAMRS_EXPERIMENTS
// Two schedules that were built, verified and measured, and did not beat the tile schedule above;
// they are compiled into libsamrs_b200_exp.so only (build.sh exp) and exercised by tests/test_gpu_kernels.py when that build is loaded.
// ---------------------------------------------------------------------------------------------------------------------------
// Stream-K schedule of the same CTA-pair pipeline, for the GEMMs that add into the fp32 residual stream (proj / lin2:
// out += A W^T + b through TMA reduce-add stores).
//
// Why: N = 1280 gives 16 x 8 = 128 tiles of 256 x 160 on 74 CTA pairs = 1.73 waves; 20 pairs idle through the second wave and
// the tile cannot widen (a 256-wide tile, which feeds the tensor pipe better, would be 80 tiles = 1.08 waves).  Here the work is
// the flat list of (tile, k-block) units, cut into 74 equal contiguous ranges; a pair's range covers whole tiles plus at most a
// HEAD piece (k-blocks 0..j of the tile where its range ends) and a TAIL piece (k-blocks j..K of the tile where it starts).
// Every piece is an ordinary pass of the pipeline whose epilogue reduce-adds its partial sums into `out`; the bias rides on
// the piece that holds k-block 0.
//
// Deterministic sums: a split tile receives exactly two reduce-adds, and their order is fixed.  A pair walks its range
// from the END, so the HEAD piece is the first thing pair p-1 does and the TAIL piece of the same tile the last thing pair p
// does; the TAIL epilogue still waits (acquire) on a per-tile counter that the 16 epilogue warps of the HEAD piece bump
// (release) once their bulk stores have completed, so `out + head + tail` is the order on every run.  A waiter only ever waits
// for a lower-numbered cluster that does not wait before signalling, so the scheme needs nothing beyond the in-order block
// scheduling that decoupled look-back scans rely on.  The last TAIL warp to pass resets the tile's two counters for the next launch.
// Requires units-per-pair >= k-blocks-per-tile (no tile is cut three ways); the host falls back to the tile schedule otherwise.
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_sk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int S = Cfg::kStages;
  static_assert(BN % 32 == 0, "stream-K tiles are whole 32-column chunks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + S * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + GEMM_EPI_SMEM);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  // this pair's contiguous range of (tile, k-block) units
  const int total = p.tiles_m * p.tiles_n * num_kb;
  const int per = total / num_pairs, extra = total % num_pairs;
  const int u_begin = pair * per + (pair < extra ? pair : extra);
  const int u_end = u_begin + per + (pair < extra ? 1 : 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 2 * GEMM_EPI_WARPS);
    mbar_init(&tempty[1], 2 * GEMM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  // pieces, last one first: (tile, kb0, kb1) with cur = end of the piece in unit space
#define SAMRS_SK_FOR_EACH_PIECE                                         \
  for (int cur = u_end, t, kb0, kb1; cur > u_begin; cur = t * num_kb + kb0) \
    if (t = (cur - 1) / num_kb, kb1 = cur - t * num_kb, kb0 = (u_begin > t * num_kb ? u_begin - t * num_kb : 0), true)

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    SAMRS_SK_FOR_EACH_PIECE {
      const int m0 = (t % p.tiles_m) * 256 + int(rank) * GEMM_BM;
      const int n0 = (t / p.tiles_m) * BN + int(rank) * (BN / 2);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + GEMM_BM * 128;
        const uint32_t lead_full = mapa_u32(&full[stage], 0);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
          tma_load_2d_pair(sa, &tmA, lead_full, kb * GEMM_BK, m0);
          tma_load_2d_pair(sb, &tmB, lead_full, kb * GEMM_BK, n0);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(256, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      SAMRS_SK_FOR_EACH_PIECE {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + GEMM_BM * 128;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sb);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              tc_mma_f16_pair(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, ((kb - kb0) | k) != 0);
            tc_commit_pair(&empty[stage], 0x3);
            if (kb == kb1 - 1) tc_commit_pair(&tfull[as], 0x3);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const float oscale = (p.out_scale != 0.f) ? p.out_scale : 1.0f;
    uint8_t* my_stage = epi_stage + (warp - 4) * GEMM_EPI_WARP_SMEM;
    uint32_t nstaged = 0;
    constexpr int NCH = BN / 32, SPLIT = (NCH + 1) / 2;
    const int cb = (warp < 8) ? 0 : SPLIT, ce = (warp < 8) ? SPLIT : NCH;
    int as = 0;
    uint32_t aphase = 0;
    SAMRS_SK_FOR_EACH_PIECE {
      const int m0 = (t % p.tiles_m) * 256 + int(rank) * GEMM_BM;
      const int n0 = (t / p.tiles_m) * BN;
      float bias_r[NCH];
      if (kb0 == 0) {
        gemm_load_bias<BN>(p, n0, lane, bias_r);
      } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c) bias_r[c] = 0.f;
      }
      int* cnt = p.sk_flags + 2 * t;                // [0] HEAD warps done, [1] TAIL warps that have seen it
      if (kb0 > 0) {
        // TAIL piece: the HEAD piece of this tile (another pair's first piece) must have landed in `out` before ours is added
        if (lane == 0) {
          while (ld_acquire_gpu(cnt) < 2 * GEMM_EPI_WARPS) __nanosleep(64);
          if (atomicAdd(cnt + 1, 1) == 2 * GEMM_EPI_WARPS - 1) { cnt[0] = 0; cnt[1] = 0; }     // everyone has passed: re-arm for the next launch
        }
        __syncwarp();
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      gemm_epilogue_warp<BN, false, 0>(p, &tmC, &tmC, t_addr, m0 + q * 32, n0, 0, oscale, my_stage, nstaged, lane, bias_r, cb, ce);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty[as], 0));
      if (kb0 == 0 && kb1 < num_kb) {
        // HEAD piece: publish it once this warp's reduce-adds have been performed (the committing lane waits for its groups)
        if (elect_one()) {
          tma_store_wait_all();
          __threadfence();
          red_release_gpu_add(cnt, 1);
        }
        __syncwarp();
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (elect_one()) tma_store_wait_all();
  }
#undef SAMRS_SK_FOR_EACH_PIECE

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Two CTA pairs per cluster (cluster of 4), stacked along M, sharing the B tile through TMA multicast.
//
// Why: back to back the pair kernel moves 11-12 TB/s from the L2 into shared memory whatever the schedule (tile waves or
// stream-K), i.e. it sits on the L2 -> SM fabric, not on the tensor pipe.  Here the 512 x BN super-tile of a cluster needs the
// B tile once: rank r = 2 * pair + half keeps B rows [half * BN/2, +BN/2) like before, but loads only ONE quarter of the tile
// (rows half * BN/2 + pair * BN/4 ...) and multicasts it to the CTA of the same half in both pairs.  Per CTA and k-block:
// 16 KB of A + BN/4 rows of B (23 KB instead of 30 KB at BN = 224).
//
// Protocol differences from gemm_tc2_kernel: a stage slot is written by this CTA and by its partner (rank ^ 2), so
// empty[s] counts the commits of BOTH pairs' MMAs (multicast to all four CTAs); the two pairs therefore run their k-loops in
// lock-step.  full[s] stays per pair (leader = even rank): every load that lands in a pair's shared memory credits that
// pair's leader, whoever issued it.
template <int BN, bool OUT_HALF, int ACT>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB /*box: BN/4 rows*/,
                const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int S = Cfg::kStages;
  static_assert(BN % 32 == 0, "quarter B boxes must be whole 8-row swizzle atoms");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + S * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + GEMM_EPI_SMEM);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0..3
  const uint32_t half = rank & 1, pr = rank >> 1, lead_rank = rank & ~1u;
  const int cl = blockIdx.x >> 2, num_cl = gridDim.x >> 2;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int tiles_m2 = (p.tiles_m + 1) >> 1;        // super-tiles of 512 rows
  const int num_super = tiles_m2 * p.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 2);                      // the MMA commits of both pairs
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 2 * GEMM_EPI_WARPS);
    mbar_init(&tempty[1], 2 * GEMM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    const uint16_t bmask = uint16_t((1u << half) | (1u << (half + 2)));
    for (int t = cl; t < num_super; t += num_cl) {
      const int m0 = ((t % tiles_m2) * 2 + int(pr)) * 256 + int(half) * GEMM_BM;
      const int n0 = (t / tiles_m2) * BN + int(half) * (BN / 2) + int(pr) * (BN / 4);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + GEMM_BM * 128 + int(pr) * (BN / 4) * 128;
        const uint32_t lead_full = mapa_u32(&full[stage], lead_rank);
        if (elect_one()) {
          if (half == 0) mbar_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
          tma_load_2d_pair(sa, &tmA, lead_full, kb * GEMM_BK, m0);
          tma_load_2d_pair_mcast(sb, &tmB, &full[stage], kb * GEMM_BK, n0, bmask);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (half == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(256, BN, 0, 0);
      const uint16_t pair_mask = uint16_t(0x3u << (2 * pr));
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cl; t < num_super; t += num_cl) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + GEMM_BM * 128;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sb);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              tc_mma_f16_pair(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0);
            tc_commit_pair(&empty[stage], 0xF);
            if (kb == num_kb - 1) tc_commit_pair(&tfull[as], pair_mask);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const float oscale = (p.out_scale != 0.f) ? p.out_scale : 1.0f;
    uint8_t* my_stage = epi_stage + (warp - 4) * GEMM_EPI_WARP_SMEM;
    uint32_t nstaged = 0;
    constexpr int NCH = BN / 32, SPLIT = OUT_HALF ? ((NCH + 1) / 4) * 2 : (NCH + 1) / 2;
    const int cb = (warp < 8) ? 0 : SPLIT, ce = (warp < 8) ? SPLIT : NCH;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = cl; t < num_super; t += num_cl) {
      const int m0 = ((t % tiles_m2) * 2 + int(pr)) * 256 + int(half) * GEMM_BM;
      const int n0 = (t / tiles_m2) * BN;
      float bias_r[NCH];
      gemm_load_bias<BN>(p, n0, lane, bias_r);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      gemm_epilogue_warp<BN, OUT_HALF, ACT>(p, &tmC, &tmC, t_addr, m0 + q * 32, n0, 0, oscale, my_stage, nstaged, lane, bias_r, cb, ce);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty[as], lead_rank));
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (elect_one()) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();                               // nobody exits (or frees TMEM) while the cluster still signals
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}
#endif  // SAMRS_EXPERIMENTS

}  // namespace samrs
