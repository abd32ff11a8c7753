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

template <int BN>
struct Gemm2Cfg {
  static constexpr int kStageBytes = GEMM_BM * 128 + (BN / 2) * 128;     // per CTA
  static constexpr int kStages = (226 * 1024 - 1280 - GEMM_EPI_SMEM) / kStageBytes > 8 ? 8 : (226 * 1024 - 1280 - GEMM_EPI_SMEM) / kStageBytes;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + GEMM_EPI_SMEM;
};

template <int BN, bool OUT_HALF, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC16 /*16-column tail box (BN % 32 == 16)*/,
                const GemmParams p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 0) { gemm_dbg(p, 0); gemm_dbg_wall(p, false); }
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
    mbar_init(&tempty[0], 2 * GEMM_EPI_WARPS);
    mbar_init(&tempty[1], 2 * GEMM_EPI_WARPS);
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
            tma_load_2d_pair(sa, &tmA, lead_full, kb * GEMM_BK, m0);
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
    uint8_t* my_stage = epi_stage + (warp - 4) * GEMM_EPI_WARP_SMEM;
    uint32_t nstaged = 0;
    constexpr int NCH = (BN + 31) / 32, SPLIT = OUT_HALF ? ((NCH + 1) / 4) * 2 : (NCH + 1) / 2;   // fp16: whole chunk pairs per group
    const int cb = (warp < 8) ? 0 : SPLIT, ce = (warp < 8) ? SPLIT : NCH;
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

}  // namespace samrs
