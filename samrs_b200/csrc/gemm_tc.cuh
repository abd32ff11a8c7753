// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )       A, B fp16 (K-major), fp32 accumulate in TMEM
//
// Covers every nn.Linear / 1x1 conv / im2col'ed conv of the image encoder
// (reference call sites: SA/modeling/image_encoder.py:227,238 qkv/proj, SA/modeling/common.py:25-26 MLP,
// image_encoder.py:391-395 patch embed, :88-104 neck).  B is the torch weight [out_features, in_features]
// as stored in the state dict, which is already K-major.
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (one accumulator row per thread).
// Pipelines: smem ring full/empty (TMA <-> MMA), two TMEM accumulators full/empty (MMA <-> epilogue),
// static persistent tile schedule (tile = blockIdx.x + i * gridDim.x).
#pragma once
#include "common.cuh"

namespace samrs {

struct GemmParams {
  int M, N, K;
  void* out;            // half or float, row-major, leading dimension ldc (elements)
  int ldc;
  const float* bias;    // [N] or nullptr
  const float* res;     // fp32 [res_mod or M, ldr] added in the epilogue, or nullptr
  int ldr;
  int res_mod;          // residual row = m % res_mod (0 -> m)
  int tiles_m, tiles_n;
  int batch;            // >= 1: independent problems sharing B; A is then a rank-3 tensor map (k, batch, m)
  int a_rank3;
  long long out_batch_stride;   // elements between consecutive batch outputs
  float out_scale;      // accumulator is multiplied by this before bias/residual (split-weight scaling); 0 -> 1
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;   // one 128-byte swizzle atom of fp16

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytes = GEMM_BM * 128 + BN * 128;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 160 ? 5 : 6);
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Epilogue of one accumulator tile for the thread that owns TMEM lane (row): BN fp32 columns -> scale, bias,
// residual, activation -> fp16 / fp32 global stores (each thread writes contiguous 64- or 128-byte runs of its row).
template <int BN, bool OUT_HALF, int ACT>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, uint32_t t_addr, int row, int n0, int bt, float oscale) {
      const float* res_row = nullptr;
        if (p.res != nullptr && row < p.M)
          res_row = p.res + size_t(p.res_mod > 0 ? row % p.res_mod : row) * p.ldr;
  #pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(t_addr + uint32_t(c * 32), v);
          tc_wait_ld();
          const int col0 = n0 + c * 32;
          if (row < p.M && col0 < p.N) {
            float f[32];
  #pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * oscale;
            const bool fullchunk = (col0 + 32 <= p.N);
            if (p.bias != nullptr) {
              if (fullchunk) {
  #pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                  f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) f[j] += __ldg(p.bias + col0 + j);
              }
            }
            if (res_row != nullptr) {
              if (fullchunk) {
  #pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 r = *reinterpret_cast<const float4*>(res_row + col0 + j);
                  f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) f[j] += res_row[col0 + j];
              }
            }
            if (ACT == 1) {
  #pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
            }
            if (OUT_HALF) {
              __half* o = reinterpret_cast<__half*>(p.out) + size_t(bt) * p.out_batch_stride + size_t(row) * p.ldc + col0;
              if (fullchunk) {
  #pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 pk;
                  __half2 h0 = __floats2half2_rn(f[j], f[j + 1]);
                  __half2 h1 = __floats2half2_rn(f[j + 2], f[j + 3]);
                  __half2 h2 = __floats2half2_rn(f[j + 4], f[j + 5]);
                  __half2 h3 = __floats2half2_rn(f[j + 6], f[j + 7]);
                  pk.x = *reinterpret_cast<uint32_t*>(&h0);
                  pk.y = *reinterpret_cast<uint32_t*>(&h1);
                  pk.z = *reinterpret_cast<uint32_t*>(&h2);
                  pk.w = *reinterpret_cast<uint32_t*>(&h3);
                  *reinterpret_cast<uint4*>(o + j) = pk;
                }
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) o[j] = __float2half_rn(f[j]);
              }
            } else {
              float* o = reinterpret_cast<float*>(p.out) + size_t(bt) * p.out_batch_stride + size_t(row) * p.ldc + col0;
              if (fullchunk) {
  #pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) o[j] = f[j];
              }
            }
          }
        }
}

template <int BN, bool OUT_HALF, int ACT>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int num_tiles = tiles_mn * p.batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 4);
    mbar_init(&tempty[1], 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int bt = t / tiles_mn, tt = t % tiles_mn;
        const int m0 = (tt % p.tiles_m) * GEMM_BM;
        const int n0 = (tt / p.tiles_m) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + GEMM_BM * 128;
          mbar_expect_tx(&full[stage], Cfg::kStageBytes);
          if (p.a_rank3) tma_load_3d(sa, &tmA, &full[stage], kb * GEMM_BK, bt, m0);
          else tma_load_2d(sa, &tmA, &full[stage], kb * GEMM_BK, m0);
          tma_load_2d(sb, &tmB, &full[stage], kb * GEMM_BK, n0);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
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
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 fp16 (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
            tc_mma_f16(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0);
          }
          tc_commit(&empty[stage]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int q = warp & 3;                       // TMEM lane quadrant of this warp
    const float oscale = (p.out_scale != 0.f) ? p.out_scale : 1.0f;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int bt = t / tiles_mn, tt = t % tiles_mn;
      const int m0 = (tt % p.tiles_m) * GEMM_BM;
      const int n0 = (tt / p.tiles_m) * BN;
      const int row = m0 + q * 32 + lane;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      gemm_epilogue_tile<BN, OUT_HALF, ACT>(p, t_addr, row, n0, bt, oscale);
      // accumulator drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// Host-side launcher: picks the N tile that wastes the fewest SM-waves for this shape.
int launch_gemm_tc(const __half* A, int lda, const __half* B, int ldb, const GemmParams& p, bool out_half, int act,
                   int num_sms, cudaStream_t stream, int force_bn = 0, const CUtensorMap* a_map_rank3 = nullptr);

}  // namespace samrs
