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
  unsigned long long* dbg;   // optional: CTA 0 records clock64() at pipeline events (tools/gemm_trace.py)
  int accumulate;       // fp32 output only: out += result (TMA reduce-add) instead of out = result
  int dbg_mode;         // tools only (results are garbage): bit0 skip TMA loads, bit1 skip MMAs, bit2 skip the epilogue
  // ACT == 3 (mask decoder, second up-scaling stage fused with the hyper-network product; see gemm_epilogue_warp_upscale2):
  const float* hyper = nullptr;   // [B][hyper_nm][32]
  float* low = nullptr;           // [B][hyper_nm][256][256] low-res mask logits
  int hyper_nm = 0;
  // 3-term split-fp16 GEMMs (mask decoder): A is stored as [hi | lo] (2K columns) and its k-blocks >= a_wrap_kb are read again
  // from the start, i.e. the operand behaves as [hi | lo | hi] without the third copy ever being written or fetched from HBM
  int a_wrap_kb = 0;
  // stream-K schedule (gemm_tc2_sk_kernel): two zero-initialised counters per output tile, owned by the engine
  int* sk_flags = nullptr;
};
// Pipeline traces and the "remove one stage" experiments of tools/gemm_trace.py exist only in builds with
// -DSAMRS_EXPERIMENTS (build.sh exp -> libsamrs_b200_exp.so); the product library compiles them away.
#ifdef SAMRS_EXPERIMENTS
constexpr bool kGemmExp = true;
__device__ __forceinline__ void gemm_dbg(const GemmParams& p, int slot) {
  if (p.dbg != nullptr && blockIdx.x == 0 && slot < 4096) p.dbg[slot] = clock64();
}
// wall-clock span of the whole grid (slots 10 / 11 = min start / max end over CTAs, ns) and of CTA 0 (12 / 13)
__device__ __forceinline__ void gemm_dbg_wall(const GemmParams& p, bool end) {
  if (p.dbg == nullptr) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (end) atomicMax(p.dbg + 11, t); else atomicMin(p.dbg + 10, t);
  if (blockIdx.x == 0) p.dbg[end ? 13 : 12] = t;
}
#else
constexpr bool kGemmExp = false;
__device__ __forceinline__ void gemm_dbg(const GemmParams&, int) {}
__device__ __forceinline__ void gemm_dbg_wall(const GemmParams&, bool) {}
#endif

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;   // one 128-byte swizzle atom of fp16
constexpr int GEMM_EPI_WARPS = 8;                    // warps 4..11: two column groups of four (TMEM lane quadrant = warp % 4)
constexpr int GEMM_THREADS = 128 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_EPI_WARP_SMEM = 8192;                  // per epilogue warp: two 32x32 fp32 or four 32x32 fp16 staging blocks
constexpr int GEMM_EPI_SMEM = GEMM_EPI_WARPS * GEMM_EPI_WARP_SMEM;

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytes = GEMM_BM * 128 + BN * 128;
  // as many stages as fit: bytes in flight per SM, not tile shape, set the achievable L2->SM feed rate
  static constexpr int kStages = (226 * 1024 - 1280 - GEMM_EPI_SMEM) / kStageBytes > 8 ? 8 : (226 * 1024 - 1280 - GEMM_EPI_SMEM) / kStageBytes;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + GEMM_EPI_SMEM;
};

// GELU(x) = x/2 (1 + erf(x / sqrt 2)) = (h + |h|) - |h| erfc(z) with h = x/2, z = |x| / sqrt 2, and
// erfc(z) ~= 2^(-z Q(z)) on [0, 4] (Q = degree-5 least-squares fit of -log2(erfc z) / z, tools/fit_gelu.py): |error| < 3e-7,
// far below the fp16 rounding of the result.  One MUFU op and ten FP32 ops per element: the epilogue of lin1 is bound
// by MUFU / issue slots, erff costs ~30 instructions and the Abramowitz-Stegun 7.1.26 form used before two MUFU ops.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float q = fmaf(z, -2.635702863e-04f, 4.330650429e-03f);
  q = fmaf(q, z, -3.223223820e-02f);
  q = fmaf(q, z, 1.509066050e-01f);
  q = fmaf(q, z, 9.176831254e-01f);
  q = fmaf(q, z, 1.627991484e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * q));               // MUFU, ~2 ulp
  const float h = 0.5f * x;
  return fmaf(-fabsf(h), e, h + fabsf(h));
}

// fp16-output epilogue (qkv, lin1): 32-column chunks are processed in pairs - both tcgen05.ld are in flight together, the
// two converted 32x32 fp16 blocks are staged side by side and handed to the TMA unit behind ONE proxy fence / elect -
// and the four 2 KiB staging blocks of the warp rotate, so the warp only waits for the store issued two pairs ago.
template <int BN, int ACT>
__device__ __forceinline__ void gemm_epilogue_warp_half(const GemmParams& p, const CUtensorMap* tmC, uint32_t t_addr, int row0, int n0,
                                                        int bt, float oscale, uint8_t* stage, uint32_t& nstaged, int lane,
                                                        const float (&bias_r)[(BN + 31) / 32], int chunk_begin, int chunk_end) {
  static_assert(BN % 32 == 0, "fp16-output tiles are whole 32-column chunks");
  constexpr int NCH = (BN + 31) / 32;
  auto convert = [&](const uint32_t (&v)[32], float bias_c, uint8_t* buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] = fmaf(__uint_as_float(v[8 * j + i]), oscale, __shfl_sync(0xffffffffu, bias_c, 8 * j + i));
        if (ACT == 1) f[i] = gelu_erf(f[i]);
      }
      __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
      __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(buf + lane * 64 + j * 16) = pk;
    }
  };
#pragma unroll
  for (int c = 0; c < NCH; c += 2) {
    if (c < chunk_begin || c >= chunk_end) continue;
    if (n0 + c * 32 >= p.N) break;
    const bool two = (c + 1 < NCH) && (c + 1 < chunk_end) && (n0 + (c + 1) * 32 < p.N);     // warp-uniform
    uint32_t v0[32], v1[32];
    tmem_ld32(t_addr + uint32_t(c * 32), v0);
    if (two) tmem_ld32(t_addr + uint32_t((c + 1 < NCH ? c + 1 : c) * 32), v1);
    uint8_t* buf = stage + (nstaged & 1) * (GEMM_EPI_WARP_SMEM / 2);
    ++nstaged;
    if (elect_one()) tma_store_wait_read<1>();   // the pair staged two steps ago has been read by the TMA unit
    __syncwarp();
    tc_wait_ld();
    convert(v0, bias_r[c], buf);
    if (two) convert(v1, bias_r[c + 1 < NCH ? c + 1 : c], buf + 2048);
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {                           // same lane every time (full-warp mask): bulk groups are per thread
      tma_store_3d(tmC, buf, n0 + c * 32, row0, bt);
      if (two) tma_store_3d(tmC, buf + 2048, n0 + (c + 1) * 32, row0, bt);
      tma_store_commit();
    }
  }
}

// fp16-output epilogue for the kernels that run twelve epilogue warps (lin1 + GELU: with eight warps the ~17 instructions per
// element of bias + erf-GELU + pack made the epilogue, not the main loop, set the tile period: 9.3 k clk against 8 k,
// profiles/r01_gemm_trace_v5.txt).  Three column groups of at most three chunks; a warp owns 4 KiB of staging = two 2 KiB
// blocks that alternate, one 32-column chunk per TMA store.
template <int BN, int ACT>
__device__ __forceinline__ void gemm_epilogue_warp_half1(const GemmParams& p, const CUtensorMap* tmC, uint32_t t_addr, int row0, int n0,
                                                         float oscale, uint8_t* stage /*4 KiB*/, uint32_t& nstaged, int lane,
                                                         const float (&bias_r)[(BN + 31) / 32], int chunk_begin, int chunk_end) {
  static_assert(BN % 32 == 0, "fp16-output tiles are whole 32-column chunks");
  constexpr int NCH = BN / 32;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c < chunk_begin || c >= chunk_end) continue;
    if (n0 + c * 32 >= p.N) break;
    uint32_t v[32];
    tmem_ld32(t_addr + uint32_t(c * 32), v);
    uint8_t* buf = stage + (nstaged & 1) * 2048;
    ++nstaged;
    if (elect_one()) tma_store_wait_read<1>();   // the block staged two chunks ago has been read by the TMA unit
    __syncwarp();
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] = fmaf(__uint_as_float(v[8 * j + i]), oscale, __shfl_sync(0xffffffffu, bias_r[c], 8 * j + i));
        if (ACT == 1) f[i] = gelu_erf(f[i]);
      }
      __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
      __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(buf + lane * 64 + j * 16) = pk;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      tma_store_3d(tmC, buf, n0 + c * 32, row0, 0);
      tma_store_commit();
    }
  }
}

// ACT == 3: the GEMM is ConvTranspose2d(64 -> 32, k2 s2) of the mask decoder's up-scaling (mask_decoder.py:58) over rows
// m = ((b * 4096 + token) * 4 + d1) (one 128 x 128 position each, 64 input channels) and columns n = d2 * 32 + c (the four
// output sub-positions x 32 channels).  The epilogue finishes the reference's chain in registers: + bias -> GELU (:59) ->
// mask[b, m', y, x] = sum_c hyper[b, m', c] * upscaled[b, c, y, x] (:163-167), so the (B, 32, 256, 256) tensor is never
// materialised: each 32-column chunk is one output pixel of the row's 2 x 2 block, written as one float per mask.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_warp_upscale2(const GemmParams& p, uint32_t t_addr, int row0, float oscale, int lane,
                                                            const float (&bias_r)[(BN + 31) / 32], int chunk_begin, int chunk_end) {
  static_assert(BN == 128, "the up-scaling epilogue expects the four 32-channel groups in one tile");
  const int m = row0 + lane;
  const bool ok = m < p.M;
  const int mm = ok ? m : 0;
  const int b = mm >> 14, rem = mm & 16383, token = rem >> 2, d1 = rem & 3;
  const int Y = 2 * (token >> 6) + (d1 >> 1), X = 2 * (token & 63) + (d1 & 1);      // position in the 128 x 128 grid
  const int bw = __shfl_sync(0xffffffffu, b, 0);                                   // 16384 rows per prompt: uniform in a tile
#pragma unroll
  for (int c = 0; c < (BN + 31) / 32; ++c) {
    if (c < chunk_begin || c >= chunk_end) continue;
    uint32_t v[32];
    tmem_ld32(t_addr + uint32_t(c * 32), v);
    tc_wait_ld();
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(fmaf(__uint_as_float(v[j]), oscale, __shfl_sync(0xffffffffu, bias_r[c], j)));
    const int y = 2 * Y + (c >> 1), x = 2 * X + (c & 1);
    for (int mk = 0; mk < p.hyper_nm; ++mk) {
      const float4* hy = reinterpret_cast<const float4*>(p.hyper + (size_t(bw) * p.hyper_nm + mk) * 32);
      float o0 = 0.f, o1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float4 h0 = __ldg(hy + j), h1 = __ldg(hy + j + 1);
        o0 = fmaf(h0.x, f[4 * j], o0); o0 = fmaf(h0.y, f[4 * j + 1], o0); o0 = fmaf(h0.z, f[4 * j + 2], o0); o0 = fmaf(h0.w, f[4 * j + 3], o0);
        o1 = fmaf(h1.x, f[4 * j + 4], o1); o1 = fmaf(h1.y, f[4 * j + 5], o1); o1 = fmaf(h1.z, f[4 * j + 6], o1); o1 = fmaf(h1.w, f[4 * j + 7], o1);
      }
      if (ok) p.low[((size_t(b) * p.hyper_nm + mk) * 256 + y) * 256 + x] = o0 + o1;
    }
  }
}

// Epilogue of one accumulator tile, executed by one warp for its 32 TMEM lanes (rows row0 .. row0+31), 32 columns at
// a time: tcgen05.ld -> scale + bias (+ residual) + activation in registers -> the warp's 32x32 block is staged in
// shared memory and written out by ONE asynchronous TMA store (rows / columns beyond M / N are clipped by the TMA
// unit).  The pipeline trace in profiles/r01_gemm_trace.txt shows why: while the TMA producer saturates the L2, every
// ld.global / st.global issued by the epilogue warps takes 0.5-6 k cycles, which made the first two epilogue versions
// (row-per-thread stores, then staged coalesced stores) as long as the main loop itself.  Here the only global loads
// left are the bias (prefetched before the accumulator is ready) and the optional residual (prefetched one chunk
// ahead), and the stores never stall the warp.
template <int BN, bool OUT_HALF, int ACT>
__device__ __forceinline__ void gemm_epilogue_warp(const GemmParams& p, const CUtensorMap* tmC, const CUtensorMap* tmC16, uint32_t t_addr, int row0, int n0,
                                                   int bt, float oscale, uint8_t* stage /*8 KiB, 1024-B aligned*/, uint32_t& nstaged,
                                                   int lane, const float (&bias_r)[(BN + 31) / 32], int chunk_begin, int chunk_end) {
  if constexpr (OUT_HALF) {
    gemm_epilogue_warp_half<BN, ACT>(p, tmC, t_addr, row0, n0, bt, oscale, stage, nstaged, lane, bias_r, chunk_begin, chunk_end);
    return;
  }
  if constexpr (ACT == 3) {
    gemm_epilogue_warp_upscale2<BN>(p, t_addr, row0, oscale, lane, bias_r, chunk_begin, chunk_end);
    return;
  }
  constexpr int NCH = (BN + 31) / 32;
  const int row = row0 + lane;
  const float* res_row = nullptr;
  if (p.res != nullptr && row < p.M) res_row = p.res + size_t(p.res_mod > 0 ? row % p.res_mod : row) * p.ldr + n0;
  float4 rn[8];
  auto load_res = [&](int c) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      rn[j] = (res_row != nullptr && n0 + c * 32 + 4 * j < p.N) ? *reinterpret_cast<const float4*>(res_row + c * 32 + 4 * j)
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (p.res != nullptr) load_res(chunk_begin);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c < chunk_begin || c >= chunk_end) continue;
    if (n0 + c * 32 >= p.N) break;
    if (BN % 32 != 0 && c == NCH - 1) {
      // 16-column tail of a tile whose width is not a multiple of 32 (BN = 144): same steps through a 16 x 32 box
      uint32_t v[16];
      tmem_ld16(t_addr + uint32_t(c * 32), v);
      tc_wait_ld();
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaf(__uint_as_float(v[j]), oscale, __shfl_sync(0xffffffffu, bias_r[c], j));
      if (p.res != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { f[4 * j] += rn[j].x; f[4 * j + 1] += rn[j].y; f[4 * j + 2] += rn[j].z; f[4 * j + 3] += rn[j].w; }
      }
      uint8_t* buf = stage + (nstaged % 2) * (GEMM_EPI_WARP_SMEM / 2);
      ++nstaged;
      if (elect_one()) tma_store_wait_read<1>();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(buf + lane * 64 + j * 16) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        if (p.accumulate) tma_reduce_add_3d(tmC16, buf, n0 + c * 32, row0, bt);
        else tma_store_3d(tmC16, buf, n0 + c * 32, row0, bt);
        tma_store_commit();
      }
      continue;
    }
    const bool trc = kGemmExp && (p.dbg != nullptr) && (threadIdx.x >> 5) == 4 && lane == 0 && nstaged < 32;
    const int tslot = 2048 + int(nstaged) * 8;
    if (trc) gemm_dbg(p, tslot + 0);
    uint32_t v[32];
    tmem_ld32(t_addr + uint32_t(c * 32), v);
    float4 rc[8];
    if (p.res != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) rc[j] = rn[j];
      if (c + 1 < chunk_end) load_res(c + 1);
    }
    tc_wait_ld();
    if (trc) gemm_dbg(p, tslot + 1);
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaf(__uint_as_float(v[j]), oscale, __shfl_sync(0xffffffffu, bias_r[c], j));
    if (p.res != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { f[4 * j] += rc[j].x; f[4 * j + 1] += rc[j].y; f[4 * j + 2] += rc[j].z; f[4 * j + 3] += rc[j].w; }
    }
    if (ACT == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
    }
    // staging blocks rotate (2 x 4 KiB fp32 / 4 x 2 KiB fp16): the block about to be overwritten was handed to the TMA
    // unit NBUF chunks ago, so the warp only waits for that store to have read it while the newer ones are in flight
    constexpr int NBUF = OUT_HALF ? 4 : 2;
    uint8_t* buf = stage + (nstaged % NBUF) * (GEMM_EPI_WARP_SMEM / NBUF);
    ++nstaged;
    if (trc) gemm_dbg(p, tslot + 2);
    if (elect_one()) tma_store_wait_read<NBUF - 1>();
    __syncwarp();
    if (trc) gemm_dbg(p, tslot + 3);
    if (OUT_HALF) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half2 h0 = __floats2half2_rn(f[8 * j], f[8 * j + 1]), h1 = __floats2half2_rn(f[8 * j + 2], f[8 * j + 3]);
        __half2 h2 = __floats2half2_rn(f[8 * j + 4], f[8 * j + 5]), h3 = __floats2half2_rn(f[8 * j + 6], f[8 * j + 7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(buf + lane * 64 + j * 16) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    }
    if (trc) gemm_dbg(p, tslot + 4);
    fence_proxy_async_smem();
    __syncwarp();
    if (trc) gemm_dbg(p, tslot + 5);
    if (elect_one()) {                        // same lane every time (full-warp mask): bulk groups are per thread
      if (!OUT_HALF && p.accumulate) tma_reduce_add_3d(tmC, buf, n0 + c * 32, row0, bt);
      else tma_store_3d(tmC, buf, n0 + c * 32, row0, bt);
      tma_store_commit();
    }
    if (trc) gemm_dbg(p, tslot + 6);
  }
}

// bias of the tile's columns, one value per lane and 32-column chunk; loaded before the accumulator is ready
template <int BN>
__device__ __forceinline__ void gemm_load_bias(const GemmParams& p, int n0, int lane, float (&bias_r)[(BN + 31) / 32]) {
#pragma unroll
  for (int c = 0; c < (BN + 31) / 32; ++c) {
    const int col = n0 + c * 32 + lane;
    bias_r[c] = (p.bias != nullptr && col < p.N) ? __ldg(p.bias + col) : 0.f;
  }
}

// CM = 2: two CTAs of a cluster work on vertically adjacent tiles (2m, n) / (2m+1, n) and share the B tile: each loads
// half of it and TMA-multicasts the half into both CTAs' shared memory, which cuts the L2->SM bytes per flop from
// (128 + BN) to (128 + BN/2) per k-block and moves the main loop from feed-bound (~80 B/clk/SM) to MMA-bound.
template <int BN, bool OUT_HALF, int ACT, int CM = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  static_assert(BN % 32 == 0, "the 1-CTA kernel takes whole 32-column chunks");
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::kStages;
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
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int num_tiles = tiles_mn * p.batch;
  // CM == 2: `unit` = pair of M tiles; tile index of this CTA = 2 * (unit's m-pair) + rank, same n
  const int crank = (CM == 2) ? int(cluster_ctarank()) : 0;
  const int first = (CM == 2) ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int stride = (CM == 2) ? int(gridDim.x >> 1) : int(gridDim.x);
  const int num_units = (CM == 2) ? num_tiles / 2 : num_tiles;
  auto unit_to_tile = [&](int u) { return (CM == 2) ? ((u % (p.tiles_m / 2)) * 2 + crank + (u / (p.tiles_m / 2)) * p.tiles_m) : u; };

  if (threadIdx.x == 0) gemm_dbg_wall(p, false);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CM);          // CM == 2: both CTAs' MMAs must have released the stage (the peer writes into it)
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], GEMM_EPI_WARPS);
    mbar_init(&tempty[1], GEMM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  if (CM == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) { gemm_dbg(p, 1); }
  pdl_trigger();
  pdl_wait();                                     // the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    // The whole warp runs the loop with warp-uniform state and one elected lane issues: under a divergent `lane == 0`
    // branch the compiler cannot prove the operands uniform and wraps every TMA / MMA instruction in an
    // ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop, which cost ~500 clk per k-block (profiles/r01_gemm_trace.txt).
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = first; u < num_units; u += stride) {
        const int t = unit_to_tile(u);
        const int bt = t / tiles_mn, tt = t % tiles_mn;
        const int m0 = (tt % p.tiles_m) * GEMM_BM;
        const int n0 = (tt / p.tiles_m) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + GEMM_BM * 128;
          if (kGemmExp && (p.dbg_mode & 1)) {
            if (elect_one()) mbar_arrive(&full[stage]);
          } else if (elect_one()) {
            mbar_expect_tx(&full[stage], Cfg::kStageBytes);
            const int ka = (p.a_wrap_kb > 0 && kb >= p.a_wrap_kb) ? kb - p.a_wrap_kb : kb;
            if (p.a_rank3) tma_load_3d(sa, &tmA, &full[stage], ka * GEMM_BK, bt, m0);
            else tma_load_2d(sa, &tmA, &full[stage], ka * GEMM_BK, m0);
            if (CM == 2) tma_load_2d_mcast(sb + crank * (BN / 2) * 128, &tmB, &full[stage], kb * GEMM_BK, n0 + crank * (BN / 2), 0x3);
            else tma_load_2d(sb, &tmB, &full[stage], kb * GEMM_BK, n0);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (warp-uniform loop, elected lane issues)
    {
      constexpr uint32_t idesc = umma_idesc_f16(GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int ti = 0;
      for (int u = first; u < num_units; u += stride, ++ti) {
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
            if (!kGemmExp || !(p.dbg_mode & 2)) {
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k) {
                // advance 16 fp16 (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
                tc_mma_f16(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0);
              }
            }
            if (CM == 2) tc_commit_mcast(&empty[stage], 0x3); else tc_commit(&empty[stage]);
            if (kb == num_kb - 1) tc_commit(&tfull[as]);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        gemm_dbg(p, 16 + ti * 64 + 60);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int q = warp & 3;                       // TMEM lane quadrant of this warp
    const float oscale = (p.out_scale != 0.f) ? p.out_scale : 1.0f;
    uint8_t* my_stage = epi_stage + (warp - 4) * GEMM_EPI_WARP_SMEM;
    uint32_t nstaged = 0;
    constexpr int NCH = (BN + 31) / 32, SPLIT = OUT_HALF ? ((NCH + 1) / 4) * 2 : (NCH + 1) / 2;   // fp16: whole chunk pairs per group
    const int cb = (warp < 8) ? 0 : SPLIT, ce = (warp < 8) ? SPLIT : NCH;
    int as = 0, eti = 0;
    uint32_t aphase = 0;
    for (int u = first; u < num_units; u += stride) {
      const int t = unit_to_tile(u);
      const int bt = t / tiles_mn, tt = t % tiles_mn;
      const int m0 = (tt % p.tiles_m) * GEMM_BM;
      const int n0 = (tt / p.tiles_m) * BN;
      float bias_r[(BN + 31) / 32];
      gemm_load_bias<BN>(p, n0, lane, bias_r);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if (warp == 4 && lane == 0) gemm_dbg(p, 16 + eti * 64 + 61);
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      if (!kGemmExp || !(p.dbg_mode & 4))
        gemm_epilogue_warp<BN, OUT_HALF, ACT>(p, &tmC, &tmC, t_addr, m0 + q * 32, n0, bt, oscale, my_stage, nstaged, lane, bias_r, cb, ce);
      // accumulator drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (warp == 4 && lane == 0) gemm_dbg(p, 16 + eti * 64 + 62);
      ++eti;
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (elect_one()) tma_store_wait_all();
  }

  tc_fence_before();
  if (threadIdx.x == 0) gemm_dbg(p, 2);
  if (CM == 2) cluster_sync_all(); else __syncthreads();     // the peer may still multicast into / signal this CTA
  if (threadIdx.x == 0) { gemm_dbg(p, 3); gemm_dbg_wall(p, true); }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// Host-side launcher: picks the N tile that wastes the fewest SM-waves for this shape.
int launch_gemm_tc(const __half* A, int lda, const __half* B, int ldb, const GemmParams& p, bool out_half, int act,
                   int num_sms, cudaStream_t stream, int force_bn = 0, const CUtensorMap* a_map_rank3 = nullptr);

}  // namespace samrs
