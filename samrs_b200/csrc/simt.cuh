// CUDA-core kernels of the samrs_b200 engine: everything on the SAM path that is not a large GEMM /
// attention contraction.  All are HBM- or latency-bound; layouts are chosen so that every warp reads
// and writes contiguous 128-byte lines.
#pragma once
#include "common.cuh"

namespace samrs {

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ------------------------------------------------------------------------------------------------
// Sam.preprocess (SA/modeling/sam.py:164-174) fused with the patch gather of PatchEmbed's 16x16/s16 conv
// (SA/modeling/image_encoder.py:391-395): u8 HWC image -> A[4096][768] fp16, k = c*256 + iy*16 + ix.
// Pixels outside (H,W) are the zero padding applied after normalisation.
// ------------------------------------------------------------------------------------------------
__global__ void preprocess_im2col_kernel(const uint8_t* __restrict__ img, int H, int W, int chw /*1: CHW input*/,
                                         __half* __restrict__ A) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // one thread per (patch, c, iy), 16 ix each
  if (idx >= 4096 * 48) return;
  const int patch = idx / 48, rem = idx % 48, c = rem / 16, iy = rem % 16;
  const int py = patch / 64, px = patch % 64;
  const int y = py * 16 + iy;
  const float mean = (c == 0) ? 123.675f : (c == 1 ? 116.28f : 103.53f);
  const float sd = (c == 0) ? 58.395f : (c == 1 ? 57.12f : 57.375f);
  __half vals[16];
#pragma unroll
  for (int ix = 0; ix < 16; ++ix) {
    const int x = px * 16 + ix;
    float v = 0.f;
    if (y < H && x < W) {
      const uint8_t u = chw ? img[(size_t(c) * H + y) * W + x] : img[(size_t(y) * W + x) * 3 + c];
      v = __fdiv_rn(__fsub_rn(float(u), mean), sd);
    }
    vals[ix] = __float2half_rn(v);
  }
  uint4* dst = reinterpret_cast<uint4*>(A + size_t(patch) * 768 + c * 256 + iy * 16);
  dst[0] = *reinterpret_cast<uint4*>(&vals[0]);
  dst[1] = *reinterpret_cast<uint4*>(&vals[8]);
}

// ------------------------------------------------------------------------------------------------
// Encoder LayerNorm, streaming form: a block of 2 warps owns 4 consecutive rows (1024 blocks for 4096 rows = 7 blocks on
// almost every one of the 148 SMs instead of 3-or-4 blocks of 8 rows), each warp normalises 2 rows and loads the second
// while it reduces / stores the first, so the read and write bursts overlap instead of following each other.
// Same arithmetic (two-pass mean / biased variance, same summation order) as ln_rows_kernel below.
// ------------------------------------------------------------------------------------------------
template <int MAXV /* float4 per lane, exact: C == 128 * MAXV */>
__global__ void __launch_bounds__(64) ln_rows_stream_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, __half* __restrict__ out,
                                                            int ld_out, int rows) {
  const int lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * 4 + (threadIdx.x >> 5) * 2;
  pdl_trigger();
  pdl_wait();
  if (row0 >= rows) return;
  constexpr float Cf = float(128 * MAXV);
  float4 v[2][MAXV];
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const float4* src = reinterpret_cast<const float4*>(in + size_t(row0 + rr < rows ? row0 + rr : row0) * ld_in);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) v[rr][i] = src[lane + 32 * i];
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    if (row0 + rr >= rows) break;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) s += (v[rr][i].x + v[rr][i].y) + (v[rr][i].z + v[rr][i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / Cf;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const float a = v[rr][i].x - mean, b = v[rr][i].y - mean, c = v[rr][i].z - mean, d = v[rr][i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / Cf + eps);
    uint2* dst = reinterpret_cast<uint2*>(out + size_t(row0 + rr) * ld_out);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int k = lane + 32 * i;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + k);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + k);
      const float y0 = (v[rr][i].x - mean) * rstd * g.x + b.x;
      const float y1 = (v[rr][i].y - mean) * rstd * g.y + b.y;
      const float y2 = (v[rr][i].z - mean) * rstd * g.z + b.z;
      const float y3 = (v[rr][i].w - mean) * rstd * g.w + b.w;
      __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      dst[k] = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Row LayerNorm (nn.LayerNorm / LayerNorm2d over a contiguous channel row): one warp per row,
// two-pass mean / biased variance in registers.  OutT = __half feeds the next tensor-core GEMM,
// OutT = float is used by the decoder.  ACT 1 = GELU(erf) (output_upscaling, mask_downscaling).
// ------------------------------------------------------------------------------------------------
template <typename OutT, int ACT, int MAXV /* float4 per lane */>
__global__ void ln_rows_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ gamma,
                               const float* __restrict__ beta, float eps, OutT* __restrict__ out, int ld_out,
                               int rows, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(in + size_t(warp) * ld_in);
  const int nv = C >> 2;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int k = lane + 32 * i;
    v[i] = (k < nv) ? src[k] : make_float4(0, 0, 0, 0);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / float(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int k = lane + 32 * i;
    if (k < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / float(C) + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int k = lane + 32 * i;
    if (k < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + k);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + k);
      float y0 = (v[i].x - mean) * rstd * g.x + b.x;
      float y1 = (v[i].y - mean) * rstd * g.y + b.y;
      float y2 = (v[i].z - mean) * rstd * g.z + b.z;
      float y3 = (v[i].w - mean) * rstd * g.w + b.w;
      if (ACT == 1) { y0 = gelu_erf_f(y0); y1 = gelu_erf_f(y1); y2 = gelu_erf_f(y2); y3 = gelu_erf_f(y3); }
      if (sizeof(OutT) == 2) {
        __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + size_t(warp) * ld_out)[k] = pk;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + size_t(warp) * ld_out)[k] = make_float4(y0, y1, y2, y3);
      }
    }
  }
}

// Token rows of the mask decoder (C = 256): x = LayerNorm(x) in place and `with_pe` = x + pe, the q / k input of the next
// attention (transformer.py:161-179) - the arithmetic of ln_rows_kernel<float, 0, 2> followed by a plain add, one warp per row.
__global__ void __launch_bounds__(256) ln256_tok_kernel(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, const float* __restrict__ pe, float* __restrict__ with_pe, int rows) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4* xr = reinterpret_cast<float4*>(x + size_t(row) * 256);
  float4 v[2];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / 256.0f;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / 256.0f + eps);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + k);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + k);
    const float4 p = reinterpret_cast<const float4*>(pe + size_t(row) * 256)[k];
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    xr[k] = y;
    reinterpret_cast<float4*>(with_pe + size_t(row) * 256)[k] = make_float4(y.x + p.x, y.y + p.y, y.z + p.z, y.w + p.w);
  }
}

// fp32 -> fp16 cast of a contiguous buffer (n multiple of 4)
__global__ void cast_f32_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n4) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(in)[i];
  __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&h0);
  pk.y = *reinterpret_cast<uint32_t*>(&h1);
  reinterpret_cast<uint2*>(out)[i] = pk;
}

// ------------------------------------------------------------------------------------------------
// Neck helpers (SA/modeling/image_encoder.py:88-104): 3x3/pad-1 conv as im2col'ed GEMM, and the final
// LayerNorm2d written both token-major (decoder input) and NCHW (the `features` tensor of the API).
// ------------------------------------------------------------------------------------------------
__global__ void neck_im2col3x3_kernel(const __half* __restrict__ in /*[4096][256]*/, __half* __restrict__ out /*[4096][2304]*/) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (token, tap, 8-channel chunk)
  if (idx >= 4096 * 9 * 32) return;
  const int chunk = idx & 31, tap = (idx >> 5) % 9, token = idx / (9 * 32);
  const int y = (token >> 6) + tap / 3 - 1, x = (token & 63) + tap % 3 - 1;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (y >= 0 && y < 64 && x >= 0 && x < 64) v = reinterpret_cast<const uint4*>(in + size_t(y * 64 + x) * 256)[chunk];
  reinterpret_cast<uint4*>(out + size_t(token) * 2304 + tap * 256)[chunk] = v;
}

__global__ void transpose_tok_to_nchw_kernel(const float* __restrict__ in /*[4096][C]*/, float* __restrict__ out /*[C][4096]*/, int C) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) tile[i][threadIdx.x] = in[size_t(t0 + i) * C + c0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) out[size_t(c0 + i) * 4096 + t0 + threadIdx.x] = tile[threadIdx.x][i];
}
__global__ void transpose_nchw_to_tok_kernel(const float* __restrict__ in /*[C][4096]*/, float* __restrict__ out /*[4096][C]*/, int C) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) tile[i][threadIdx.x] = in[size_t(c0 + i) * 4096 + t0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) out[size_t(t0 + i) * C + c0 + threadIdx.x] = tile[threadIdx.x][i];
}

// ------------------------------------------------------------------------------------------------
// fp32 SGEMM for the mask decoder (precision-sensitive, 2 % of the FLOPs):
//   C[M,N] = act(A[M,K] W[N,K]^T + bias[N] + R[m % rmod][N])      K % 16 == 0
// 128x64 block tile, 16-deep k slices, 8x4 outputs per thread, operands transposed into smem.
// ------------------------------------------------------------------------------------------------
struct SgemmParams {
  const float* A; int lda;
  const float* W; int ldw;
  float* C; int ldc;
  const float* bias;
  const float* R; int ldr; int rmod;
  int M, N, K;
  int act;              // 0 none, 1 relu, 2 gelu
};

__global__ void __launch_bounds__(256)
sgemm_tn_kernel(const SgemmParams p) {
  __shared__ __align__(16) float As[2][16][128 + 4];
  __shared__ __align__(16) float Ws[2][16][64 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 64;
  const int ty = tid / 16, tx = tid % 16;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // global->register staging: A tile 128x16 = 512 float4 (2 per thread), W tile 64x16 = 256 float4 (1 per thread)
  const int a_row0 = tid / 4, a_k4 = (tid % 4) * 4;            // rows a_row0 and a_row0 + 64
  const int w_row = tid / 4, w_k4 = (tid % 4) * 4;
  float4 ra0, ra1, rw;
  auto gload = [&](int k0) {
    const int r0 = m0 + a_row0, r1 = r0 + 64;
    ra0 = (r0 < p.M) ? *reinterpret_cast<const float4*>(p.A + size_t(r0) * p.lda + k0 + a_k4) : make_float4(0, 0, 0, 0);
    ra1 = (r1 < p.M) ? *reinterpret_cast<const float4*>(p.A + size_t(r1) * p.lda + k0 + a_k4) : make_float4(0, 0, 0, 0);
    const int wr = n0 + w_row;
    rw = (wr < p.N) ? *reinterpret_cast<const float4*>(p.W + size_t(wr) * p.ldw + k0 + w_k4) : make_float4(0, 0, 0, 0);
  };
  auto sstore = [&](int buf) {
    As[buf][a_k4 + 0][a_row0] = ra0.x; As[buf][a_k4 + 1][a_row0] = ra0.y; As[buf][a_k4 + 2][a_row0] = ra0.z; As[buf][a_k4 + 3][a_row0] = ra0.w;
    As[buf][a_k4 + 0][a_row0 + 64] = ra1.x; As[buf][a_k4 + 1][a_row0 + 64] = ra1.y; As[buf][a_k4 + 2][a_row0 + 64] = ra1.z; As[buf][a_k4 + 3][a_row0 + 64] = ra1.w;
    Ws[buf][w_k4 + 0][w_row] = rw.x; Ws[buf][w_k4 + 1][w_row] = rw.y; Ws[buf][w_k4 + 2][w_row] = rw.z; Ws[buf][w_k4 + 3][w_row] = rw.w;
  };
  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = p.K / 16;
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    const float* rrow = p.R ? p.R + size_t(p.rmod > 0 ? m % p.rmod : m) * p.ldr : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + n);
      if (rrow) v += rrow[n];
      if (p.act == 1) v = fmaxf(v, 0.f);
      else if (p.act == 2) v = gelu_erf_f(v);
      p.C[size_t(m) * p.ldc + n] = v;
    }
  }
}

// Small-M variant for the token side of the decoder (M = prompts x tokens <= ~2k rows): these GEMMs are latency- not
// throughput-bound (a few MFLOP each, ~60 of them per decode), so the kernel is built around ONE round trip to memory:
// a block of 128 threads owns a 32 x 32 output tile and at most 256 of K; it brings its whole A and W panels (32 rows x
// 256 floats each, k contiguous, row pitch 260 floats = conflict-free float4 reads for rows 1 apart) into shared memory with
// cp.async - every load of the block in flight at once - and then runs 64 float4 steps of 2 x 4 dot products per thread.
// K > 256 is split across blockIdx.z into a workspace that splitk_reduce_kernel folds in a fixed order; the k order
// inside a block is ascending, so results equal the previous kernel's (64-deep slices through transposed smem, 4 serial
// load -> sync -> compute rounds and 8-way bank conflicts on the transposing stores: 17 us per launch against ~4 us).
constexpr int SGT_KMAX = 256;
constexpr int SGT_PITCH = SGT_KMAX + 4;
constexpr int SGT_SMEM = 2 * 32 * SGT_PITCH * 4;
template <int TN /*output columns per block: 32, or 16 when 32-wide tiles would leave most SMs without a block*/>
__global__ void __launch_bounds__(128)
sgemm_small_kernel(const SgemmParams p, int k_per_split, float* __restrict__ ws /*[splits][M][N] or null*/) {
  extern __shared__ __align__(16) float sgt_smem[];
  float* As = sgt_smem;
  float* Ws = sgt_smem + 32 * SGT_PITCH;
  const int tid = threadIdx.x;
  constexpr int NJ = TN / 8;
  const int m0 = blockIdx.x * 32, n0 = blockIdx.y * TN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kn = min(p.K, kbeg + k_per_split) - kbeg;          // <= 256, multiple of 16 (host)
  const int k4n = kn >> 2;
  for (int i = tid; i < 32 * k4n; i += 128) {
    const int r = i / k4n, c4 = i - r * k4n;
    const int ar = m0 + r, wr = n0 + r;
    float* da = As + r * SGT_PITCH + 4 * c4;
    float* dw = Ws + r * SGT_PITCH + 4 * c4;
    if (ar < p.M) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(da)), "l"(p.A + size_t(ar) * p.lda + kbeg + 4 * c4) : "memory");
    } else {
      *reinterpret_cast<float4*>(da) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (r < TN) {
      if (wr < p.N) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dw)), "l"(p.W + size_t(wr) * p.ldw + kbeg + 4 * c4) : "memory");
      } else {
        *reinterpret_cast<float4*>(dw) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const int tm = tid >> 3, tn = tid & 7;                        // rows tm, tm + 16; columns tn, tn + 8, tn + 16, tn + 24
  float acc[2][NJ];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int c4 = 0; c4 < k4n; ++c4) {
    float4 a[2], w[NJ];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = *reinterpret_cast<const float4*>(As + (tm + 16 * i) * SGT_PITCH + 4 * c4);
#pragma unroll
    for (int j = 0; j < NJ; ++j) w[j] = *reinterpret_cast<const float4*>(Ws + (tn + 8 * j) * SGT_PITCH + 4 * c4);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float v = acc[i][j];
        v = fmaf(a[i].x, w[j].x, v); v = fmaf(a[i].y, w[j].y, v); v = fmaf(a[i].z, w[j].z, v); v = fmaf(a[i].w, w[j].w, v);
        acc[i][j] = v;
      }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + tm + 16 * i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int n = n0 + tn + 8 * j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (ws) { ws[(size_t(blockIdx.z) * p.M + m) * p.N + n] = v; continue; }
      if (p.bias) v += __ldg(p.bias + n);
      if (p.R) v += p.R[size_t(p.rmod > 0 ? m % p.rmod : m) * p.ldr + n];
      if (p.act == 1) v = fmaxf(v, 0.f);
      else if (p.act == 2) v = gelu_erf_f(v);
      p.C[size_t(m) * p.ldc + n] = v;
    }
  }
}
__global__ void splitk_reduce_kernel(const SgemmParams p, const float* __restrict__ ws, int splits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.M * p.N) return;
  const int m = i / p.N, n = i % p.N;
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += ws[size_t(z) * p.M * p.N + i];
  if (p.bias) v += __ldg(p.bias + n);
  if (p.R) v += p.R[size_t(p.rmod > 0 ? m % p.rmod : m) * p.ldr + n];
  if (p.act == 1) v = fmaxf(v, 0.f);
  else if (p.act == 2) v = gelu_erf_f(v);
  p.C[size_t(m) * p.ldc + n] = v;
}


// ------------------------------------------------------------------------------------------------
// Prompt encoder (SA/modeling/prompt_encoder.py:73-100,128-173,190-219) + decoder token assembly
// (SA/modeling/mask_decoder.py:127-129).  tokens[b] = [iou, mask0..3, points..., (pad), box corners].
// One block per prompt, one thread per channel pair (sin at c, cos at c + 128).
// ------------------------------------------------------------------------------------------------
struct PromptParams {
  const float* gauss;          // [2][128]
  const float* point_emb;      // [4][256]
  const float* not_a_point;    // [256]
  const float* iou_token;      // [256]
  const float* mask_tokens;    // [4][256]
  const float* points;         // [B][NP][2] or null
  const int* labels;           // [B][NP]
  const float* boxes;          // [B][4] or null
  int NP;                      // points per prompt (without pad)
  int pad;                     // append the (0,0)/-1 pad point (points given, no boxes)
  int T;                       // tokens per prompt
  float* tokens;               // [B][T][256]
};

__device__ __forceinline__ void pe_pair(const float* gauss, float x, float y, int c, float& s, float& co) {
  // forward_with_coords (:212-219) -> _pe_encoding (:190-197): coords already include the +0.5 shift
  float cx = x / 1024.0f, cy = y / 1024.0f;
  cx = 2.0f * cx - 1.0f;
  cy = 2.0f * cy - 1.0f;
  float v = __fadd_rn(__fmul_rn(cx, gauss[c]), __fmul_rn(cy, gauss[128 + c]));
  v = 6.283185307179586f * v;
  s = sinf(v);
  co = cosf(v);
}

__global__ void prompt_tokens_kernel(const PromptParams p) {
  const int b = blockIdx.x, c = threadIdx.x;     // 128 threads
  float* tok = p.tokens + size_t(b) * p.T * 256;
  tok[c] = p.iou_token[c];
  tok[c + 128] = p.iou_token[c + 128];
  for (int i = 0; i < 4; ++i) {
    tok[(1 + i) * 256 + c] = p.mask_tokens[i * 256 + c];
    tok[(1 + i) * 256 + c + 128] = p.mask_tokens[i * 256 + c + 128];
  }
  int t = 5;
  if (p.points) {
    for (int i = 0; i < p.NP + p.pad; ++i, ++t) {
      float x = 0.f, y = 0.f;
      int lab = -1;
      if (i < p.NP) {
        x = p.points[(size_t(b) * p.NP + i) * 2 + 0];
        y = p.points[(size_t(b) * p.NP + i) * 2 + 1];
        lab = p.labels[size_t(b) * p.NP + i];
      }
      float s, co;
      pe_pair(p.gauss, x + 0.5f, y + 0.5f, c, s, co);
      if (lab == -1) { s = 0.f + p.not_a_point[c]; co = 0.f + p.not_a_point[c + 128]; }
      else if (lab == 0) { s += p.point_emb[c]; co += p.point_emb[c + 128]; }
      else if (lab == 1) { s += p.point_emb[256 + c]; co += p.point_emb[256 + c + 128]; }
      tok[t * 256 + c] = s;
      tok[t * 256 + c + 128] = co;
    }
  }
  if (p.boxes) {
    for (int i = 0; i < 2; ++i, ++t) {
      const float x = p.boxes[size_t(b) * 4 + 2 * i], y = p.boxes[size_t(b) * 4 + 2 * i + 1];
      float s, co;
      pe_pair(p.gauss, x + 0.5f, y + 0.5f, c, s, co);
      s += p.point_emb[(2 + i) * 256 + c];
      co += p.point_emb[(2 + i) * 256 + c + 128];
      tok[t * 256 + c] = s;
      tok[t * 256 + c + 128] = co;
    }
  }
}

// get_dense_pe (:62-71,199-210): pe[token][c], token = y*64+x, coords ((x+0.5)/64, (y+0.5)/64)
__global__ void dense_pe_kernel(const float* __restrict__ gauss, float* __restrict__ pe /*[4096][256]*/) {
  const int token = blockIdx.x, c = threadIdx.x;   // 128 threads
  const float fx = (float(token & 63) + 0.5f) / 64.0f, fy = (float(token >> 6) + 0.5f) / 64.0f;
  const float cx = 2.0f * fx - 1.0f, cy = 2.0f * fy - 1.0f;
  float v = __fadd_rn(__fmul_rn(cx, gauss[c]), __fmul_rn(cy, gauss[128 + c]));
  v = 6.283185307179586f * v;
  pe[size_t(token) * 256 + c] = sinf(v);
  pe[size_t(token) * 256 + c + 128] = cosf(v);
}

// mask_downscaling (SA/modeling/prompt_encoder.py:51-59,102-105) fused with `src = features + dense`
// (SA/modeling/mask_decoder.py:136-137): one thread per (prompt, token) -> src[b][token][0..255].
struct MaskEmbedParams {
  const float* mask;      // [B][256][256]
  const float *w0, *b0, *g1, *be1, *w3, *b3, *g4, *be4, *w6, *b6;
  const float* feat;      // [4096][256] token-major image embedding
  float* src;             // [B][4096][256]
};
__global__ void mask_embed_src_kernel(const MaskEmbedParams p, int B) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 4096) return;
  const int b = idx / 4096, token = idx % 4096, ty = token >> 6, tx = token & 63;
  const float* m = p.mask + size_t(b) * 65536;
  float h2[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) h2[o] = p.b3[o];
#pragma unroll
  for (int py = 0; py < 2; ++py)
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      // conv0 (1->4, k2 s2) at 128x128 position (2ty+py, 2tx+px)
      const int y0 = (2 * ty + py) * 2, x0 = (2 * tx + px) * 2;
      const float i00 = m[y0 * 256 + x0], i01 = m[y0 * 256 + x0 + 1], i10 = m[(y0 + 1) * 256 + x0], i11 = m[(y0 + 1) * 256 + x0 + 1];
      float h[4];
      float mu = 0.f;
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        h[o] = p.b0[o] + p.w0[o * 4 + 0] * i00 + p.w0[o * 4 + 1] * i01 + p.w0[o * 4 + 2] * i10 + p.w0[o * 4 + 3] * i11;
        mu += h[o];
      }
      mu *= 0.25f;
      float var = 0.f;
#pragma unroll
      for (int o = 0; o < 4; ++o) var += (h[o] - mu) * (h[o] - mu);
      var *= 0.25f;
      const float rs = 1.0f / sqrtf(var + 1e-6f);
#pragma unroll
      for (int o = 0; o < 4; ++o) h[o] = gelu_erf_f(p.g1[o] * ((h[o] - mu) * rs) + p.be1[o]);
      // conv3 (4->16, k2 s2): weight [16][4][2][2]
#pragma unroll
      for (int o = 0; o < 16; ++o)
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) h2[o] += p.w3[((o * 4 + ci) * 2 + py) * 2 + px] * h[ci];
    }
  float mu = 0.f;
#pragma unroll
  for (int o = 0; o < 16; ++o) mu += h2[o];
  mu *= (1.0f / 16.0f);
  float var = 0.f;
#pragma unroll
  for (int o = 0; o < 16; ++o) var += (h2[o] - mu) * (h2[o] - mu);
  var *= (1.0f / 16.0f);
  const float rs = 1.0f / sqrtf(var + 1e-6f);
#pragma unroll
  for (int o = 0; o < 16; ++o) h2[o] = gelu_erf_f(p.g4[o] * ((h2[o] - mu) * rs) + p.be4[o]);
  float* dst = p.src + (size_t(b) * 4096 + token) * 256;
  const float* f = p.feat + size_t(token) * 256;
  for (int c = 0; c < 256; ++c) {
    float v = p.b6[c];
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) v += p.w6[c * 16 + ci] * h2[ci];
    dst[c] = f[c] + v;
  }
}

// ------------------------------------------------------------------------------------------------
// Decoder attention (SA/modeling/transformer.py:218-240), three shapes.
// ------------------------------------------------------------------------------------------------
// (1) token self-attention: q,k,v [B][T][256] already projected, 8 heads x 32.  One block per prompt.
__global__ void tok_self_attn_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                     int ld /*row pitch of q, k, v*/, float* __restrict__ out, int T) {
  extern __shared__ float sm[];
  float* sq = sm;
  float* sk = sq + T * 256;
  float* sv = sk + T * 256;
  float* sc = sv + T * 256;               // [8][T][T]
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < T * 256; i += blockDim.x) {
    const size_t g = (size_t(b) * T + (i >> 8)) * ld + (i & 255);
    sq[i] = q[g];
    sk[i] = k[g];
    sv[i] = v[g];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * T * T; i += blockDim.x) {
    const int h = i / (T * T), tq = (i / T) % T, tk = i % T;
    float s = 0.f;
    for (int c = 0; c < 32; ++c) s = fmaf(sq[tq * 256 + h * 32 + c], sk[tk * 256 + h * 32 + c], s);
    sc[i] = s / sqrtf(32.0f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * T; i += blockDim.x) {
    float* row = sc + i * T;
    float mx = -INFINITY;
    for (int j = 0; j < T; ++j) mx = fmaxf(mx, row[j]);
    float sum = 0.f;
    for (int j = 0; j < T; ++j) { row[j] = expf(row[j] - mx); sum += row[j]; }
    for (int j = 0; j < T; ++j) row[j] /= sum;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * 256; i += blockDim.x) {
    const int tq = i / 256, ch = i % 256, h = ch / 32;
    float acc = 0.f;
    for (int j = 0; j < T; ++j) acc = fmaf(sc[(h * T + tq) * T + j], sv[j * 256 + ch], acc);
    out[size_t(b) * T * 256 + i] = acc;
  }
}

// (2) token -> image cross attention: q [B][T][128]; K,V rows of `ld` floats, prompt stride kv_bstride
//     (0 = shared by all prompts); 8 heads x 16.  One block per (prompt, head, key chunk of 4096 / gridDim.z keys), one WARP
//     per query token: the block stages 128 keys' K and V head slices in shared memory (coalesced 64-byte row segments),
//     lane l of every warp then owns keys l, l+32, ... of the tile and runs its own online softmax over them (running max,
//     sum and 16 weighted-V accumulators in registers), so K/V are read once per block and nothing is reduced across
//     threads until the very end: one 18-value warp merge per token.  The first version reduced 17 values across the
//     block for every token and key pair (595 shuffles + 21 barriers per thread) and ran at 1.3 TB/s; this one is bound by
//     the K/V read.  Per-chunk partials (acc[16], sum, max) go to `part`; t2i_combine_kernel merges the chunks.
constexpr int T2I_TILE = 128;       // keys per shared-memory tile
constexpr int T2I_PITCH = 20;       // floats per staged row (16 + 4 pad: conflict-free float4 reads at row stride 1)
__global__ void __launch_bounds__(512)
t2i_attn_kernel(const float* __restrict__ q, const float* __restrict__ K, const float* __restrict__ V, int ld,
                size_t kv_bstride, float* __restrict__ part /*[B][8][T][nch][18]*/, int T) {
  const int b = blockIdx.x, h = blockIdx.y, ch = blockIdx.z, nch = gridDim.z;
  const int tid = threadIdx.x, lane = tid & 31, t = tid >> 5;       // warp t = query token t (blockDim.x = 32 * T)
  const int keys_per_chunk = 4096 / nch;
  const float* Kb = K + size_t(b) * kv_bstride + h * 16;
  const float* Vb = V + size_t(b) * kv_bstride + h * 16;
  __shared__ __align__(16) float sk[2][T2I_TILE * T2I_PITCH];       // two stages: the next tile streams in (cp.async)
  __shared__ __align__(16) float sv[2][T2I_TILE * T2I_PITCH];       // while this one is consumed
  float qv[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) qv[c] = q[(size_t(b) * T + t) * 128 + h * 16 + c] * 0.25f;     // 1 / sqrt(16) folded into q
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  auto load_tile = [&](int buf, int k0) {
    for (int i = tid; i < T2I_TILE * 4; i += blockDim.x) {           // 4 x 16 bytes per key row and operand
      const int r = i >> 2, c4 = i & 3;
      const size_t key = size_t(ch) * keys_per_chunk + k0 + r;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&sk[buf][r * T2I_PITCH + 4 * c4])), "l"(Kb + key * ld + 4 * c4) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&sv[buf][r * T2I_PITCH + 4 * c4])), "l"(Vb + key * ld + 4 * c4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int ntiles = keys_per_chunk / T2I_TILE;
  load_tile(0, 0);
  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < ntiles) {
      load_tile(buf ^ 1, (it + 1) * T2I_TILE);
      asm volatile("cp.async.wait_group 1;" ::: "memory");           // this tile has landed, the next may still be in flight
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < T2I_TILE / 32; ++j) {
      const float* kr = &sk[buf][(lane + 32 * j) * T2I_PITCH];
      const float* vr = &sv[buf][(lane + 32 * j) * T2I_PITCH];
      float sc = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + 4 * c4);
        sc = fmaf(qv[4 * c4], kk.x, sc); sc = fmaf(qv[4 * c4 + 1], kk.y, sc); sc = fmaf(qv[4 * c4 + 2], kk.z, sc); sc = fmaf(qv[4 * c4 + 3], kk.w, sc);
      }
      if (sc > m) {                                                  // new running max: rescale what has been accumulated
        const float a = expf(m - sc);                                // exp(-inf) = 0 on the first key
        l *= a;
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] *= a;
        m = sc;
      }
      const float pr = expf(sc - m);
      l += pr;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 vv = *reinterpret_cast<const float4*>(vr + 4 * c4);
        acc[4 * c4] = fmaf(pr, vv.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(pr, vv.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(pr, vv.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(pr, vv.w, acc[4 * c4 + 3]);
      }
    }
    __syncthreads();                                                 // the stage is overwritten by the load after next
  }
  // merge the 32 lanes' (max, sum, acc) of this token
  float M = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  const float w = expf(m - M);
  l *= w;
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] *= w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l += __shfl_xor_sync(0xffffffffu, l, o);
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  }
  if (lane == 0) {
    float* o = part + ((((size_t(b) * 8 + h) * T + t) * nch) + ch) * 18;
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = acc[c];
    o[16] = l;
    o[17] = M;
  }
}
// merge the key-chunk partials: out[b][t][h*16 + c] = sum_i acc_i e^(m_i - M) / sum_i l_i e^(m_i - M)
__global__ void t2i_combine_kernel(const float* __restrict__ part, float* __restrict__ out /*[B][T][128]*/, int B, int T, int nch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // (b, h, t, c)
  if (i >= B * 8 * T * 16) return;
  const int c = i % 16, t = (i / 16) % T, h = (i / (16 * T)) % 8, b = i / (16 * T * 8);
  const float* p = part + (((size_t(b) * 8 + h) * T + t) * nch) * 18;
  float M = -INFINITY;
  for (int k = 0; k < nch; ++k) M = fmaxf(M, p[k * 18 + 17]);
  float num = 0.f, den = 0.f;
  for (int k = 0; k < nch; ++k) {
    const float w = expf(p[k * 18 + 17] - M);
    num = fmaf(p[k * 18 + c], w, num);
    den = fmaf(p[k * 18 + 16], w, den);
  }
  out[(size_t(b) * T + t) * 128 + h * 16 + c] = num / den;
}

// split an fp32 value into fp16 hi + lo (hi + lo == v to ~2^-22 relative): operands of the 3-term tensor-core GEMM
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// src0 = image embedding + no_mask_embed (mask_decoder.py:136 with the dense "no mask" prompt, prompt_encoder.py:165-168), written
// as fp32 (residual of the first image-side update) and as the split-fp16 A operand [hi | lo] of the layer-0 projection GEMM
__global__ void add_rowvec_split_kernel(const float* __restrict__ a, const float* __restrict__ vec, float* __restrict__ out,
                                        __half* __restrict__ out_split /*[rows][2 * C]*/, int rows, int C) {
  const size_t i4 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;          // 4 consecutive channels
  if (i4 * 4 >= size_t(rows) * C) return;
  const int r = int(i4 * 4 / C), c = int(i4 * 4 % C);
  const float4 x = *reinterpret_cast<const float4*>(a + i4 * 4);
  const float4 v = *reinterpret_cast<const float4*>(vec + c);
  const float y[4] = {x.x + v.x, x.y + v.y, x.z + v.z, x.w + v.w};
  *reinterpret_cast<float4*>(out + i4 * 4) = make_float4(y[0], y[1], y[2], y[3]);
  __half hi[4], lo[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) split_f16(y[t], hi[t], lo[t]);
  __half* o = out_split + size_t(r) * 2 * C + c;
  *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(o + C) = *reinterpret_cast<const uint2*>(lo);
}

// (3) image -> token cross attention: Q rows of `ldq` floats, prompt stride q_bstride (0 = shared); k,v [B][T][128].
//     One thread per (image token, head): the 8 threads of a token read its 512-byte Q row and write the two 256-byte
//     segments of its output row together, so every warp-wide access is four whole rows.  The prompt's k,v sit in shared
//     memory with a head pitch of 20 floats: the eight heads a warp touches at once then fall into eight different bank
//     groups (with the natural pitch of 16 they collide 4-way, which made the first per-head version 2.4x slower than the
//     one-thread-per-token kernel it replaced).  The result feeds the out_proj tensor-core GEMM and is written directly
//     as the split-fp16 operand [B*4096][hi(128) | lo(128)].
constexpr int I2T_HP = 20;                 // floats per (token, head) slice in shared memory
constexpr int I2T_TOK_PER_BLOCK = 128;     // image tokens per block (4 passes of 32 tokens x 8 heads)
__global__ void __launch_bounds__(256)
i2t_attn_kernel(const float* __restrict__ Q, int ldq, size_t q_bstride, const float* __restrict__ k, const float* __restrict__ v,
                __half* __restrict__ out_split /*[B][4096][384]*/, int T) {
  extern __shared__ __align__(16) float sm[];
  float* sk = sm;                          // [T][8][I2T_HP]
  float* sv = sm + T * 8 * I2T_HP;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < T * 128; i += blockDim.x) {
    const int t = i >> 7, c = i & 127;
    sk[(t * 8 + (c >> 4)) * I2T_HP + (c & 15)] = k[size_t(b) * T * 128 + i];
    sv[(t * 8 + (c >> 4)) * I2T_HP + (c & 15)] = v[size_t(b) * T * 128 + i];
  }
  __syncthreads();
  const int h = threadIdx.x & 7;
#pragma unroll 1
  for (int pass = 0; pass < I2T_TOK_PER_BLOCK / 32; ++pass) {
    const int token = blockIdx.x * I2T_TOK_PER_BLOCK + pass * 32 + (threadIdx.x >> 3);
    const float4* qr = reinterpret_cast<const float4*>(Q + size_t(b) * q_bstride + size_t(token) * ldq + h * 16);
    const float4 q0 = qr[0], q1 = qr[1], q2 = qr[2], q3 = qr[3];
    const float qv[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) {
      const float4* kk = reinterpret_cast<const float4*>(sk + (t * 8 + h) * I2T_HP);
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w = kk[j];
        a = fmaf(qv[4 * j], w.x, a); a = fmaf(qv[4 * j + 1], w.y, a); a = fmaf(qv[4 * j + 2], w.z, a); a = fmaf(qv[4 * j + 3], w.w, a);
      }
      m = fmaxf(m, a * 0.25f);
    }
    float l = 0.f;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    for (int t = 0; t < T; ++t) {
      const float4* kk = reinterpret_cast<const float4*>(sk + (t * 8 + h) * I2T_HP);
      const float4* vv = reinterpret_cast<const float4*>(sv + (t * 8 + h) * I2T_HP);
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w = kk[j];
        a = fmaf(qv[4 * j], w.x, a); a = fmaf(qv[4 * j + 1], w.y, a); a = fmaf(qv[4 * j + 2], w.z, a); a = fmaf(qv[4 * j + 3], w.w, a);
      }
      const float pexp = expf(a * 0.25f - m);
      l += pexp;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w = vv[j];
        acc[4 * j] = fmaf(pexp, w.x, acc[4 * j]); acc[4 * j + 1] = fmaf(pexp, w.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(pexp, w.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(pexp, w.w, acc[4 * j + 3]);
      }
    }
    const float inv = 1.0f / l;
    __half hi[16], lo[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) split_f16(acc[c] * inv, hi[c], lo[c]);
    __half* o = out_split + (size_t(b) * 4096 + token) * 256 + h * 16;
    reinterpret_cast<uint4*>(o)[0] = reinterpret_cast<uint4*>(hi)[0];
    reinterpret_cast<uint4*>(o)[1] = reinterpret_cast<uint4*>(hi)[1];
    reinterpret_cast<uint4*>(o + 128)[0] = reinterpret_cast<uint4*>(lo)[0];
    reinterpret_cast<uint4*>(o + 128)[1] = reinterpret_cast<uint4*>(lo)[1];
  }
}

// LayerNorm over 256 channels (decoder norm4, eps 1e-5) writing the fp32 result (optional) and its split-fp16
// form [hi(256) | lo(256)] for the following tensor-core GEMMs.  One warp per row.
__global__ void ln256_split_kernel(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float* __restrict__ out_f32, __half* __restrict__ out_split, int rows) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(in + size_t(row) * 256);
  const float4 a = src[lane], b = src[lane + 32];
  float s = (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 256.0f);
  float d[8] = {a.x - mean, a.y - mean, a.z - mean, a.w - mean, b.x - mean, b.y - mean, b.z - mean, b.w - mean};
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) q = fmaf(d[i], d[i], q);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / 256.0f) + eps);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int k = lane + 32 * half;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + k);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + k);
    const float y0 = d[4 * half] * rstd * g.x + be.x, y1 = d[4 * half + 1] * rstd * g.y + be.y;
    const float y2 = d[4 * half + 2] * rstd * g.z + be.z, y3 = d[4 * half + 3] * rstd * g.w + be.w;
    if (out_f32) reinterpret_cast<float4*>(out_f32 + size_t(row) * 256)[k] = make_float4(y0, y1, y2, y3);
    __half hi[4], lo[4];
    split_f16(y0, hi[0], lo[0]); split_f16(y1, hi[1], lo[1]); split_f16(y2, hi[2], lo[2]); split_f16(y3, hi[3], lo[3]);
    __half* o = out_split + size_t(row) * 512 + 4 * k;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<uint2*>(hi);
    *reinterpret_cast<uint2*>(o + 256) = *reinterpret_cast<uint2*>(lo);
  }
}


// LayerNorm2d(64, eps 1e-6) + GELU of output_upscaling (SA/modeling/mask_decoder.py:54-57, common.py:31-43) on the 64-channel
// groups base[row * ld + off + g * 64 + c], g < 4 (the ConvT1 columns of the fused projection GEMM); half a warp per group.
// The result leaves as the split-fp16 A operand [hi | lo] of the 3-term tensor-core ConvTranspose2 GEMM (which reads the hi block
// twice): row (token * 4 + group) of out, 128 halves (precision recipe).
__global__ void ln64_gelu_split_kernel(const float* __restrict__ base, int ld, int off, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, int rows, __half* __restrict__ out /*[rows*4][128]*/) {
  const int gidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;       // (row, group)
  const int l16 = threadIdx.x & 15;
  if (gidx >= rows * 4) return;
  const float4 a = *(reinterpret_cast<const float4*>(base + size_t(gidx >> 2) * ld + off + (gidx & 3) * 64) + l16);
  float s = (a.x + a.y) + (a.z + a.w);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 64.0f);
  const float d0 = a.x - mean, d1 = a.y - mean, d2 = a.z - mean, d3 = a.w - mean;
  float q = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q * (1.0f / 64.0f) + 1e-6f);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + l16);
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + l16);
  const float y[4] = {gelu_erf_f(g.x * (d0 * rstd) + be.x), gelu_erf_f(g.y * (d1 * rstd) + be.y),
                      gelu_erf_f(g.z * (d2 * rstd) + be.z), gelu_erf_f(g.w * (d3 * rstd) + be.w)};
  __half hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_f16(y[i], hi[i], lo[i]);
  __half* o = out + size_t(gidx) * 128 + l16 * 4;
  *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(o + 64) = *reinterpret_cast<const uint2*>(lo);
}
// weight [N][K] fp32 -> split-fp16 [N][3K] = scale * [hi | hi | lo], matching activations [hi | lo | hi] (stored [hi | lo], hi read twice)
__global__ void split_weight_kernel(const float* __restrict__ w, int N, int K, float scale, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n = i / K, k = i % K;
  __half hi, lo;
  split_f16(w[i] * scale, hi, lo);
  __half* o = out + size_t(n) * 3 * K;
  o[k] = hi;
  o[K + k] = hi;
  o[2 * K + k] = lo;
}



// ------------------------------------------------------------------------------------------------
// Fused post-processing (SA/modeling/sam.py:133-162 + SA/predictor.py:242-243 + the driver's painter,
// Generate Dataset/main_sam_hbox_semantic.py:162,195-199) for tiles whose input and original size are
// both 1024x1024 (second interpolate is the identity).  ATen's align_corners=False rule
// (ATen/native/UpSample.h area_pixel_compute_source_index): src = 0.25*(dst+0.5)-0.5 clamped at 0.
// Arithmetic is spelled with explicit round-to-nearest mul/add (no FMA contraction) in the order
//   r(y) = l0x*a[y][x0] + l1x*a[y][x1];   v = l0y*r(y0) + l1y*r(y1)
// so the thresholded result is bit-identical to the oracle's restatement.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index_x4(int d, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = __fsub_rn(__fmul_rn(0.25f, __fadd_rn(float(d), 0.5f)), 0.5f);
  s = fmaxf(s, 0.f);
  i0 = int(s);
  i1 = min(i0 + 1, in_size - 1);
  l1 = __fsub_rn(s, float(i0));
  l0 = __fsub_rn(1.0f, l1);
}
__device__ __forceinline__ float bilerp(const float* __restrict__ a, int y0, int y1, float ly0, float ly1, int x0, int x1,
                                        float lx0, float lx1) {
  const float r0 = __fadd_rn(__fmul_rn(lx0, a[y0 * 256 + x0]), __fmul_rn(lx1, a[y0 * 256 + x1]));
  const float r1 = __fadd_rn(__fmul_rn(lx0, a[y1 * 256 + x0]), __fmul_rn(lx1, a[y1 * 256 + x1]));
  return __fadd_rn(__fmul_rn(ly0, r0), __fmul_rn(ly1, r1));
}

// masks (u8 0/1 == torch.bool) and/or full-resolution logits for NB low-res maps
// Row-blocked: output rows 4k+2 .. 4k+5 interpolate between the SAME two low-res rows (k, k+1), so one thread owns a
// 4 x 4 output block, forms the 8 horizontal interpolations once and reuses them for its four rows (24 instead of 48
// interpolations per 16 pixels, 6 instead of 16 loads).  Row weights come from src_index_x4 per row and the order of the
// roundings is bilerp()'s, so the result is bit-identical to the one-row-per-thread form.  blockIdx.y = k + 1 (k = -1
// covers output rows 0 and 1, k = 255 rows 1022 and 1023).
__global__ void upsample4_threshold_kernel(const float* __restrict__ low /*[NB][256][256]*/, uint8_t* __restrict__ masks /*[NB][1024][1024] or null*/,
                                           float* __restrict__ logits /*or null*/, int NB) {
  const int x4 = blockIdx.x * blockDim.x + threadIdx.x;     // group of 4 output columns
  const int k = int(blockIdx.y) - 1, b = blockIdx.z;
  if (x4 >= 256) return;
  const float* a = low + size_t(b) * 65536;
  const int r0i = min(max(k, 0), 255), r1i = min(r0i + 1, 255);
  float h0[4], h1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int x0, x1;
    float lx0, lx1;
    src_index_x4(x4 * 4 + i, 256, x0, x1, lx0, lx1);
    h0[i] = __fadd_rn(__fmul_rn(lx0, a[r0i * 256 + x0]), __fmul_rn(lx1, a[r0i * 256 + x1]));
    h1[i] = __fadd_rn(__fmul_rn(lx0, a[r1i * 256 + x0]), __fmul_rn(lx1, a[r1i * 256 + x1]));
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = 4 * k + 2 + j;
    if (y < 0 || y > 1023) continue;
    int y0, y1;
    float ly0, ly1;
    src_index_x4(y, 256, y0, y1, ly0, ly1);                  // (y0, y1) == (r0i, r1i) wherever ly1 != 0
    uint32_t pk = 0;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = __fadd_rn(__fmul_rn(ly0, h0[i]), __fmul_rn(ly1, h1[i]));
      pk |= (v[i] > 0.0f ? 1u : 0u) << (8 * i);
    }
    const size_t o = (size_t(b) * 1024 + y) * 1024 + x4 * 4;
    if (masks) *reinterpret_cast<uint32_t*>(masks + o) = pk;
    if (logits) *reinterpret_cast<float4*>(logits + o) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// semantic label map: label of the highest-index prompt whose mask is true, else the existing canvas value
__global__ void upsample4_paint_kernel(const float* __restrict__ low /*[NB][256][256]*/, const int* __restrict__ labels, int NB,
                                       uint8_t* __restrict__ canvas /*[1024][1024]*/) {
  // same 4 x 4 row-blocking as upsample4_threshold_kernel; the scan over the boxes stops once all 16 pixels are decided
  const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = int(blockIdx.y) - 1;
  if (x4 >= 256) return;
  const int r0i = min(max(k, 0), 255), r1i = min(r0i + 1, 255);
  int x0[4], x1[4];
  float lx0[4], lx1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) src_index_x4(x4 * 4 + i, 256, x0[i], x1[i], lx0[i], lx1[i]);
  float ly0[4], ly1[4];
  uint32_t cur[4];
  uint32_t undecided = 0;                                   // bit (4 j + i): row j, column i still shows the canvas value
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = 4 * k + 2 + j;
    int y0, y1;
    src_index_x4(min(max(y, 0), 1023), 256, y0, y1, ly0[j], ly1[j]);
    cur[j] = 0;
    if (y >= 0 && y <= 1023) {
      cur[j] = *reinterpret_cast<uint32_t*>(canvas + size_t(y) * 1024 + x4 * 4);
      undecided |= 0xFu << (4 * j);
    }
  }
  const uint32_t rows_present = undecided;
  for (int b = NB - 1; b >= 0 && undecided; --b) {
    const float* a = low + size_t(b) * 65536;
    const uint32_t lab = uint32_t(labels[b]) & 0xFF;
    float h0[4], h1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h0[i] = __fadd_rn(__fmul_rn(lx0[i], a[r0i * 256 + x0[i]]), __fmul_rn(lx1[i], a[r0i * 256 + x1[i]]));
      h1[i] = __fadd_rn(__fmul_rn(lx0[i], a[r1i * 256 + x0[i]]), __fmul_rn(lx1[i], a[r1i * 256 + x1[i]]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if ((undecided >> (4 * j + i)) & 1) {
          if (__fadd_rn(__fmul_rn(ly0[j], h0[i]), __fmul_rn(ly1[j], h1[i])) > 0.0f) {
            cur[j] = (cur[j] & ~(0xFFu << (8 * i))) | (lab << (8 * i));
            undecided &= ~(1u << (4 * j + i));
          }
        }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if ((rows_present >> (4 * j)) & 1) *reinterpret_cast<uint32_t*>(canvas + size_t(4 * k + 2 + j) * 1024 + x4 * 4) = cur[j];
}

// ------------------------------------------------------------------------------------------------
// Painter reduce over bool masks of any size (the driver's loop, main_sam_hbox_semantic.py:195-199, for tiles whose
// original size is not 1024 x 1024: the masks come from the general-size postprocess).  One thread per 4 pixels scans
// the boxes from the last to the first and stops at the first mask that is set: "highest box index wins" (SURVEY.md F1).
// ------------------------------------------------------------------------------------------------
__global__ void paint_masks_kernel(const uint8_t* __restrict__ masks /*[NB][H*W]*/, const int* __restrict__ labels, int NB,
                                   size_t HW, uint8_t* __restrict__ canvas /*[H*W]*/) {
  const size_t i0 = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i0 >= HW) return;
  const int n = (HW - i0 >= 4) ? 4 : int(HW - i0);
  uint32_t undecided = (1u << n) - 1u;
  uint8_t cur[4];
  for (int i = 0; i < n; ++i) cur[i] = canvas[i0 + i];
  for (int b = NB - 1; b >= 0 && undecided; --b) {
    const uint8_t* m = masks + size_t(b) * HW + i0;
    const uint8_t lab = uint8_t(labels[b] & 0xFF);
    for (int i = 0; i < n; ++i)
      if (((undecided >> i) & 1) && m[i]) { cur[i] = lab; undecided &= ~(1u << i); }
  }
  for (int i = 0; i < n; ++i) canvas[i0 + i] = cur[i];
}

// ------------------------------------------------------------------------------------------------
// Pillow-exact 8-bit bilinear image resize (ResizeLongestSide.apply_image, SA/utils/transforms.py:26-31 ->
// PIL.Image.resize(BILINEAR) -> libImaging/Resample.c): two separable passes over uint8 HWC data with 22-bit fixed-point
// taps; each pass accumulates 2^21 + sum(pixel * tap) in int32, shifts by 22 and clamps to [0, 255]; the horizontal pass
// runs first and its uint8 result feeds the vertical one.  Tap tables (first tap, count, weights) come from the host
// (engine.cu pil_coeffs, double arithmetic as in Resample.c precompute_coeffs / normalize_coeffs_8bpc).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t pil_clip8(int acc) { return uint8_t(min(max(acc >> 22, 0), 255)); }

__global__ void pil_resize_h_kernel(const uint8_t* __restrict__ src /*[H][W][3]*/, int W, uint8_t* __restrict__ dst /*[H][W1][3]*/, int W1,
                                    const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const int x1 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x1 >= W1) return;
  const int lo = bounds[2 * x1], n = bounds[2 * x1 + 1];
  const int* k = kk + size_t(x1) * ksize;
  const uint8_t* p = src + (size_t(y) * W + lo) * 3;
  int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
  for (int i = 0; i < n; ++i) {
    const int kv = k[i];
    a0 += int(p[3 * i]) * kv;
    a1 += int(p[3 * i + 1]) * kv;
    a2 += int(p[3 * i + 2]) * kv;
  }
  uint8_t* o = dst + (size_t(y) * W1 + x1) * 3;
  o[0] = pil_clip8(a0); o[1] = pil_clip8(a1); o[2] = pil_clip8(a2);
}

// vertical pass: one thread per byte of an output row (x * 3 + c), consecutive threads read consecutive bytes
__global__ void pil_resize_v_kernel(const uint8_t* __restrict__ src /*[H][row]*/, int row_bytes, uint8_t* __restrict__ dst /*[H1][row]*/,
                                    const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x, y1 = blockIdx.y;
  if (e >= row_bytes) return;
  const int lo = bounds[2 * y1], n = bounds[2 * y1 + 1];
  const int* k = kk + size_t(y1) * ksize;
  int a = 1 << 21;
  for (int i = 0; i < n; ++i) a += int(src[size_t(lo + i) * row_bytes + e]) * k[i];
  dst[size_t(y1) * row_bytes + e] = pil_clip8(a);
}

// general bilinear resize (align_corners=False, ATen scale = in/out) of a cropped source view
__global__ void bilinear_resize_kernel(const float* __restrict__ in, int in_ld, int in_plane, int in_h, int in_w,
                                       float* __restrict__ out_f, uint8_t* __restrict__ out_mask, int out_h, int out_w, int NB) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= out_w) return;
  const float sy = float(in_h) / float(out_h), sx = float(in_w) / float(out_w);
  float fy = fmaxf(__fsub_rn(__fmul_rn(sy, __fadd_rn(float(y), 0.5f)), 0.5f), 0.f);
  float fx = fmaxf(__fsub_rn(__fmul_rn(sx, __fadd_rn(float(x), 0.5f)), 0.5f), 0.f);
  const int y0 = int(fy), x0 = int(fx);
  const int y1 = y0 + (y0 < in_h - 1 ? 1 : 0), x1 = x0 + (x0 < in_w - 1 ? 1 : 0);
  const float ly1 = __fsub_rn(fy, float(y0)), lx1 = __fsub_rn(fx, float(x0));
  const float ly0 = __fsub_rn(1.0f, ly1), lx0 = __fsub_rn(1.0f, lx1);
  const float* a = in + size_t(b) * in_plane;
  const float r0 = __fadd_rn(__fmul_rn(lx0, a[y0 * in_ld + x0]), __fmul_rn(lx1, a[y0 * in_ld + x1]));
  const float r1 = __fadd_rn(__fmul_rn(lx0, a[y1 * in_ld + x0]), __fmul_rn(lx1, a[y1 * in_ld + x1]));
  const float v = __fadd_rn(__fmul_rn(ly0, r0), __fmul_rn(ly1, r1));
  const size_t o = (size_t(b) * out_h + y) * out_w + x;
  if (out_f) out_f[o] = v;
  if (out_mask) out_mask[o] = v > 0.0f ? 1 : 0;
}

}  // namespace samrs
