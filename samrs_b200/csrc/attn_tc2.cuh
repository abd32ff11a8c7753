// tcgen05 attention, second generation: two softmax warpgroups per CTA ping-pong on one tensor core.
//
// Same mathematics as attn_tc.cuh (SA/modeling/image_encoder.py:224-240, 325-361; zero-filled padded keys, K/V bias
// folded away), different schedule.  The first kernel ran   load -> S = QK^T -> softmax -> P (smem) -> O = PV -> rescale
// strictly one after the other per query tile and reached ~0.1-0.18 PFLOP/s (profiles/r01_*).  Here:
//   * a CTA owns TWO query tiles that share every K/V tile (windowed: the two halves of one 14x14 window and head;
//     global: two adjacent 128-query tiles).  Warps 4-7 / 8-11 are the softmax warpgroups of tile 0 / 1; while one
//     group is in its exp2 pass the tensor core works for the other.
//   * P never touches shared memory: each softmax thread packs its row to fp16 and tcgen05.st's it over the first half
//     of its own S columns; the PV MMA then reads A = P from tensor memory (tcgen05.mma with a TMEM A operand).
//   * O accumulates in tensor memory across key tiles.  The running max used for scaling (m_ref) is only raised when
//     a row's tile max exceeds it by more than 2^8 (P stays <= 256, far inside fp16); only then is O rescaled in TMEM.
//     The exact result is independent of m_ref because O and the row sum l carry the same factor.
//   * the decomposed rel-pos bias is added by the tensor core, not by the softmax threads: bias[q,k] = R[q,:] . E[k,:]
//     with R the query's rel-pos terms (fp16, written once per unit into TMEM by the row's thread) and E a constant
//     one-hot matrix in shared memory (global: E[k][kw'] = [k mod 64 == kw'], the two relh terms of a key tile stay
//     scalar; windowed: E[k][kh'] = [k div 14 == kh'], E[k][14+kw'] = [k mod 14 == kw']).  One extra MMA per key tile
//     (K = 64 / 32) replaces a load + add per score, which halves the instruction count of both softmax passes.
//   * windowed blocks compute the rel-pos terms themselves: G = Q . T^T with T = log2e [rel_pos_h ; rel_pos_w] (27 + 27
//     rows, resident in shared memory) is one more N = 64 MMA per query tile into the still-unused S columns; the row's
//     thread reads its 54 values, selects the 14 + 14 its query position needs (a 4-stage barrel shift: no dynamic
//     register indexing, no memory), scales them and writes them as the fp16 operand R.  The separate head-batched
//     rel-pos GEMM and its [heads][4096][64] table (28 launches, 8 MB written and gathered per layer) are gone; global
//     blocks (4 of 32) keep the GEMM + gather, whose table rows depend on the key tile.
// TMEM (512 columns): warpgroup w owns columns [256w, 256w+256): S at +0 (208 or 128 fp32 columns), P aliased on
// S's first half, O at +112 (windowed: inside the dead upper half of S) or +128 (global), R at +208; windowed G at +0
// (read and replaced by R before the S MMA is issued).
#pragma once
#include <type_traits>

#include "attn_tc.cuh"

namespace samrs {

template <int HD, int BX, int QBY, int KBY, int NKT>
struct Attn2Cfg {
  static constexpr int NATOM = (HD + 63) / 64;
  static constexpr int KR = BX * KBY;
  static constexpr int SN = (KR + 15) / 16 * 16;
  static constexpr int QR = BX * QBY;
  static constexpr int KV_STAGES = (NKT > 1) ? 2 : 1;
  static constexpr int Q_TILE_BYTES = NATOM * 128 * 128;
  static constexpr int KV_ATOM_BYTES = SN * 128;
  static constexpr int KV_BYTES = NATOM * KV_ATOM_BYTES;
  static constexpr int O_OFF = (NKT > 1) ? 128 : 112;
  static constexpr int R_OFF = 208;                          // TMEM column of the packed fp16 rel-pos operand R
  static constexpr int RK = (NKT > 1) ? 64 : 32;             // K extent of the bias MMA (kw' | kh',kw')
  static constexpr int E_BYTES = SN * 128;                   // one-hot matrix E: SN key rows x 64 fp16 (SW128 atom rows)
  static constexpr int T_ROWS = 64;                          // windowed: rel-pos table rows (27 + 27, zero padded)
  static constexpr int T_BYTES = (NKT == 1) ? NATOM * T_ROWS * 128 : 0;
  static constexpr int kSmemBytes = 2 * Q_TILE_BYTES + 2 * KV_STAGES * KV_BYTES + E_BYTES + T_BYTES + 1024 + 256;
  static_assert(SN % 16 == 0 && SN <= 208, "bad S tile");
  static_assert(O_OFF >= SN / 2 && O_OFF + HD <= 256, "O must not overlap P");
};

// compile-time loop: guarantees that every chunk offset is a constant, so relh[] / relw[] stay in registers
template <int I, int N, int S, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + S, N, S>(f);
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / integer pipes (Cody-Waite split + degree-4 polynomial on [-0.5, 0.5], relative error 3.6e-6, far
// below the fp16 rounding of P).  The MUFU unit retires one warp-wide ex2 per 8 clocks per SM sub-partition and the
// windowed softmax was bound by it (and by the fp32->fp16 packs that share it): every third exponential of the
// windowed kernel goes here.  The global kernel is issue-bound instead: with every 3rd / 4th exponential on the FMA pipe it
// ran 13 % / 7 % slower (tools/gpu_ab.sh, -DSAMRS_GLOBAL_POLY_DIV=3|4), so it keeps MUFU only.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float r = x + 12582912.0f;                 // 1.5 * 2^23: the integer part of x lands in the low mantissa bits
  const float f = x - (r - 12582912.0f);           // fractional part in [-0.5, 0.5]
  float q = fmaf(f, 0.009676037f, 0.055922036f);
  q = fmaf(q, f, 0.24022107f);
  q = fmaf(q, f, 0.69312103f);
  q = fmaf(q, f, 1.0000001f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(r) << 23));
}
// DIV = 0: MUFU only; DIV = n: every n-th exponential goes to the FMA pipe
#ifndef SAMRS_GLOBAL_POLY_DIV
#define SAMRS_GLOBAL_POLY_DIV 0
#endif
template <int I, int DIV>
__device__ __forceinline__ float ex2_mixed(float x) {
  if constexpr (DIV > 0 && I % (DIV > 0 ? DIV : 1) == (DIV > 0 ? DIV : 1) - 1) return ex2_poly(x);
  else return ex2_approx(x);
}

template <int HD, int BX, int QBY, int KBY, int NKT>
__global__ void __launch_bounds__(384, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmT /*windowed: rel-pos table [64][HD] fp16*/, const AttnParams p) {
  using C = Attn2Cfg<HD, BX, QBY, KBY, NKT>;
  constexpr int NATOM = C::NATOM, SN = C::SN, KR = C::KR, QR = C::QR, ST = C::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2 tiles][NATOM atoms][128 rows x 128 B]
  uint8_t* sK = sQ + 2 * C::Q_TILE_BYTES;
  uint8_t* sV = sK + ST * C::KV_BYTES;
  uint8_t* sE = sV + ST * C::KV_BYTES;
  uint8_t* sT = sE + C::E_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + C::T_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;        // [2]
  uint64_t* k_empty = bars + 4;
  uint64_t* v_full = bars + 6;
  uint64_t* v_empty = bars + 8;
  uint64_t* s_full = bars + 10;       // [2] per warpgroup
  uint64_t* p_full = bars + 12;       // [2] per warpgroup, 128 arrivals
  uint64_t* o_full = bars + 14;       // [2] per warpgroup: last PV of the unit done
  uint64_t* o_free = bars + 16;       // [2] per warpgroup, 128 arrivals: O has been read, S/O region reusable
  uint64_t* r_full = bars + 18;       // [2] per warpgroup, 128 arrivals: this unit's R operand is in TMEM
  uint64_t* g_full = bars + 20;       // [2] per warpgroup (windowed): G = Q T^T is in TMEM
  uint64_t* t_full = bars + 22;       // windowed: the rel-pos table has landed in shared memory (once per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_units = p.num_qtiles * p.heads;         // num_qtiles = query-tile PAIRS here

  for (int i = threadIdx.x; i < (2 * C::Q_TILE_BYTES + 2 * ST * C::KV_BYTES + C::E_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // one-hot selector E (K-major SW128 rows of 64 fp16): constant for the whole kernel
  for (int k = threadIdx.x; k < KR; k += blockDim.x) {
    const __half one = __float2half_rn(1.0f);
    if (NKT > 1) {
      const int c = k % 64;
      *reinterpret_cast<__half*>(sE + sw128_offset(k, c >> 3) + (c & 7) * 2) = one;
    } else {
      const int c0 = k / BX, c1 = KBY + k % BX;
      *reinterpret_cast<__half*>(sE + sw128_offset(k, c0 >> 3) + (c0 & 7) * 2) = one;
      *reinterpret_cast<__half*>(sE + sw128_offset(k, c1 >> 3) + (c1 & 7) * 2) = one;
    }
  }
  fence_proxy_async_smem();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    if (NKT == 1) tma_prefetch_desc(&tmT);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 128);
      mbar_init(&r_full[i], 128);
      mbar_init(&g_full[i], 1);
    }
    mbar_init(t_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // register re-allocation between warpgroups: the control warpgroup (TMA / MMA issue) keeps 96 registers per thread,
  // the two softmax warpgroups get 200 (128*96 + 256*200 = 63488 <= 384*168, the pool the CTA was launched with)
  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
  pdl_trigger();
  pdl_wait();                                           // zero fill, selector matrix, barriers and TMEM came before

  auto unit_coords = [&](int unit, int& head, int& qy0, int& qy1, int& x0, int& ky0) {
    head = unit % p.heads;
    const int u = unit / p.heads;
    if (NKT == 1) {              // windowed: u = window, the two tiles are its upper / lower 7 rows
      x0 = (u % 5) * BX;
      ky0 = (u / 5) * KBY;
      qy0 = ky0;
      qy1 = ky0 + QBY;
    } else {                     // global: u = pair of 128-query tiles
      x0 = 0;
      ky0 = 0;
      qy0 = u * 2 * QBY;
      qy1 = qy0 + QBY;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // warp-uniform loop, one elected lane issues (a divergent `lane == 0` branch makes the compiler wrap every TMA / MMA
    // instruction in an R2UR.BROADCAST waterfall loop)
    {
      uint32_t qph = 0, kst = 0, kph = 0, vst = 0, vph = 0;
      int ui = 0;
      if (NKT == 1 && int(blockIdx.x) < num_units) {
        if (elect_one()) {
          mbar_expect_tx(t_full, C::T_BYTES);
#pragma unroll
          for (int a = 0; a < NATOM; ++a) tma_load_2d(sT + a * C::T_ROWS * 128, &tmT, t_full, a * 64, 0);
        }
        __syncwarp();
      }
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++ui) {
        int head, qy0, qy1, x0, ky0;
        unit_coords(unit, head, qy0, qy1, x0, ky0);
        attn_dbg(p, 64 * ui + 0);
        mbar_wait(q_empty, qph ^ 1);
        attn_dbg(p, 64 * ui + 1);
        if (elect_one()) {
          mbar_expect_tx(q_full, 2 * NATOM * QR * 128);
#pragma unroll
          for (int a = 0; a < NATOM; ++a) {
            tma_load_3d(sQ + a * 128 * 128, &tmQ, q_full, head * HD + a * 64, x0, qy0);
            tma_load_3d(sQ + C::Q_TILE_BYTES + a * 128 * 128, &tmQ, q_full, head * HD + a * 64, x0, qy1);
          }
        }
        __syncwarp();
        qph ^= 1;
        for (int j = 0; j < NKT; ++j) {
          mbar_wait(&k_empty[kst], kph ^ 1);
          if (j == 0) attn_dbg(p, 64 * ui + 2);
          if (elect_one()) {
            mbar_expect_tx(&k_full[kst], NATOM * KR * 128);
#pragma unroll
            for (int a = 0; a < NATOM; ++a)
              tma_load_3d(sK + kst * C::KV_BYTES + a * C::KV_ATOM_BYTES, &tmKV, &k_full[kst], p.D + head * HD + a * 64, x0, ky0 + j * KBY);
          }
          __syncwarp();
          if (++kst == ST) { kst = 0; kph ^= 1; }
          mbar_wait(&v_empty[vst], vph ^ 1);
          if (j == 0) attn_dbg(p, 64 * ui + 3);
          if (elect_one()) {
            mbar_expect_tx(&v_full[vst], NATOM * KR * 128);
#pragma unroll
            for (int a = 0; a < NATOM; ++a)
              tma_load_3d(sV + vst * C::KV_BYTES + a * C::KV_ATOM_BYTES, &tmKV, &v_full[vst], 2 * p.D + head * HD + a * 64, x0, ky0 + j * KBY);
          }
          __syncwarp();
          if (++vst == ST) { vst = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, SN, 0, 0);
      constexpr uint32_t idesc_o64 = umma_idesc_f16(128, 64, 0, 1);
      constexpr uint32_t idesc_o16 = umma_idesc_f16(128, 16, 0, 1);
      constexpr uint32_t idesc_o80 = umma_idesc_f16(128, 80, 0, 1);
      uint32_t qph = 0, kst = 0, kph = 0, vst = 0, vph = 0, pph[2] = {0, 0}, fph[2] = {0, 0}, rph = 0;
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aE = smem_u32(sE), aT = smem_u32(sT);
      constexpr uint32_t idesc_g = umma_idesc_f16(128, 64, 0, 0);
      auto issue_g = [&](int w) {                       // windowed: G_w = Q_w T^T into the (still unused) first S columns
        const uint32_t d = tmem_base + uint32_t(w * 256);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint32_t a = aQ + w * C::Q_TILE_BYTES + (k / 4) * (128 * 128) + (k % 4) * 32;
            const uint32_t b = aT + (k / 4) * (C::T_ROWS * 128) + (k % 4) * 32;
            tc_mma_f16(d, umma_desc_sw128(a), umma_desc_sw128(b), idesc_g, k != 0);
          }
          tc_commit(&g_full[w]);
        }
        __syncwarp();
      };
      if (NKT == 1 && int(blockIdx.x) < num_units) mbar_wait(t_full, 0);
      auto issue_s = [&](int w) {                       // S_w = Q_w K^T into warpgroup w's columns
        const uint32_t d = tmem_base + uint32_t(w * 256);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint32_t a = aQ + w * C::Q_TILE_BYTES + (k / 4) * (128 * 128) + (k % 4) * 32;
            const uint32_t b = aK + kst * C::KV_BYTES + (k / 4) * C::KV_ATOM_BYTES + (k % 4) * 32;
            tc_mma_f16(d, umma_desc_sw128(a), umma_desc_sw128(b), idesc_s, k != 0);
          }
          // + rel-pos bias: S += R_w (TMEM, fp16) . E^T (one-hot, smem)
#pragma unroll
          for (int k = 0; k < C::RK / 16; ++k)
            tc_mma_f16_ts(d, tmem_base + uint32_t(w * 256 + C::R_OFF + k * 8), umma_desc_sw128(aE + k * 32), idesc_s, 1u);
          tc_commit(&s_full[w]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int w, int j) {               // O_w (+)= P_w V_j, A = P from tensor memory
        const uint32_t d = tmem_base + uint32_t(w * 256 + C::O_OFF);
        const uint32_t pa = tmem_base + uint32_t(w * 256);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < SN / 16; ++k) {
            const uint32_t b = aV + vst * C::KV_BYTES + k * 2048;
            const uint32_t acc = (j > 0 || k > 0) ? 1u : 0u;
            if (NATOM == 2 && !p.pv_split) {
              // head dim 80 = one 64-wide swizzle atom + 16 columns of the next: a single N=80 MMA whose descriptor strides
              // over the two atoms (LBO), instead of an N=64 and an N=16 MMA - the issue count, not the flops, bounds PV
              tc_mma_f16_ts(d, pa + uint32_t(k * 8), umma_desc_sw128_mn(b, C::KV_ATOM_BYTES), idesc_o80, acc);
            } else {
              tc_mma_f16_ts(d, pa + uint32_t(k * 8), umma_desc_sw128(b), idesc_o64, acc);
              if (NATOM == 2) tc_mma_f16_ts(d + 64, pa + uint32_t(k * 8), umma_desc_sw128(b + C::KV_ATOM_BYTES), idesc_o16, acc);
            }
          }
        }
        __syncwarp();
      };
      int ui = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++ui) {
        attn_dbg(p, 64 * ui + 8);
        mbar_wait(q_full, qph);
        qph ^= 1;
        attn_dbg(p, 64 * ui + 9);
        if (NKT == 1) {
          tc_fence_after();
          for (int w = 0; w < 2; ++w) {                 // the warpgroup has read the previous unit's O (aliases S and G)
            mbar_wait(&o_free[w], fph[w] ^ 1);
            fph[w] ^= 1;
            tc_fence_after();
            issue_g(w);
          }
        }
        mbar_wait(&k_full[kst], kph);
        tc_fence_after();
        attn_dbg(p, 64 * ui + 10);
        for (int w = 0; w < 2; ++w) {                   // the warpgroup has read the previous unit's O (aliases S)
          if (NKT > 1) {
            mbar_wait(&o_free[w], fph[w] ^ 1);
            fph[w] ^= 1;
          }
          mbar_wait(&r_full[w], rph);                   // ... and has stored this unit's R operand
          tc_fence_after();
          issue_s(w);
          attn_dbg(p, 64 * ui + 11 + w);
        }
        rph ^= 1;
        if (elect_one()) tc_commit(&k_empty[kst]);
        if (++kst == ST) { kst = 0; kph ^= 1; }
        if ((NKT == 1) && elect_one()) tc_commit(q_empty);
        for (int j = 0; j < NKT; ++j) {
          const bool more = (j + 1 < NKT);
          mbar_wait(&v_full[vst], vph);
          if (j == 0) attn_dbg(p, 64 * ui + 13);
          if (more) mbar_wait(&k_full[kst], kph);
          for (int w = 0; w < 2; ++w) {
            mbar_wait(&p_full[w], pph[w]);
            pph[w] ^= 1;
            tc_fence_after();
            if (j == 0) attn_dbg(p, 64 * ui + 14 + w);
            issue_pv(w, j);
            if ((!more) && elect_one()) tc_commit(&o_full[w]);
            if (more) issue_s(w);                       // S_w of the next key tile overwrites the P it just consumed
          }
          if (elect_one()) tc_commit(&v_empty[vst]);
          if (++vst == ST) { vst = 0; vph ^= 1; }
          if (more) {
            if (elect_one()) tc_commit(&k_empty[kst]);
            if (++kst == ST) { kst = 0; kph ^= 1; }
            if ((j + 2 == NKT) && elect_one()) tc_commit(q_empty);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax warpgroups (warps 4-7 and 8-11)
    const int w = (warp - 4) >> 2;                      // warpgroup = query tile
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t wg_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(w * 256);
    constexpr int NP = (BX == 64) ? 256 : 64;
    constexpr int SS = BX;
    uint32_t sph = 0, oph = 0, gph = 0;
    int ui = 0;
    const bool tr = (r == 0);
    // rel-pos terms of this thread's query row for unit `un`, divided by the score scale (the accumulator is scaled
    // afterwards: t = scale_log2e * (q.k + R.E)) and packed to fp16 pairs = one row of the TMEM operand R
    uint32_t rk_next[C::RK / 2];
    auto load_rk = [&](int un) {
      int head, qy0, qy1, x0, ky0;
      unit_coords(un, head, qy0, qy1, x0, ky0);
      const int qy = w ? qy1 : qy0;
      const int ty = qy + r / BX, tx = x0 + r % BX;
      const bool valid = (r < QR) && ty < 64 && tx < 64;
      const int qw = r % BX;
      if constexpr (NKT > 1) {
        // the rel-pos GEMM's epilogue wrote fp16(G / scale_log2e), i.e. the entries of the bias operand R themselves
        const __half* rrow = p.rel16 + (size_t(head) * 4096 + (valid ? ty * 64 + tx : 0)) * NP;
        const __half zero = __float2half_rn(0.f);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const __half a0 = valid ? __ldg(rrow + (2 * SS - 1) + qw + (SS - 1) - 2 * i) : zero;
          const __half a1 = valid ? __ldg(rrow + (2 * SS - 1) + qw + (SS - 1) - (2 * i + 1)) : zero;
          __half2 h = __halves2half2(a0, a1);
          rk_next[i] = *reinterpret_cast<uint32_t*>(&h);
        }
      }
    };
    if (NKT > 1 && int(blockIdx.x) < num_units) load_rk(blockIdx.x);
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++ui) {
      int head, qy0, qy1, x0, ky0;
      unit_coords(unit, head, qy0, qy1, x0, ky0);
      if (tr) attn_dbg(p, 64 * ui + 20 + 10 * w);
      const int qy = w ? qy1 : qy0;
      const int ty = qy + r / BX, tx = x0 + r % BX;
      const bool valid = (r < QR) && ty < 64 && tx < 64;
      const int token = ty * 64 + tx;
      const __half* relrow = p.rel16 + (size_t(head) * 4096 + (valid ? token : 0)) * NP;
      const int qh = qy - ky0 + r / BX;
      if constexpr (NKT > 1) {
        // this row's packed rel-pos operand R was prefetched during the previous unit (rk_next): store it to tensor memory
        tmem_st32(wg_addr + C::R_OFF, rk_next);
      } else {
        // windowed: G = q . log2e [rel_pos_h ; rel_pos_w]^T (54 values) is in this row's first 64 S columns.  The query at
        // window position (qh, qw) needs rel_h[kh] = G[qh - kh + 13] and rel_w[kw] = G[27 + qw - kw + 13]: shift each
        // 27-entry half down by qh / qw with a 4-stage barrel shifter (warp-divergent selects, no indexing), reverse,
        // scale by 1 / (scale log2e) (the accumulator is multiplied by scale log2e afterwards) and pack to fp16.
        mbar_wait(&g_full[w], gph);
        gph ^= 1;
        tc_fence_after();
        uint32_t rk[16];
        rk[14] = 0u;
        rk[15] = 0u;
        // one 27-entry half at a time (32 live registers): columns [0,27) = rel_pos_h terms, [27,54) = rel_pos_w terms
        static_for<0, 2, 1>([&](auto hc) {
          constexpr int half = decltype(hc)::value;
          uint32_t gv[32];
          tmem_ld32(wg_addr + half * 27, gv);
          tc_wait_ld();
          const int sft = half ? (r % BX) : qh;
#pragma unroll
          for (int b = 8; b >= 1; b >>= 1) {             // largest shift first: every stage reads entries that the earlier
            const bool m = (sft & b) != 0;               // (larger) stages have already brought into range
#pragma unroll
            for (int j = 0; j < 14 + 8; ++j)             // after all stages only entries 0..13 are used
              if (j + b < 27) gv[j] = m ? gv[j + b] : gv[j];
              else gv[j] = m ? 0u : gv[j];
          }
#pragma unroll
          for (int i = 0; i < 7; ++i) {                  // operand columns half*14 + {2i, 2i+1} <- shifted entries 13-2i, 12-2i
            __half2 h = __floats2half2_rn(__uint_as_float(gv[13 - 2 * i]) * p.rel_scale, __uint_as_float(gv[12 - 2 * i]) * p.rel_scale);
            rk[half * 7 + i] = *reinterpret_cast<uint32_t*>(&h);
          }
        });
        tmem_st16(wg_addr + C::R_OFF, rk);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&r_full[w]);
      float m_ref = 0.f, l_run = 0.f;

      for (int j = 0; j < NKT; ++j) {
        // global blocks: the two key rows of this tile share rel_h terms that stay scalar; windowed: all bias is in the MMA
        constexpr int NG = (NKT > 1) ? KBY : 1;
        float relh[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) relh[i] = (NKT > 1 && valid) ? __half2float(__ldg(relrow + qh + (SS - 1) - (j * KBY + i))) * p.scale_log2e : 0.f;
        if (tr && j == 0) attn_dbg(p, 64 * ui + 21 + 10 * w);
        mbar_wait(&s_full[w], sph);
        sph ^= 1;
        tc_fence_after();
        if (tr && j == 0) attn_dbg(p, 64 * ui + 22 + 10 * w);
        // pass 1: row max (log2 domain).  The accumulator already holds q.k + bias/scale: one FMNMX per score, the
        // scale (> 0) and the scalar rel_h term are applied to the group maxima
        float m_g[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) m_g[i] = -INFINITY;
        static_for<0, SN, 32>([&](auto c0c) {
          constexpr int c0 = decltype(c0c)::value;
          constexpr int W = (SN - c0 >= 32) ? 32 : 16;
          uint32_t v[W];
          if constexpr (W == 32) tmem_ld32(wg_addr + c0, v); else tmem_ld16(wg_addr + c0, v);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < W; ++i) {
            const int c = c0 + i;
            if (c < KR) m_g[NKT > 1 ? c / BX : 0] = fmaxf(m_g[NKT > 1 ? c / BX : 0], __uint_as_float(v[i]));
          }
        });
        float m_tile = -INFINITY;
#pragma unroll
        for (int i = 0; i < NG; ++i) m_tile = fmaxf(m_tile, fmaf(m_g[i], p.scale_log2e, relh[i]));
        if (tr && j == 0) attn_dbg(p, 64 * ui + 23 + 10 * w);
        if (j == 0) {
          m_ref = m_tile;
        } else {
          // lazy rescale: s_full(j) also certifies that PV_{j-1} has completed, so O is quiescent here
          const bool need = m_tile > m_ref + 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            const float m_new = need ? m_tile : m_ref;
            const float alpha = ex2_approx(m_ref - m_new);
            static_for<0, HD, 32>([&](auto c0c) {
              constexpr int c0 = decltype(c0c)::value;
              if constexpr (HD - c0 >= 32) {
                uint32_t v[32];
                tmem_ld32(wg_addr + C::O_OFF + c0, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                tmem_st32(wg_addr + C::O_OFF + c0, v);
              } else {
                uint32_t v[16];
                tmem_ld16(wg_addr + C::O_OFF + c0, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                tmem_st16(wg_addr + C::O_OFF + c0, v);
              }
            });
            l_run *= alpha;
            m_ref = m_new;
          }
        }
        // pass 2: P = exp2(scale * acc + rel_h - m_ref) -> fp16 pairs -> tensor memory (over the first half of S)
        float dh[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) dh[i] = relh[i] - m_ref;
        float l_tile = 0.f;
        static_for<0, SN, 32>([&](auto c0c) {
          constexpr int c0 = decltype(c0c)::value;
          constexpr int W = (SN - c0 >= 32) ? 32 : 16;
          uint32_t v[W];
          if constexpr (W == 32) tmem_ld32(wg_addr + c0, v); else tmem_ld16(wg_addr + c0, v);
          tc_wait_ld();
          uint32_t pk[W / 2];
          static_for<0, W, 2>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr int c = c0 + i;
            float e0 = 0.f, e1 = 0.f;
            if constexpr (c < KR) e0 = ex2_mixed<i, (NKT == 1 ? 3 : SAMRS_GLOBAL_POLY_DIV)>(fmaf(__uint_as_float(v[i]), p.scale_log2e, dh[NKT > 1 ? c / BX : 0]));
            if constexpr (c + 1 < KR) e1 = ex2_mixed<i + 1, (NKT == 1 ? 3 : SAMRS_GLOBAL_POLY_DIV)>(fmaf(__uint_as_float(v[i + 1]), p.scale_log2e, dh[NKT > 1 ? (c + 1) / BX : 0]));
            l_tile += e0 + e1;
            __half2 h = __floats2half2_rn(e0, e1);
            pk[i / 2] = *reinterpret_cast<uint32_t*>(&h);
          });
          if constexpr (W == 32) tmem_st16(wg_addr + c0 / 2, pk); else tmem_st8(wg_addr + c0 / 2, pk);
        });
        l_run += l_tile;
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[w]);
        if (tr && j == 0) attn_dbg(p, 64 * ui + 24 + 10 * w);
      }
      if (NKT > 1 && unit + int(gridDim.x) < num_units) load_rk(unit + gridDim.x);    // prefetch the next unit's R while PV runs
      // O complete: normalise and write the fp16 output row
      mbar_wait(&o_full[w], oph);
      oph ^= 1;
      tc_fence_after();
      if (tr) attn_dbg(p, 64 * ui + 25 + 10 * w);
      const float inv = 1.0f / l_run;
      __half* o = p.out + size_t(valid ? token : 0) * p.D + head * HD;
      static_for<0, HD, 32>([&](auto c0c) {
        constexpr int c0 = decltype(c0c)::value;
        if constexpr (HD - c0 >= 32) {
          uint32_t v[32];
          tmem_ld32(wg_addr + C::O_OFF + c0, v);
          tc_wait_ld();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              __half2 h0 = __floats2half2_rn(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
              __half2 h1 = __floats2half2_rn(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
              __half2 h2 = __floats2half2_rn(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
              __half2 h3 = __floats2half2_rn(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(o + c0 + i) = pk;
            }
          }
        } else {
          uint32_t v[16];
          tmem_ld16(wg_addr + C::O_OFF + c0, v);
          tc_wait_ld();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              __half2 h0 = __floats2half2_rn(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
              __half2 h1 = __floats2half2_rn(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
              __half2 h2 = __floats2half2_rn(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
              __half2 h3 = __floats2half2_rn(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(o + c0 + i) = pk;
            }
          }
        }
      });
      tc_fence_before();
      mbar_arrive(&o_free[w]);
      if (tr) attn_dbg(p, 64 * ui + 26 + 10 * w);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace samrs
