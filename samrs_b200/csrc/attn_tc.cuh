// tcgen05 attention with decomposed relative-position bias for the SAM image encoder.
//
// One kernel template serves both block kinds of `Attention.forward`
// (SA/modeling/image_encoder.py:224-240) + `add_decomposed_rel_pos` (:325-361):
//   * 14x14 windowed blocks: one query tile = half a window (7 rows x 14 tokens = 98 queries),
//     one key tile = the whole window (196 keys, zero-padded tokens included);
//   * global blocks: query tile = 2 grid rows (128 queries), 32 key tiles of 2 grid rows (128 keys),
//     streaming (online) softmax so the 4096x4096 score matrix is never materialised.
// Tokens are addressed through one 3-D TMA tensor map over the qkv activation viewed as
// [y=64][x=64][3*D] fp16; out-of-grid (padded) tokens are zero-filled by TMA.  Because the K bias only
// shifts every score of a row by the same amount and the V bias adds bv to every output row
// (softmax rows sum to 1), the qkv GEMM omits both and the V bias is folded into the proj bias;
// zero-filled padded keys then reproduce the reference's "padded k,v = qkv bias" semantics exactly.
//
// Per CTA (persistent over (query tile, head) items), 256 threads:
//   warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 softmax (1 query row / thread).
//   S = Q K^T    : tcgen05.mma M128 x N(SN) x K16, Q and K tiles K-major SW128 in smem, fp32 S in TMEM
//   softmax      : tcgen05.ld S -> scale + rel_h[kh] + rel_w[kw] (registers) -> exp2 -> fp16 P -> smem (SW128)
//   O_j = P V_j  : A = P (K-major), B = V tile as loaded ([key][hd] = MN-major SW128), N = 64 (+16 for hd 80)
//   O accumulates in registers with the usual running-max rescale; normalised and written as fp16.
#pragma once
#include "common.cuh"

namespace samrs {

struct AttnParams {
  const float* rel;     // [heads][4096][NP] fp32 (NP = 256 global / 64 windowed), log2(e) * q.[rel_pos_h ; rel_pos_w]:
                        //   rel_h[kh] = rel[qh - kh + S-1],  rel_w[kw] = rel[(2S-1) + qw - kw + S-1]
  const __half* rel16;  // attention v2: the same table already divided by scale_log2e and rounded to fp16
                        //   (written by the rel-pos GEMM's epilogue; it is exactly the value the kernel used to compute)
  __half* out;          // [4096][D] fp16, head-major columns (h*HD + c)
  int D;                // embed dim
  int heads;
  int num_qtiles;       // windowed: 25 windows * 2 halves; global: 32
  float scale_log2e;    // hd^-0.5 * log2(e)
  unsigned long long* dbg;   // optional pipeline trace of CTA 0 (tools/attn_trace.py)
  int pv_split;         // attention v2, head dim 80: 1 = issue P.V as an N=64 and an N=16 MMA per k-step (first version), 0 = one N=80 MMA
};
__device__ __forceinline__ void attn_dbg(const AttnParams& p, int slot) {
  if (p.dbg != nullptr && blockIdx.x == 0 && slot < 4096) p.dbg[slot] = clock64();
}

template <int HD, int BX, int QBY, int KBY, int NKT>
struct AttnCfg {
  static constexpr int NATOM = (HD + 63) / 64;
  static constexpr int KR = BX * KBY;                       // keys per tile (196 / 128)
  static constexpr int SN = (KR + 15) / 16 * 16;            // MMA N for S (208 / 128)
  static constexpr int QR = BX * QBY;                       // query rows per tile (98 / 128)
  static constexpr int PATOMS = (SN + 63) / 64;             // P tile = PATOMS x [128 x 64] fp16
  static constexpr int KV_STAGES = (NKT > 1) ? 2 : 1;
  static constexpr int Q_BYTES = NATOM * 128 * 128;
  static constexpr int KV_ATOM_BYTES = SN * 128;            // SN rows x 128 B, multiple of 1024 (SN % 8 == 0)
  static constexpr int KV_BYTES = NATOM * KV_ATOM_BYTES;
  static constexpr int P_BYTES = PATOMS * 128 * 128;
  static constexpr int SBUF = (NKT > 1) ? 128 : 256;        // TMEM column stride between S buffers
  static constexpr int O_COL = 256;                         // TMEM column of the O tile
  static constexpr int kSmemBytes = Q_BYTES + 2 * KV_STAGES * KV_BYTES + P_BYTES + 1024 + 256;
  static_assert(SN % 8 == 0 && SN <= 256, "bad S tile");
  static_assert(NKT == 1 || SN == 128, "streaming path assumes 128-key tiles");
};

template <int HD, int BX, int QBY, int KBY, int NKT>
__global__ void __launch_bounds__(256, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  using C = AttnCfg<HD, BX, QBY, KBY, NKT>;
  constexpr int NATOM = C::NATOM, SN = C::SN, KR = C::KR, QR = C::QR, ST = C::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::Q_BYTES;
  uint8_t* sV = sK + ST * C::KV_BYTES;
  uint8_t* sP = sV + ST * C::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;        // [ST]
  uint64_t* k_empty = bars + 4;       // [ST]
  uint64_t* v_full = bars + 6;        // [ST]
  uint64_t* v_empty = bars + 8;       // [ST]
  uint64_t* s_full = bars + 10;       // [2]
  uint64_t* p_full = bars + 12;
  uint64_t* o_full = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = p.num_qtiles * p.heads;

  // zero the operand tiles once: rows that TMA never writes (beyond the box) must stay finite
  for (int i = threadIdx.x; i < (C::Q_BYTES + 2 * ST * C::KV_BYTES + C::P_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
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
    }
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> coordinates
  auto item_coords = [&](int item, int& head, int& qy0, int& x0, int& ky0) {
    head = item % p.heads;
    const int qt = item / p.heads;
    if (NKT == 1) {              // windowed: qt = window * 2 + half
      const int win = qt >> 1, half = qt & 1;
      const int wy = win / 5, wx = win % 5;
      x0 = wx * BX;
      ky0 = wy * KBY;
      qy0 = ky0 + half * QBY;
    } else {                     // global: qt = pair of grid rows
      x0 = 0;
      ky0 = 0;
      qy0 = qt * QBY;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t qph = 0, kst = 0, kph = 0, vst = 0, vph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int head, qy0, x0, ky0;
        item_coords(item, head, qy0, x0, ky0);
        mbar_wait(q_empty, qph ^ 1);
        mbar_expect_tx(q_full, NATOM * QR * 128);
        for (int a = 0; a < NATOM; ++a) tma_load_3d(sQ + a * 128 * 128, &tmQ, q_full, head * HD + a * 64, x0, qy0);
        qph ^= 1;
        for (int j = 0; j < NKT; ++j) {
          mbar_wait(&k_empty[kst], kph ^ 1);
          mbar_expect_tx(&k_full[kst], NATOM * KR * 128);
          for (int a = 0; a < NATOM; ++a)
            tma_load_3d(sK + kst * C::KV_BYTES + a * C::KV_ATOM_BYTES, &tmKV, &k_full[kst],
                        p.D + head * HD + a * 64, x0, ky0 + j * KBY);
          if (++kst == ST) { kst = 0; kph ^= 1; }
          mbar_wait(&v_empty[vst], vph ^ 1);
          mbar_expect_tx(&v_full[vst], NATOM * KR * 128);
          for (int a = 0; a < NATOM; ++a)
            tma_load_3d(sV + vst * C::KV_BYTES + a * C::KV_ATOM_BYTES, &tmKV, &v_full[vst],
                        2 * p.D + head * HD + a * 64, x0, ky0 + j * KBY);
          if (++vst == ST) { vst = 0; vph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, SN, 0, 0);
      constexpr uint32_t idesc_o64 = umma_idesc_f16(128, 64, 0, 1);
      constexpr uint32_t idesc_o16 = umma_idesc_f16(128, 16, 0, 1);
      uint32_t qph = 0, kst = 0, kph = 0, vst = 0, vph = 0, pph = 0;
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      auto issue_s = [&](int buf) {
        mbar_wait(&k_full[kst], kph);
        tc_fence_after();
        const uint32_t d = tmem_base + uint32_t(buf * C::SBUF);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t a = aQ + (k / 4) * (128 * 128) + (k % 4) * 32;
          const uint32_t b = aK + kst * C::KV_BYTES + (k / 4) * C::KV_ATOM_BYTES + (k % 4) * 32;
          tc_mma_f16(d, umma_desc_sw128(a), umma_desc_sw128(b), idesc_s, k != 0);
        }
        tc_commit(&k_empty[kst]);
        tc_commit(&s_full[buf]);
        if (++kst == ST) { kst = 0; kph ^= 1; }
      };
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        mbar_wait(q_full, qph);
        tc_fence_after();
        issue_s(0);
        for (int j = 0; j < NKT; ++j) {
          if (j + 1 < NKT) issue_s((j + 1) & 1);
          if (j + 1 == NKT) tc_commit(q_empty);            // every S MMA of this item has been issued
          mbar_wait(p_full, pph);
          pph ^= 1;
          mbar_wait(&v_full[vst], vph);
          tc_fence_after();
          const uint32_t d = tmem_base + C::O_COL;
#pragma unroll
          for (int k = 0; k < SN / 16; ++k) {
            const uint32_t a = aP + (k / 4) * (128 * 128) + (k % 4) * 32;
            const uint32_t b = aV + vst * C::KV_BYTES + k * 2048;           // 16 key rows x 128 B
            tc_mma_f16(d, umma_desc_sw128(a), umma_desc_sw128(b), idesc_o64, k != 0);
            if (NATOM == 2)
              tc_mma_f16(d + 64, umma_desc_sw128(a), umma_desc_sw128(b + C::KV_ATOM_BYTES), idesc_o16, k != 0);
          }
          tc_commit(&v_empty[vst]);
          tc_commit(o_full);
          if (++vst == ST) { vst = 0; vph ^= 1; }
        }
        qph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax + output
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    uint32_t sph[2] = {0, 0};
    uint32_t oph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int head, qy0, x0, ky0;
      item_coords(item, head, qy0, x0, ky0);
      const int ty = qy0 + r / BX, tx = x0 + r % BX;
      const bool valid = (r < QR) && ty < 64 && tx < 64;
      const int token = ty * 64 + tx;
      constexpr int NP = (BX == 64) ? 256 : 64;
      constexpr int SS = BX;                                  // 64 (global) or 14 (window): rel-pos table half-size
      const float* relrow = p.rel + (size_t(head) * 4096 + (valid ? token : 0)) * NP;
      const int qh = qy0 - ky0 + r / BX, qw = r % BX;
      float relw[BX];
#pragma unroll
      for (int i = 0; i < BX; ++i) relw[i] = valid ? __ldg(relrow + (2 * SS - 1) + qw + (SS - 1) - i) : 0.f;
      float o_acc[HD];
#pragma unroll
      for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;
      float m_run = -INFINITY, l_run = 0.f;

      for (int j = 0; j < NKT; ++j) {
        const int buf = j & 1;
        float relh[KBY];
#pragma unroll
        for (int i = 0; i < KBY; ++i) relh[i] = valid ? __ldg(relrow + qh + (SS - 1) - ((NKT == 1 ? 0 : j * KBY) + i)) : 0.f;
        mbar_wait(&s_full[buf], sph[buf]);
        sph[buf] ^= 1;
        tc_fence_after();
        const uint32_t s_addr = lane_addr + uint32_t(buf * C::SBUF);
        // pass 1: row max of the biased, scaled scores (log2 domain)
        float m_tile = -INFINITY;
#pragma unroll
        for (int c0 = 0; c0 < SN; c0 += 32) {
          if (SN - c0 >= 32) {
            uint32_t v[32];
            tmem_ld32(s_addr + c0, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = c0 + i;
              if (c < KR) m_tile = fmaxf(m_tile, fmaf(__uint_as_float(v[i]), p.scale_log2e, relh[c / BX] + relw[c % BX]));
            }
          } else {
            uint32_t v[16];
            tmem_ld16(s_addr + c0, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = c0 + i;
              if (c < KR) m_tile = fmaxf(m_tile, fmaf(__uint_as_float(v[i]), p.scale_log2e, relh[c / BX] + relw[c % BX]));
            }
          }
        }
        const float m_new = fmaxf(m_run, m_tile);
        const float alpha = exp2f(m_run - m_new);       // 0 on the first tile (m_run = -inf)
        float l_tile = 0.f;
        // pass 2: P = exp2(s - m_new) -> fp16 -> swizzled smem (A operand of the PV MMA)
#pragma unroll
        for (int c0 = 0; c0 < SN; c0 += 32) {
          if (SN - c0 >= 32) {
            uint32_t v[32];
            tmem_ld32(s_addr + c0, v);
            tc_wait_ld();
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float e0 = 0.f, e1 = 0.f;
              const int c = c0 + i;
              if (c < KR) e0 = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2e, relh[c / BX] + relw[c % BX]) - m_new);
              if (c + 1 < KR) e1 = exp2f(fmaf(__uint_as_float(v[i + 1]), p.scale_log2e, relh[(c + 1) / BX] + relw[(c + 1) % BX]) - m_new);
              if (!valid) { e0 = 0.f; e1 = 0.f; }
              l_tile += e0 + e1;
              __half2 h = __floats2half2_rn(e0, e1);
              pk[i / 2] = *reinterpret_cast<uint32_t*>(&h);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = c0 + g * 8;
              uint8_t* dst = sP + (col / 64) * (128 * 128) + sw128_offset(r, (col % 64) / 8);
              *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
            }
          } else {
            uint32_t v[16];
            tmem_ld16(s_addr + c0, v);
            tc_wait_ld();
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float e0 = 0.f, e1 = 0.f;
              const int c = c0 + i;
              if (c < KR) e0 = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2e, relh[c / BX] + relw[c % BX]) - m_new);
              if (c + 1 < KR) e1 = exp2f(fmaf(__uint_as_float(v[i + 1]), p.scale_log2e, relh[(c + 1) / BX] + relw[(c + 1) % BX]) - m_new);
              if (!valid) { e0 = 0.f; e1 = 0.f; }
              l_tile += e0 + e1;
              __half2 h = __floats2half2_rn(e0, e1);
              pk[i / 2] = *reinterpret_cast<uint32_t*>(&h);
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int col = c0 + g * 8;
              uint8_t* dst = sP + (col / 64) * (128 * 128) + sw128_offset(r, (col % 64) / 8);
              *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
            }
          }
        }
        l_run = l_run * alpha + l_tile;
        m_run = m_new;
        // publish P to the tensor core (generic-proxy writes -> async proxy), S buffer is free again
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_full);
        // O_j = P V_j
        mbar_wait(o_full, oph);
        oph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 32) {
          if (HD - c0 >= 32) {
            uint32_t v[32];
            tmem_ld32(lane_addr + C::O_COL + c0, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o_acc[c0 + i] = fmaf(o_acc[c0 + i], alpha, __uint_as_float(v[i]));
          } else {
            uint32_t v[16];
            tmem_ld16(lane_addr + C::O_COL + c0, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o_acc[c0 + i] = fmaf(o_acc[c0 + i], alpha, __uint_as_float(v[i]));
          }
        }
        tc_fence_before();
      }
      if (valid) {
        const float inv = 1.0f / l_run;
        __half* o = p.out + size_t(token) * p.D + head * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 8) {
          __half2 h0 = __floats2half2_rn(o_acc[c] * inv, o_acc[c + 1] * inv);
          __half2 h1 = __floats2half2_rn(o_acc[c + 2] * inv, o_acc[c + 3] * inv);
          __half2 h2 = __floats2half2_rn(o_acc[c + 4] * inv, o_acc[c + 5] * inv);
          __half2 h3 = __floats2half2_rn(o_acc[c + 6] * inv, o_acc[c + 7] * inv);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2);
          pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(o + c) = pk;
        }
      }
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
