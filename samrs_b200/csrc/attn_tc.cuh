// Parameters and tile geometry shared by the tcgen05 attention kernel (attn_tc2.cuh) of the SAM image encoder.
//
// `Attention.forward` (SA/modeling/image_encoder.py:224-240) + `add_decomposed_rel_pos` (:325-361) for both block kinds:
//   * 14x14 windowed blocks: one query tile = half a window (7 rows x 14 tokens = 98 queries), one key tile = the whole
//     window (196 keys, zero-padded tokens included);
//   * global blocks: query tile = 2 grid rows (128 queries), 32 key tiles of 2 grid rows (128 keys), streaming (online)
//     softmax so the 4096x4096 score matrix is never materialised.
// Tokens are addressed through one 3-D TMA tensor map over the qkv activation viewed as [y=64][x=64][3*D] fp16;
// out-of-grid (padded) tokens are zero-filled by TMA.  Because the K bias only shifts every score of a row by the same
// amount and the V bias adds bv to every output row (softmax rows sum to 1), the qkv GEMM omits both and the V bias is
// folded into the proj bias; zero-filled padded keys then reproduce the reference's "padded k,v = qkv bias" semantics.
#pragma once
#include "common.cuh"

namespace samrs {

struct AttnParams {
  const float* rel;     // [heads][4096][NP] fp32 (NP = 256 global / 64 windowed), log2(e) * q.[rel_pos_h ; rel_pos_w]:
                        //   rel_h[kh] = rel[qh - kh + S-1],  rel_w[kw] = rel[(2S-1) + qw - kw + S-1]
  const __half* rel16;  // global blocks: the same table already divided by scale_log2e and rounded to fp16 (written by the
                        //   rel-pos GEMM's epilogue); windowed blocks compute their terms in the kernel and ignore it
  __half* out;          // [4096][D] fp16, head-major columns (h*HD + c)
  int D;                // embed dim
  int heads;
  int num_qtiles;       // query-tile PAIRS: windowed 25 (windows, two 7-row halves each); global 16
  float scale_log2e;    // hd^-0.5 * log2(e)
  float rel_scale;      // windowed blocks: 1 / scale_log2e, applied to G = q . log2e [rel_pos_h ; rel_pos_w]^T before it becomes R
  unsigned long long* dbg;   // optional pipeline trace of CTA 0 (tools/attn_trace.py)
  int pv_split;         // head dim 80: 1 = issue P.V as an N=64 and an N=16 MMA per k-step (first version), 0 = one N=80 MMA
};
#ifdef SAMRS_EXPERIMENTS
__device__ __forceinline__ void attn_dbg(const AttnParams& p, int slot) {
  if (p.dbg != nullptr && blockIdx.x == 0 && slot < 4096) p.dbg[slot] = clock64();
}
#else
__device__ __forceinline__ void attn_dbg(const AttnParams&, int) {}
#endif

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

}  // namespace samrs
