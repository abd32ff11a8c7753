// Instance payload on the device: uncompressed COCO run-length encoding + area of every mask of a tile.
//
// Replaces, per mask, the driver's D2H of a 1 MiB bool mask followed by `maskUtils.encode(np.asfortranarray(mask))`
// and `np.sum(mask)` (Generate Dataset/main_sam_hbox_semantic.py:200-203).  The run semantics are the ones the
// reference pins in-repo (segment_anything/utils/amg.py:107-135, mask_to_rle_pytorch): pixels in column-major
// (Fortran) order, alternating run lengths starting with a run of zeros - a mask whose first pixel is set gets a
// leading 0 - and the lengths sum to H*W.
//
// Three passes of integer / bit work, all HBM- or latency-bound (no tensor cores):
//   1. pack: every thread owns one column x and 32 rows: packed[b][wy][x] bit r = mask[b][32 wy + r][x].  Loads are
//      coalesced along x (row-major source), the stores of consecutive threads are consecutive words.  The fused
//      variant computes the bits straight from the 256x256 low-res logits with the same explicitly rounded
//      bilinear arithmetic as upsample4_threshold_kernel, so the 1 MiB/mask bool tensor is never materialised.
//   2. count: one CTA per mask walks the packed words in column-major order (thread t owns a contiguous range of
//      columns); transitions of a word are  w ^ ((w << 1) | carry)  with the carry taken from the previous pixel in
//      Fortran order (the last row of the previous column; 0 before the first pixel); popcounts give the number of
//      runs and the area.  A one-thread scan over the B masks turns run counts into offsets.
//   3. emit: the same walk again; a block-wide exclusive scan gives every thread the rank of its first transition and
//      the position of the last transition before it, so run k is written as position_k - position_{k-1}.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace samrs {

// bits of column x, rows 32*wy .. 32*wy+31 of a row-major u8 mask (non-zero = set); rows >= H read as 0
__global__ void rle_pack_kernel(const uint8_t* __restrict__ masks, int H, int W, int HW32, uint32_t* __restrict__ packed) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int wy = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const uint8_t* m = masks + (size_t(b) * H + size_t(wy) * 32) * W + x;
  const int nr = min(32, H - wy * 32);
  uint32_t w = 0;
#pragma unroll 8
  for (int r = 0; r < nr; ++r) w |= (m[size_t(r) * W] != 0 ? 1u : 0u) << r;
  packed[(size_t(b) * HW32 + wy) * W + x] = w;
}

// same bits from the low-res logits of a 1024x1024 tile: bit = (bilinear x4 upsample > 0), arithmetic of simt.cuh bilerp.
// The 32 output rows of a word touch low-res rows 8 wy - 1 .. 8 wy + 8 only, so the horizontal interpolation
// r(j) = lx0 a[j][x0] + lx1 a[j][x1] is evaluated once per low-res row (10 instead of 64 times) and reused; the values and
// the order of the roundings are those of bilerp(), so the bits equal upsample4_threshold_kernel's.  Rows clamped at the
// borders carry weight 0 (top) or duplicate row 255 (bottom), exactly as ATen's index clamp does.
__global__ void rle_pack_lowres_kernel(const float* __restrict__ low /*[B][256][256]*/, uint32_t* __restrict__ packed /*[B][32][1024]*/) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int wy = blockIdx.y, b = blockIdx.z;
  if (x >= 1024) return;
  const float* a = low + size_t(b) * 65536;
  int x0, x1;
  float lx0, lx1;
  src_index_x4(x, 256, x0, x1, lx0, lx1);
  float hrow[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const int j = min(max(8 * wy - 1 + k, 0), 255);
    hrow[k] = __fadd_rn(__fmul_rn(lx0, a[j * 256 + x0]), __fmul_rn(lx1, a[j * 256 + x1]));
  }
  uint32_t w = 0;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    int y0, y1;
    float ly0, ly1;
    src_index_x4(wy * 32 + r, 256, y0, y1, ly0, ly1);                 // only the weights are used; the rows are hrow[i0], hrow[i0 + 1]
    const int i0 = (r >> 2) + ((r & 3) >= 2 ? 1 : 0);
    const float v = __fadd_rn(__fmul_rn(ly0, hrow[i0]), __fmul_rn(ly1, hrow[i0 + 1]));
    w |= (v > 0.0f ? 1u : 0u) << r;
  }
  packed[(size_t(b) * 32 + wy) * 1024 + x] = w;
}

constexpr int RLE_THREADS = 1024;

// transitions of one packed word; `carry` = value of the previous pixel in Fortran order
__device__ __forceinline__ uint32_t rle_transitions(uint32_t w, uint32_t valid, uint32_t carry) {
  return (w ^ ((w << 1) | carry)) & valid;
}

// pass 2 (emit == false): runs[b] = transitions + 1, area[b];  pass 3 (emit == true): counts written at offsets[b]
template <bool EMIT>
__global__ void __launch_bounds__(RLE_THREADS) rle_scan_kernel(const uint32_t* __restrict__ packed, int H, int W, int HW32,
                                                               long long* __restrict__ runs, long long* __restrict__ area,
                                                               const long long* __restrict__ offsets, uint32_t* __restrict__ counts,
                                                               long long capacity, long long* __restrict__ tstate /*[B][1024][2]*/) {
  const int b = blockIdx.x, t = threadIdx.x;
  const uint32_t* pk = packed + size_t(b) * HW32 * W;
  const int cpt = (W + RLE_THREADS - 1) / RLE_THREADS;            // columns per thread, contiguous in Fortran order
  const int xa = min(W, t * cpt), xb = min(W, xa + cpt);
  const int last_bits = H - (HW32 - 1) * 32;                       // valid rows of a column's last word (1..32)
  const uint32_t last_valid = last_bits == 32 ? 0xFFFFFFFFu : ((1u << last_bits) - 1u);

  // column-major walk over this thread's words
  auto walk = [&](auto&& on_word) {
    if (xa >= xb) return;
    uint32_t carry = 0;
    if (xa > 0) carry = (pk[size_t(HW32 - 1) * W + (xa - 1)] >> (last_bits - 1)) & 1u;
    for (int x = xa; x < xb; ++x) {
#pragma unroll 8
      for (int wy = 0; wy < HW32; ++wy) {
        const bool lastw = (wy == HW32 - 1);
        const uint32_t valid = lastw ? last_valid : 0xFFFFFFFFu;
        const uint32_t w = pk[size_t(wy) * W + x] & valid;
        on_word(x, wy, w, rle_transitions(w, valid, carry));
        carry = (w >> ((lastw ? last_bits : 32) - 1)) & 1u;
      }
    }
  };

  // per-thread totals: computed by the count pass and kept in `tstate` for the emit pass
  unsigned long long n_tr = 0, n_set = 0;
  long long last_pos = -1;
  long long* ts = tstate + (size_t(b) * RLE_THREADS + t) * 2;
  if (!EMIT) {
    walk([&](int x, int wy, uint32_t w, uint32_t tr) {
      n_tr += __popc(tr);
      n_set += __popc(w);
      if (tr) last_pos = (long long)x * H + wy * 32 + (31 - __clz(tr));
    });
    ts[0] = (long long)n_tr;
    ts[1] = last_pos;
  } else {
    n_tr = (unsigned long long)ts[0];
    last_pos = ts[1];
  }

  // block-wide exclusive scans: sum of transitions, max of last positions (and the total area)
  __shared__ unsigned long long s_sum[RLE_THREADS / 32], s_area[RLE_THREADS / 32];
  __shared__ long long s_max[RLE_THREADS / 32];
  const int lane = t & 31, warp = t >> 5;
  unsigned long long inc = n_tr, ar = n_set;
  long long mx = last_pos;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
    const long long m = __shfl_up_sync(0xffffffffu, mx, o);
    if (lane >= o) { inc += u; mx = max(mx, m); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ar += __shfl_xor_sync(0xffffffffu, ar, o);
  if (lane == 31) { s_sum[warp] = inc; s_max[warp] = mx; }
  if (lane == 0) s_area[warp] = ar;
  __syncthreads();
  unsigned long long base = 0, total = 0, area_total = 0;
  long long prev_max = -1, max_total = -1;
  for (int i = 0; i < RLE_THREADS / 32; ++i) {
    if (i < warp) { base += s_sum[i]; prev_max = max(prev_max, s_max[i]); }
    total += s_sum[i];
    area_total += s_area[i];
    max_total = max(max_total, s_max[i]);
  }
  const unsigned long long rank0 = base + inc - n_tr;             // transitions before this thread
  long long prev = __shfl_up_sync(0xffffffffu, mx, 1);            // last transition position before this thread
  prev = (lane == 0) ? prev_max : max(prev, prev_max);

  if (!EMIT) {
    if (t == 0) { runs[b] = (long long)total + 1; area[b] = (long long)area_total; }
    return;
  }
  const long long off = offsets[b];
  unsigned long long rank = rank0;
  long long pp = prev < 0 ? 0 : prev;                             // run 0 is measured from pixel 0
  walk([&](int x, int wy, uint32_t w, uint32_t tr) {
    while (tr) {
      const int bit = __ffs(tr) - 1;
      tr &= tr - 1;
      const long long p = (long long)x * H + wy * 32 + bit;
      if (off + (long long)rank < capacity) counts[off + rank] = uint32_t(p - pp);
      pp = p;
      ++rank;
    }
  });
  if (t == 0) {                                                    // final run: from the last transition to H*W
    const long long lastp = max_total < 0 ? 0 : max_total;
    if (off + (long long)total < capacity) counts[off + total] = uint32_t((long long)H * W - lastp);
  }
}

// ------------------------------------------------------------------------------------------------
// Compressed COCO string of the runs (what `maskUtils.encode` returns in rle["counts"], main_sam_hbox_semantic.py:200-201):
// every run - from the fourth on, its difference to the run two before - is written as 5-bit groups, least significant
// first, bit 5 = "more groups follow", bit 4 of the last group = sign, each group + 48 as one ASCII character
// (pycocotools rleToString; parity with that un-vendored library is unpinned, the host restatement is samrs_b200/rle.py).
// One CTA per mask, 1024 runs per step: a block-wide exclusive scan of the group counts gives every run its position.
// EMIT == false: nchars[b] only.  EMIT == true: the characters at char_offsets[b].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int coco_groups(long long x) {
  int n = 0;
  bool more = true;
  while (more) {
    const int c = int(x & 0x1f);
    x >>= 5;
    more = (c & 0x10) ? (x != -1) : (x != 0);
    ++n;
  }
  return n;
}

template <bool EMIT>
__global__ void __launch_bounds__(RLE_THREADS) coco_string_kernel(const uint32_t* __restrict__ counts, const long long* __restrict__ offsets,
                                                                  long long run_capacity, long long* __restrict__ nchars,
                                                                  const long long* __restrict__ char_offsets, uint8_t* __restrict__ chars,
                                                                  long long char_capacity) {
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long base = offsets[b];
  long long n = offsets[b + 1] - base;
  if (base + n > run_capacity) n = 0;                              // the runs themselves did not fit: nothing to encode
  const uint32_t* c = counts + base;
  __shared__ int s_warp[RLE_THREADS / 32];
  long long running = 0;
  const long long out0 = EMIT ? char_offsets[b] : 0;
  for (long long i0 = 0; i0 < n; i0 += RLE_THREADS) {
    const long long i = i0 + t;
    long long x = 0;
    int k = 0;
    if (i < n) {
      x = (long long)c[i];
      if (i > 2) x -= (long long)c[i - 2];
      k = coco_groups(x);
    }
    int inc = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < RLE_THREADS / 32; ++w) {
      if (w < warp) wbase += s_warp[w];
      total += s_warp[w];
    }
    if (EMIT && i < n) {
      long long p = out0 + running + wbase + inc - k;
      bool more = true;
      while (more) {
        int ch = int(x & 0x1f);
        x >>= 5;
        more = (ch & 0x10) ? (x != -1) : (x != 0);
        if (more) ch |= 0x20;
        if (p < char_capacity) chars[p] = uint8_t(ch + 48);
        ++p;
      }
    }
    running += total;
    __syncthreads();
  }
  if (!EMIT && t == 0) nchars[b] = running;
}

__global__ void rle_offsets_kernel(const long long* __restrict__ runs, int B, long long* __restrict__ offsets) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long acc = 0;
    for (int b = 0; b < B; ++b) { offsets[b] = acc; acc += runs[b]; }
    offsets[B] = acc;
  }
}

}  // namespace samrs
