// Rotated-box -> mask-prompt rasterisation on the device.
//
// Replaces the per-box host loop of `Generate Dataset/main_sam_rbox_mask_instance.py:125-141`:
//     cv2.fillPoly(canvas, [poly.astype(int32)], 255) -> +-1000 float64 -> cv2.resize(long side 1024, INTER_LINEAR)
//     -> cv2.copyMakeBorder(bottom / right, -1000) -> cv2.resize((256, 256), INTER_LINEAR) -> float32
// The arithmetic lives in the un-vendored dependency OpenCV (4.13 installed; the reference pins none), so its rules were
// pinned against cv2 itself before this was written (oracle/rbox_prompt_oracle.py, tests/test_rbox_prompt.py):
//   * fillPoly (line_type 8, shift 0) = the four edges drawn with the 8-connected Bresenham line of cv::LineIterator
//     (left-to-right, err = dx - 2 dy, minor step when err < 0) UNION the scan-line fill of the edge table: per edge
//     x(y) = (x_top << 16) + (y - y_top) * ((dx << 16) / dy) (C integer division), rows [y_top, y_bottom), active edges sorted
//     by x, consecutive pairs filled from ceil(x_left / 2^16) to floor(x_right / 2^16).  Bit-exact on 7 000 random quads
//     (convex, clipped and self-intersecting) against cv2.fillPoly.
//   * resize of CV_64F with INTER_LINEAR: src = (dst + 0.5) * (1 / (dsize / ssize)) - 0.5 in double, i0 = floor, weight
//     w1 = src - i0 (double), i0 < 0 -> (0, w1 = 0), i0 >= n - 1 -> (n - 1, w1 = 0); horizontal pass first, then vertical,
//     each `a * w0 + b * w1` in double.  Identical to cv2 except for isolated 1-ulp (float32) differences (15 of 7.8 M
//     elements in the pin run): OpenCV's vectorised pass rounds a handful of sums differently.
// Vertices must lie inside the image (the drivers' annotations do; cv2's line clipping of outside vertices is not restated).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace samrs {

struct RboxEdge { int y0, y1; long long x, dx; };

// One block per polygon: threads 0..3 draw the four boundary lines, then one thread per image row fills the spans.
__global__ void __launch_bounds__(256) rbox_fill_kernel(const float* __restrict__ polys /*[B][4][2]*/, int H, int W,
                                                        uint8_t* __restrict__ mask /*[B][H][W], zeroed*/, int* __restrict__ bad) {
  const int b = blockIdx.x;
  uint8_t* m = mask + size_t(b) * H * W;
  __shared__ int px[4], py[4];
  __shared__ RboxEdge edges[4];
  __shared__ int n_edges;
  if (threadIdx.x < 4) {
    px[threadIdx.x] = int(polys[(size_t(b) * 4 + threadIdx.x) * 2]);          // astype(np.int32): truncation toward zero
    py[threadIdx.x] = int(polys[(size_t(b) * 4 + threadIdx.x) * 2 + 1]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    bool inside = true;
    for (int i = 0; i < 4; ++i) inside = inside && px[i] >= 0 && px[i] < W && py[i] >= 0 && py[i] < H;
    if (!inside) atomicExch(bad, 1);
    for (int i = 0; i < 4 && inside; ++i) {
      const int j = (i + 3) & 3;                                               // edge from vertex j (previous) to vertex i
      if (py[j] == py[i]) continue;                                            // horizontal edges only draw their line
      const long long X0 = (long long)px[j] << 16, X1 = (long long)px[i] << 16;
      RboxEdge e;
      e.dx = (X1 - X0) / (long long)(py[i] - py[j]);                            // C division: truncates toward zero
      if (py[j] < py[i]) { e.y0 = py[j]; e.y1 = py[i]; e.x = X0; }
      else { e.y0 = py[i]; e.y1 = py[j]; e.x = X1; }
      edges[n++] = e;
    }
    n_edges = inside ? n : 0;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    // cv::Line(img, p_prev, p_cur, color, 8): LineIterator(connectivity 8, leftToRight = true)
    const int i = threadIdx.x, j = (i + 3) & 3;
    int x0 = px[j], y0 = py[j], x1 = px[i], y1 = py[i];
    const bool inside = x0 >= 0 && x0 < W && y0 >= 0 && y0 < H && x1 >= 0 && x1 < W && y1 >= 0 && y1 < H;
    if (inside) {
      int dx = x1 - x0, dy = y1 - y0, sy = 1;
      if (dx < 0) { dx = -dx; dy = -dy; x0 = x1; y0 = y1; }
      if (dy < 0) { dy = -dy; sy = -1; }
      const bool vert = dy > dx;
      if (vert) { const int t = dx; dx = dy; dy = t; }
      int err = dx - (dy + dy);
      const int plus = dx + dx, minus = -(dy + dy);
      int x = x0, y = y0;
      for (int k = 0; k <= dx; ++k) {
        m[size_t(y) * W + x] = 1;
        const bool step = err < 0;
        err += minus + (step ? plus : 0);
        if (vert) { y += sy; if (step) x += 1; }
        else { x += 1; if (step) y += sy; }
      }
    }
  }
  __syncthreads();
  const int ne = n_edges;
  for (int y = threadIdx.x; y < H; y += blockDim.x) {
    long long xs[4];
    int na = 0;
    for (int i = 0; i < ne; ++i)
      if (y >= edges[i].y0 && y < edges[i].y1) xs[na++] = edges[i].x + (long long)(y - edges[i].y0) * edges[i].dx;
    for (int i = 1; i < na; ++i)                                                // insertion sort by x (at most 4 entries)
      for (int k = i; k > 0 && xs[k - 1] > xs[k]; --k) { const long long t = xs[k]; xs[k] = xs[k - 1]; xs[k - 1] = t; }
    for (int i = 0; i + 1 < na; i += 2) {
      long long a = (xs[i] + 65535) >> 16, c = xs[i + 1] >> 16;                 // ceil(left) .. floor(right)
      if (a < W && c >= 0 && a <= c) {
        if (a < 0) a = 0;
        if (c >= W) c = W - 1;
        for (long long x = a; x <= c; ++x) m[size_t(y) * W + x] = 1;
      }
    }
  }
}

// cv2.resize(INTER_LINEAR) source index and weight of destination index d (double, as OpenCV computes them for CV_64F)
__device__ __forceinline__ void cv_lin_coeff(int d, int ssize, int dsize, int& i0, int& i1, double& w0, double& w1) {
  const double scale = 1.0 / (double(dsize) / double(ssize));
  double f = __dadd_rn(__dmul_rn(double(d) + 0.5, scale), -0.5);
  int s = int(floor(f));
  f = f - double(s);
  if (s < 0) { s = 0; f = 0.0; }
  if (s >= ssize - 1) { s = ssize - 1; f = 0.0; }
  i0 = s;
  i1 = min(s + 1, ssize - 1);
  w0 = 1.0 - f;
  w1 = f;
}
__device__ __forceinline__ double cv_lerp(double a, double w0, double b, double w1) {
  return __dadd_rn(__dmul_rn(a, w0), __dmul_rn(b, w1));                         // no contraction: two products, one sum
}

// value of the long-side-1024 image padded to 1024 x 1024 at (Y, X): stage 1 of the recipe evaluated on demand
__device__ __forceinline__ double rbox_stage1(const uint8_t* __restrict__ m, int H, int W, int nh, int nw, int Y, int X) {
  if (Y >= nh || X >= nw) return -1000.0;                                       // copyMakeBorder value
  auto px = [&](int y, int x) { return m[size_t(y) * W + x] ? 1000.0 : -1000.0; };
  if (nh == H && nw == W) return px(Y, X);                                      // cv2.resize to the same size copies
  int x0, x1, y0, y1;
  double a0, a1, b0, b1;
  cv_lin_coeff(X, W, nw, x0, x1, a0, a1);
  cv_lin_coeff(Y, H, nh, y0, y1, b0, b1);
  const double r0 = cv_lerp(px(y0, x0), a0, px(y0, x1), a1);
  const double r1 = cv_lerp(px(y1, x0), a0, px(y1, x1), a1);
  return cv_lerp(r0, b0, r1, b1);
}

__global__ void __launch_bounds__(256) rbox_prompt_kernel(const uint8_t* __restrict__ mask /*[B][H][W]*/, int H, int W, int nh, int nw,
                                                          float* __restrict__ out /*[B][256][256]*/) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y, b = blockIdx.z;
  if (ox >= 256) return;
  const uint8_t* m = mask + size_t(b) * H * W;
  int x0, x1, y0, y1;
  double a0, a1, b0, b1;
  cv_lin_coeff(ox, 1024, 256, x0, x1, a0, a1);
  cv_lin_coeff(oy, 1024, 256, y0, y1, b0, b1);
  const double r0 = cv_lerp(rbox_stage1(m, H, W, nh, nw, y0, x0), a0, rbox_stage1(m, H, W, nh, nw, y0, x1), a1);
  const double r1 = cv_lerp(rbox_stage1(m, H, W, nh, nw, y1, x0), a0, rbox_stage1(m, H, W, nh, nw, y1, x1), a1);
  out[(size_t(b) * 256 + oy) * 256 + ox] = float(cv_lerp(r0, b0, r1, b1));
}

}  // namespace samrs
