// samrs_b200 engine: weight ingest, encoder / decoder orchestration and the C ABI (include/samrs_b200.h).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/samrs_b200.h"
#include "attn_tc.cuh"
#include "attn_tc2.cuh"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "simt.cuh"
#include "rle.cuh"
#include "rbox.cuh"

namespace samrs {

static thread_local std::string g_last_error;

int fail(const char* file, int line, const char* msg) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s:%d: %s", file, line, msg);
  g_last_error = buf;
  return 1;
}
#define SAMRS_FAIL(msg) return samrs::fail(__FILE__, __LINE__, msg)
#define SAMRS_TRY(expr)          \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                 uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) SAMRS_FAIL("cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) SAMRS_FAIL("cuTensorMapEncodeTiled(2d) failed");
  return 0;
}
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes,
                 uint64_t pitch2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) SAMRS_FAIL("cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) SAMRS_FAIL("cuTensorMapEncodeTiled(3d) failed");
  return 0;
}

int make_tmap_out(CUtensorMap* out, const void* base, bool half, uint64_t N, uint64_t M, uint64_t batch, uint64_t ld_elems,
                  uint64_t batch_stride_elems, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) SAMRS_FAIL("cuTensorMapEncodeTiled unavailable");
  const uint64_t es = half ? 2 : 4;
  cuuint64_t dims[3] = {N, M, batch};
  cuuint64_t strides[2] = {ld_elems * es, (batch > 1 ? batch_stride_elems : ld_elems * M) * es};
  cuuint32_t box[3] = {box_cols, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, (half || box_cols != 32) ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) SAMRS_FAIL("cuTensorMapEncodeTiled(out) failed");
  return 0;
}

// ------------------------------------------------------------------ per-engine launch state
// Everything a launch helper needs from "the engine this ABI call serves" lives in the engine's LaunchCtx; the ABI entry
// points publish it for the duration of the call through a thread_local pointer (LaunchScope), so two host threads can
// drive two engines concurrently and an engine on another device opts its kernels into large shared memory itself.
enum ProfCat { PC_GEMM = 0, PC_ATTN_WIN, PC_ATTN_GLOB, PC_RELPOS, PC_LN, PC_ENC_OTHER, PC_DECODER, PC_EPILOGUE, PC_COUNT };
struct ProfRec { int cat; cudaEvent_t a, b; };
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};
struct LaunchCtx {
  int64_t launches = 0;
  Profiler prof;
  float* splitk_ws = nullptr;                     // split-K partial sums of the token-side SGEMMs
  size_t splitk_ws_floats = 0;
  int* sk_flags = nullptr;                        // experiments build only: stream-K tile counters (zero between launches)
  std::unordered_set<const void*> smem_opted;     // kernels whose dynamic-smem limit was raised on this engine's device
  bool capturing = false;                         // a CUDA graph is being captured: no events, no attribute calls
  bool pdl = true;                                // programmatic dependent launch for the kernels that support it
};
static thread_local LaunchCtx* t_ctx = nullptr;
static inline void count_launch(int n = 1) {
  if (t_ctx) t_ctx->launches += n;
}
// launch with programmatic stream serialization (see common.cuh): only for kernels that call pdl_wait() themselves
static bool g_pdl_default = true;
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = (t_ctx ? t_ctx->pdl : g_pdl_default) ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename K>
static int opt_in_smem(K kernel, int bytes) {
  const void* key = reinterpret_cast<const void*>(kernel);
  if (t_ctx && t_ctx->smem_opted.count(key)) return 0;
  SAMRS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (t_ctx) t_ctx->smem_opted.insert(key);
  return 0;
}

template <int BN, bool OH, int ACT>
static int launch_gemm_inst(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const GemmParams& p, int grid, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  SAMRS_TRY(opt_in_smem(gemm_tc_kernel<BN, OH, ACT>, Cfg::kSmemBytes));
  SAMRS_CUDA_OK(launch_pdl(gemm_tc_kernel<BN, OH, ACT>, dim3(grid), dim3(GEMM_THREADS), Cfg::kSmemBytes, st, tA, tB, tC, p));
  count_launch();
  return 0;
}

#ifdef SAMRS_EXPERIMENTS
// cluster-multicast variant (CM = 2): launched with a (2,1,1) cluster
template <int BN, bool OH, int ACT>
static int launch_gemm_mc_inst(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const GemmParams& p, int grid, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  SAMRS_TRY(opt_in_smem(gemm_tc_kernel<BN, OH, ACT, 2>, Cfg::kSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SAMRS_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, OH, ACT, 2>, tA, tB, tC, p));
  count_launch();
  return 0;
}
#endif

template <int BN, bool OH, int ACT, int EW = GEMM_EPI_WARPS>
static int launch_gemm2_inst(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const CUtensorMap& tC16, const GemmParams& p, int grid, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN, EW>;
  SAMRS_TRY(opt_in_smem(gemm_tc2_kernel<BN, OH, ACT, EW>, Cfg::kSmemBytes));
  SAMRS_CUDA_OK(launch_pdl(gemm_tc2_kernel<BN, OH, ACT, EW>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, tA, tB, tC, tC16, p));
  count_launch();
  return 0;
}

#ifdef SAMRS_EXPERIMENTS
// clusters of four: how many fit on the device at once depends on how the SMs are spread over GPCs, so the grid is sized from
// the occupancy query (once per kernel and engine device)
template <int BN, bool OH, int ACT>
static int launch_gemm4_inst(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const GemmParams& p, int supers, int num_sms, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN>;
  SAMRS_TRY(opt_in_smem(gemm_tc4_kernel<BN, OH, ACT>, Cfg::kSmemBytes));
  static int max_clusters = -1;
  if (max_clusters < 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms & ~3); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tc4_kernel<BN, OH, ACT>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms / 4; }
    max_clusters = n < num_sms / 4 ? n : num_sms / 4;
    if (getenv("SAMRS_VERBOSE")) fprintf(stderr, "gemm_tc4<%d>: %d clusters of 4 co-resident\n", BN, max_clusters);
  }
  const int grid = 4 * (supers < max_clusters ? supers : max_clusters);
  SAMRS_CUDA_OK(launch_pdl(gemm_tc4_kernel<BN, OH, ACT>, dim3(grid), dim3(GEMM_THREADS), Cfg::kSmemBytes, st, tA, tB, tC, p));
  count_launch();
  return 0;
}

#endif

#ifndef SAMRS_GELU_EPI_WARPS
#define SAMRS_GELU_EPI_WARPS 8       // epilogue warps of the fp16 + GELU pair kernel; 12 (three column groups) measured the same (profiles/r02_gemm_epi12_ab.txt)
#endif
constexpr int GELU_EPI_WARPS = SAMRS_GELU_EPI_WARPS;
#ifdef SAMRS_EXPERIMENTS
constexpr int SK_MAX_TILES = 1024;   // counters the engine owns (two ints per tile)
#ifndef SAMRS_SK_BN
#define SAMRS_SK_BN 0                // N tile of the stream-K GEMMs on the ViT-H shapes (256 / 160); 0 = tile schedule, which measured faster
#endif
constexpr int SK_BN = SAMRS_SK_BN;
#endif

#ifdef SAMRS_EXPERIMENTS
template <int BN>
static int launch_gemm2_sk_inst(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC, const GemmParams& p, int grid, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN>;
  SAMRS_TRY(opt_in_smem(gemm_tc2_sk_kernel<BN>, Cfg::kSmemBytes));
  SAMRS_CUDA_OK(launch_pdl(gemm_tc2_sk_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), Cfg::kSmemBytes, st, tA, tB, tC, p));
  count_launch();
  return 0;
}
#endif

// model of one tile's duration in SM clocks: tensor-pipe time vs the L2->SM feed (~50 B/clk/SM measured)
// (feed rates fitted to tools/gemm_sweep.py on a B200: ~80 B/clk/SM for the 1-CTA kernel, ~52 for the CTA-pair kernel)
static double tile_clk(int bn, int K, bool pair) {
  const double mma = double(bn) * K / 32.0;
  const double l2 = (128.0 + (pair ? bn / 2.0 : double(bn))) * K * 2.0 / (pair ? 52.0 : 80.0);
  return (mma > l2 ? mma : l2) + 600.0;
}

int launch_gemm_tc(const __half* A, int lda, const __half* B, int ldb, const GemmParams& pin, bool out_half, int act,
                   int num_sms, cudaStream_t stream, int force_bn, const CUtensorMap* a_map_rank3) {
  GemmParams p = pin;
  if (p.batch < 1) p.batch = 1;
  p.a_rank3 = a_map_rank3 ? 1 : 0;
  if (p.K % 8 != 0 || lda % 8 != 0 || ldb % 8 != 0) SAMRS_FAIL("gemm: K and leading dimensions must be multiples of 8");
  // force_bn: 0 = choose; 128/160/224/256 = 1-CTA kernel with that N tile; 1000 + bn = CTA-pair kernel;
  // 2000 + bn = cluster of two 1-CTA tiles sharing B through TMA multicast
  // 3000 + bn = CTA-pair kernel with the stream-K schedule (fp32 accumulate epilogue only)
  int bn = force_bn % 1000;
  // 4000 + bn = two CTA pairs per cluster sharing the B tile through TMA multicast (gemm_tc4_kernel)
  bool pair = (force_bn >= 1000 && force_bn < 2000) || force_bn >= 3000;
  bool mcast = force_bn >= 2000 && force_bn < 3000;
  bool streamk = force_bn >= 3000 && force_bn < 4000;
  bool quad = force_bn >= 4000;
  if (force_bn == 0 && p.M == 4096 && p.batch == 1 && (p.N == 1280 || p.N == 3840 || p.N == 5120) && num_sms == 148) {
    // ViT-H shapes: measured best configurations (profiles/r01_gemm_sweep_v4.json): the CTA-pair kernel everywhere
    // (it halves the B bytes each SM pulls from L2; 1-CTA tiles are feed-bound once the issue loop is tight)
    bn = (p.N == 1280) ? 160 : 224;
    pair = true;
#ifdef SAMRS_EXPERIMENTS
    // -DSAMRS_SK_BN=160|256: proj / lin2 (out += ...) on the stream-K schedule, see gemm_tc2_sk_kernel
    if (SK_BN > 0 && p.N == 1280 && p.accumulate && !out_half && act == 0 && p.sk_flags != nullptr) { bn = SK_BN; streamk = true; }
    // experiment hook: SAMRS_BN="n1280,n3840,n5120" with force_bn codes (e.g. "160,224,256" = 1-CTA kernels)
    static int env_bn[3] = {-1, 0, 0};
    if (env_bn[0] < 0) {
      env_bn[0] = 0;
      if (const char* s = getenv("SAMRS_BN")) sscanf(s, "%d,%d,%d", &env_bn[0], &env_bn[1], &env_bn[2]);
    }
    const int ov = env_bn[p.N == 1280 ? 0 : (p.N == 3840 ? 1 : 2)];
    if (ov > 0) { bn = ov % 1000; pair = ov >= 1000 && ov < 2000; }
    static const bool use_mcast = getenv("SAMRS_GEMM_MCAST") != nullptr;
    mcast = use_mcast;
#endif
  } else if (force_bn == 0) {
    // pick the (kernel, N tile) with the smallest modelled duration = waves x per-tile time
    const int cands[4] = {256, 224, 160, 128};
    double best = 1e30;
    for (int pr = 0; pr < 2; ++pr) {
      if (pr == 1 && (p.batch != 1 || a_map_rank3 || p.M < 256 || (num_sms & 1))) continue;
      for (int i = 0; i < 4; ++i) {
        const int c = cands[i];
        const long tm = (p.M + (pr ? 255 : 127)) / (pr ? 256 : 128), tn = (p.N + c - 1) / c;
        const long units = pr ? num_sms / 2 : num_sms;
        const long waves = (tm * tn * p.batch + units - 1) / units;
        const double cost = double(waves) * tile_clk(c, p.K, pr != 0);
        if (cost < best - 1e-9) { best = cost; bn = c; pair = (pr != 0); }
      }
    }
  }
  p.tiles_m = (p.M + (pair ? 255 : 127)) / (pair ? 256 : 128);
  p.tiles_n = (p.N + bn - 1) / bn;
  CUtensorMap tA, tB;
  if (a_map_rank3) tA = *a_map_rank3;
  else SAMRS_TRY(make_tmap_2d(&tA, A, uint64_t(p.a_wrap_kb > 0 ? p.a_wrap_kb * GEMM_BK : p.K), uint64_t(p.M), uint64_t(lda) * 2, GEMM_BK, GEMM_BM));
  if (p.a_wrap_kb > 0 && (quad || streamk || mcast || a_map_rank3 || p.K != 3 * (p.a_wrap_kb / 2) * GEMM_BK || (p.a_wrap_kb & 1)))
    SAMRS_FAIL("gemm: a wrapped A operand is [hi | lo] of a 3-term split GEMM (K' = 3K, K a multiple of 64) on the plain tile schedules");
  if (mcast && (((p.M + 127) / 128) % 2 != 0 || (num_sms & 1) || a_map_rank3)) mcast = false;
  SAMRS_TRY(make_tmap_2d(&tB, B, uint64_t(p.K), uint64_t(p.N), uint64_t(ldb) * 2, GEMM_BK, uint32_t(quad ? bn / 4 : ((pair || mcast) ? bn / 2 : bn))));
  CUtensorMap tC;
  SAMRS_TRY(make_tmap_out(&tC, p.out, out_half, uint64_t(p.N), uint64_t(p.M), uint64_t(p.batch), uint64_t(p.ldc), uint64_t(p.out_batch_stride), 32));
  CUtensorMap tC16 = tC;
  if (bn % 32 != 0) {
    if (!pair || out_half || bn % 32 != 16) SAMRS_FAIL("gemm: tiles that are not a multiple of 32 wide exist for the fp32 CTA-pair kernel only");
    SAMRS_TRY(make_tmap_out(&tC16, p.out, false, uint64_t(p.N), uint64_t(p.M), uint64_t(p.batch), uint64_t(p.ldc), uint64_t(p.out_batch_stride), 16));
  }
  if (p.res != nullptr && (p.ldr % 4 != 0)) SAMRS_FAIL("gemm: residual leading dimension must be a multiple of 4");
  const int tiles = p.tiles_m * p.tiles_n * p.batch;
#ifndef SAMRS_EXPERIMENTS
  if (quad || streamk) SAMRS_FAIL("gemm: the stream-K and cluster-of-4 schedules exist in -DSAMRS_EXPERIMENTS builds only");
#else
  if (quad) {
    if (p.batch != 1 || a_map_rank3 || bn % 32 != 0 || (num_sms & 3)) SAMRS_FAIL("gemm: the cluster-of-4 kernel takes plain 2-D problems and 32-column tile multiples");
    const int supers = ((p.tiles_m + 1) / 2) * p.tiles_n;
#define SAMRS_GEMM4_CASE(BN_)                                                                       \
  if (bn == BN_) {                                                                                  \
    if (out_half && act == 0) return launch_gemm4_inst<BN_, true, 0>(tA, tB, tC, p, supers, num_sms, stream);     \
    if (out_half && act == 1) return launch_gemm4_inst<BN_, true, 1>(tA, tB, tC, p, supers, num_sms, stream);     \
    if (!out_half && act == 0) return launch_gemm4_inst<BN_, false, 0>(tA, tB, tC, p, supers, num_sms, stream);   \
    SAMRS_FAIL("gemm: unsupported epilogue");                                                       \
  }
    SAMRS_GEMM4_CASE(224)
    SAMRS_GEMM4_CASE(160)
    SAMRS_GEMM4_CASE(256)
#undef SAMRS_GEMM4_CASE
    SAMRS_FAIL("gemm: unsupported N tile");
  }
  if (streamk) {
    const int pairs = num_sms / 2, num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
    const bool ok = !out_half && act == 0 && p.accumulate && p.batch == 1 && !a_map_rank3 && p.sk_flags != nullptr && bn % 32 == 0 &&
                    p.res == nullptr && tiles <= SK_MAX_TILES && tiles >= pairs && (long(tiles) * num_kb) / pairs >= num_kb;
    if (!ok) {
      if (force_bn >= 3000) SAMRS_FAIL("gemm: stream-K needs the fp32 accumulate epilogue, engine counters and at least one tile per CTA pair");
      streamk = false;                                   // shape does not qualify: tile schedule of the same kernel family
    }
  }
  if (streamk) {
    const int grid2 = 2 * (num_sms / 2);
    if (bn == 256) return launch_gemm2_sk_inst<256>(tA, tB, tC, p, grid2, stream);
    if (bn == 160) return launch_gemm2_sk_inst<160>(tA, tB, tC, p, grid2, stream);
    if (bn == 128) return launch_gemm2_sk_inst<128>(tA, tB, tC, p, grid2, stream);
    SAMRS_FAIL("gemm: unsupported stream-K N tile");
  }
#endif
  if (pair) {
    const int pairs = num_sms / 2;
    const int grid2 = 2 * (tiles < pairs ? tiles : pairs);
#define SAMRS_GEMM2_CASE(BN_)                                                                       \
  if (bn == BN_) {                                                                                  \
    if (out_half && act == 0) return launch_gemm2_inst<BN_, true, 0>(tA, tB, tC, tC16, p, grid2, stream);     \
    if (out_half && act == 1) return launch_gemm2_inst<BN_, true, 1, GELU_EPI_WARPS>(tA, tB, tC, tC16, p, grid2, stream);     \
    if (!out_half && act == 0) return launch_gemm2_inst<BN_, false, 0>(tA, tB, tC, tC16, p, grid2, stream);   \
    SAMRS_FAIL("gemm: unsupported epilogue");                                                       \
  }
    if (bn == 144) {
      if (out_half || act != 0) SAMRS_FAIL("gemm: the 144-wide pair tile has the fp32 epilogue only");
      return launch_gemm2_inst<144, false, 0>(tA, tB, tC, tC16, p, grid2, stream);
    }
    SAMRS_GEMM2_CASE(256)
    SAMRS_GEMM2_CASE(224)
    SAMRS_GEMM2_CASE(160)
    SAMRS_GEMM2_CASE(128)
#undef SAMRS_GEMM2_CASE
    SAMRS_FAIL("gemm: unsupported N tile");
  }
#ifdef SAMRS_EXPERIMENTS
  if (mcast) {
    const int gridm = (tiles < num_sms ? tiles : num_sms) & ~1;
#define SAMRS_GEMMMC_CASE(BN_)                                                                        \
  if (bn == BN_) {                                                                                  \
    if (out_half && act == 0) return launch_gemm_mc_inst<BN_, true, 0>(tA, tB, tC, p, gridm, stream);   \
    if (out_half && act == 1) return launch_gemm_mc_inst<BN_, true, 1>(tA, tB, tC, p, gridm, stream);   \
    if (!out_half && act == 0) return launch_gemm_mc_inst<BN_, false, 0>(tA, tB, tC, p, gridm, stream); \
    SAMRS_FAIL("gemm: unsupported epilogue");                                                       \
  }
    SAMRS_GEMMMC_CASE(256)
    SAMRS_GEMMMC_CASE(224)
    SAMRS_GEMMMC_CASE(160)
#undef SAMRS_GEMMMC_CASE
    SAMRS_FAIL("gemm: unsupported N tile for the multicast kernel");
  }
#else
  if (mcast) SAMRS_FAIL("gemm: the TMA-multicast variant exists only in -DSAMRS_EXPERIMENTS builds");
#endif
  const int grid = tiles < num_sms ? tiles : num_sms;
  if (act == 3) {
    if (bn != 128 || out_half || p.hyper == nullptr || p.low == nullptr) SAMRS_FAIL("gemm: the up-scaling epilogue needs a 128-wide fp32 tile and its operands");
    return launch_gemm_inst<128, false, 3>(tA, tB, tC, p, grid, stream);
  }
#define SAMRS_GEMM_CASE(BN_)                                                                        \
  if (bn == BN_) {                                                                                  \
    if (out_half && act == 0) return launch_gemm_inst<BN_, true, 0>(tA, tB, tC, p, grid, stream);       \
    if (out_half && act == 1) return launch_gemm_inst<BN_, true, 1>(tA, tB, tC, p, grid, stream);       \
    if (!out_half && act == 0) return launch_gemm_inst<BN_, false, 0>(tA, tB, tC, p, grid, stream);     \
    SAMRS_FAIL("gemm: unsupported epilogue");                                                       \
  }
  SAMRS_GEMM_CASE(256)
  SAMRS_GEMM_CASE(224)
  SAMRS_GEMM_CASE(160)
  SAMRS_GEMM_CASE(128)
#undef SAMRS_GEMM_CASE
  SAMRS_FAIL("gemm: unsupported N tile");
}

// ------------------------------------------------------------------ per-category device timing (bench.py roofline)
struct ProfScope {
  cudaStream_t st; ProfRec r; bool active;
  ProfScope(int cat, cudaStream_t s) : st(s), active(t_ctx && t_ctx->prof.on && !t_ctx->capturing) {
    if (active) { r.cat = cat; r.a = t_ctx->prof.get(); r.b = t_ctx->prof.get(); cudaEventRecord(r.a, st); }
  }
  ~ProfScope() { if (active) { cudaEventRecord(r.b, st); t_ctx->prof.recs.push_back(r); } }
};

// ------------------------------------------------------------------ engine state
struct BlockWeights {
  float *ln1w, *ln1b, *ln2w, *ln2b;
  __half *wqkv, *wproj, *w1, *w2;
  float *bqkv_eff, *bproj_eff, *b1, *b2;
  __half* reltab;       // [256 or 64][hd] fp16: log2(e) * [rel_pos_h ; rel_pos_w ; 0]
  bool global;
};

struct DecAttn {  // decoder Attention weights (fp32)
  float *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo;
  int internal;
};
struct DecLayer {
  DecAttn self_attn, t2i, i2t;
  float *n1w, *n1b, *n2w, *n2b, *n3w, *n3b, *n4w, *n4b;
  float *m1w, *m1b, *m2w, *m2b;
};
struct Mlp3 { float *w[3], *b[3]; };

struct Engine {
  int device = 0;
  int D = 0, depth = 0, heads = 0, hd = 0;
  std::vector<int> global_idx;
  int num_sms = 148;
  LaunchCtx ctx;
  std::string err;
  std::vector<void*> allocs;             // activations and scratch: live until samrs_destroy or release()
  std::vector<void*> weight_allocs;      // everything load_weights_impl allocates: freed when weights are loaded again
  bool loading_weights = false;
  bool weights_loaded = false, image_set = false;
  // CUDA graphs of the two static-shape launch sequences (encode body; decode body per (prompts, tokens, flags)): the first
  // call of a shape runs eagerly (it also opts the kernels into large shared memory), the second captures on `cap_stream`,
  // later calls replay the instantiated graph into the caller's stream.  Kernels that touch caller-owned pointers (image
  // in, features / logits / IoU out, prompts) stay outside the graphs.  Profiling (per-launch events) runs eagerly.
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int launches = 0; int eager_runs = 0; };
  bool graphs_enabled = true;
  cudaStream_t cap_stream = nullptr;
  // second stream of the decode body: the token-side chain of the next layer and the up-scaling LayerNorm run beside the
  // image-side kernels they do not depend on (fork / join through these events; inside a capture they become parallel graph branches)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
  GraphEntry enc_graph;
  std::unordered_map<uint64_t, GraphEntry> dec_graphs;
  void drop_graphs() {
    if (enc_graph.exec) cudaGraphExecDestroy(enc_graph.exec);
    enc_graph = GraphEntry();
    for (auto& kv : dec_graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    dec_graphs.clear();
  }

  // encoder weights
  __half* w_patch = nullptr; float* b_patch = nullptr; float* pos_embed = nullptr;
  std::vector<BlockWeights> blocks;
  __half* w_neck0 = nullptr; __half* w_neck2 = nullptr;
  float *neck1w = nullptr, *neck1b = nullptr, *neck3w = nullptr, *neck3b = nullptr;
  // prompt encoder
  float *gauss = nullptr, *point_emb = nullptr, *not_a_point = nullptr, *no_mask = nullptr;
  float *md_w0, *md_b0, *md_g1, *md_be1, *md_w3, *md_b3, *md_g4, *md_be4, *md_w6, *md_b6;
  // decoder
  DecLayer dl[2];
  DecAttn final_attn;
  float *nfw, *nfb, *iou_token, *mask_tokens;
  float *up_w1r, *up_b1r, *up_lnw, *up_lnb, *up_w2r, *up_b2;
  Mlp3 hyper[4], iou_head;
  // split-fp16 ([hi|hi|lo] x 256) weights of the four tensor-core decoder GEMMs and their epilogue operands
  __half *wd_p1 = nullptr, *wd_p2 = nullptr, *wd_o0 = nullptr, *wd_o1 = nullptr, *wd_up2 = nullptr;
  float* up_b2r = nullptr;             // ConvT2 bias tiled over the four output sub-positions [4][32]
  float *bias_p1 = nullptr, *bias_p2 = nullptr, *R1 = nullptr, *R2 = nullptr;
  float* dense_pe = nullptr;           // [4096][256]
  float* pek[5] = {nullptr};           // dense_pe * W^T for: l0.t2i.k, l0.i2t.q, l1.t2i.k, l1.i2t.q, final.k  [4096][128]

  // encoder activations
  __half *a_pe, *xn, *qkv, *attn_o, *hid, *x16, *neck_ln16, *neck_col;
  float *x, *rel, *neck0, *neck2, *feat_tok, *feat_nchw;
  // per-image decoder cache
  float *src0, *KVQ0;                  // KVQ0 [4096][384] = layer-0 [K (t2i) | V (t2i) | Q (i2t)] of the image tokens
  float *b_kvq0 = nullptr, *R_kvq0 = nullptr;                      // their concatenated biases / positional terms
  __half *wd_kvq0 = nullptr, *src0A = nullptr;                     // split-fp16 [Wk ; Wv ; Wq] (384 x 768 = [hi | hi | lo]) and src0 (4096 x 512 = [hi | lo])
  float* pp_full = nullptr;            // postprocess scratch for non-1024 sizes
  struct ResizeTab { int in, out, ksize; int* bounds; int* kk; std::vector<int> h_bounds, h_kk; unsigned long long used; };
  std::vector<ResizeTab> resize_tabs;   // Pillow tap tables per (input size, output size), built on first use; LRU of 32
  unsigned long long resize_clock = 0;
  uint8_t* resize_tmp = nullptr;       // horizontal-pass result [H][out_w][3]
  size_t resize_tmp_bytes = 0;
  uint8_t* rbox_mask = nullptr;        // rotated-box rasteriser scratch: [B][H][W] fill masks
  size_t rbox_mask_bytes = 0;
  uint32_t* rle_packed = nullptr;      // run-length encoder scratch: column-major bit planes [B][ceil(H/32)][W]
  size_t rle_packed_words = 0;
  long long* rle_runs = nullptr;       // [rle_runs_cap] runs per mask
  long long* rle_tstate = nullptr;     // [rle_runs_cap][1024][2] per-thread (transitions, last position) of the count pass
  int rle_runs_cap = 0;
  float* d_t2i_part = nullptr;         // [cap][8][16][8][18] key-chunk partials of the token->image attention
  // decoder scratch (sized for dec_cap prompts)
  int dec_cap = 0;
  float *d_tok0 = nullptr, *d_q = nullptr, *d_qkv = nullptr /*[BT][768] self-attention q | k | v*/, *d_tmp256d = nullptr;
  float *d_tmp128a = nullptr, *d_tmp128b = nullptr, *d_tmp128c = nullptr, *d_mlp = nullptr;
  float *d_keys = nullptr, *d_P = nullptr, *d_hyper = nullptr, *d_hy_a = nullptr, *d_hy_b = nullptr, *d_iou_all = nullptr, *d_low = nullptr;
  __half *d_keysA = nullptr, *d_ioA = nullptr, *d_up1 = nullptr;   // split-fp16 [hi | lo] operands; d_up1: GELU(LN(ConvT1)) [cap*16384][128]
  int mask_cap = 0;                    // mask-prompt path scratch (per-prompt layer-0 operands), allocated on first use
  float *d_Kp = nullptr, *d_Vp = nullptr, *d_Qp = nullptr, *d_src = nullptr;

  template <typename T>
  int alloc(T** p, size_t n) {
    void* q = nullptr;
    if (cudaMalloc(&q, n * sizeof(T) + 256) != cudaSuccess) return samrs::fail(__FILE__, __LINE__, "cudaMalloc failed");
    (loading_weights ? weight_allocs : allocs).push_back(q);
    *p = reinterpret_cast<T*>(q);
    return 0;
  }
  // give a superseded scratch buffer back (cudaFree waits for the device: only reached when a buffer has to grow)
  template <typename T>
  void release(T** p) {
    if (*p == nullptr) return;
    void* q = reinterpret_cast<void*>(*p);
    for (size_t i = 0; i < allocs.size(); ++i)
      if (allocs[i] == q) { allocs[i] = allocs.back(); allocs.pop_back(); cudaFree(q); break; }
    drop_graphs();                                   // captured launches may hold the old address
    *p = nullptr;
  }
  void free_weights() {
    if (weight_allocs.empty()) return;
    cudaDeviceSynchronize();
    drop_graphs();
    for (void* q : weight_allocs) cudaFree(q);
    weight_allocs.clear();
    weights_loaded = image_set = false;
  }
};

static int set_err(Engine* e, int rc) {
  if (rc != 0 && e) e->err = g_last_error;
  return rc;
}

// Runs `body(stream)` eagerly on `st`, or - from the third call of this shape on - replays its captured CUDA graph in `st`.
template <class F>
static int run_graphed(Engine* e, Engine::GraphEntry& ge, cudaStream_t st, F&& body) {
  const bool can = e->graphs_enabled && e->cap_stream != nullptr && t_ctx != nullptr && !t_ctx->prof.on;
  if (!can) return body(st);
  if (ge.eager_runs == 0) {                            // first call: eager (kernel attributes, lazily built tables)
    ge.eager_runs = 1;
    return body(st);
  }
  if (ge.exec == nullptr) {
    const int64_t l0 = t_ctx->launches;
    cudaGraph_t g = nullptr;
    int rc = 1;
    if (cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      t_ctx->capturing = true;
      rc = body(e->cap_stream);
      t_ctx->capturing = false;
      if (cudaStreamEndCapture(e->cap_stream, &g) != cudaSuccess) rc = 1;
    }
    const int n = int(t_ctx->launches - l0);
    t_ctx->launches = l0;
    if (rc == 0 && g != nullptr && cudaGraphInstantiate(&ge.exec, g, 0) != cudaSuccess) { ge.exec = nullptr; rc = 1; }
    if (g != nullptr) cudaGraphDestroy(g);
    if (rc != 0) {                                     // capture is an optimisation: fall back to direct launches for good
      cudaGetLastError();
      e->graphs_enabled = false;
      ge.exec = nullptr;
      return body(st);
    }
    ge.launches = n;
  }
  SAMRS_CUDA_OK(cudaGraphLaunch(ge.exec, st));
  t_ctx->launches += ge.launches;
  return 0;
}

#ifndef SAMRS_SGEMM_NARROW
#define SAMRS_SGEMM_NARROW 1
#endif
constexpr bool SGEMM_NARROW = SAMRS_SGEMM_NARROW != 0;
static void* g_attn_dbg = nullptr;            // device buffer for attention pipeline traces (tools only)
// ------------------------------------------------------------------ small launch helpers
static int sgemm(cudaStream_t st, const float* A, int lda, const float* W, int ldw, float* C, int ldc, const float* bias,
                 const float* R, int ldr, int rmod, int M, int N, int K, int act) {
  if (K % 16 != 0 || lda % 4 != 0 || ldw % 4 != 0) SAMRS_FAIL("sgemm: K must be a multiple of 16");
  SgemmParams p{A, lda, W, ldw, C, ldc, bias, R, ldr, rmod, M, N, K, act};
  if (M <= 2048 && K % 64 == 0) {
    // token-side GEMM: latency-bound; one shared-memory panel of at most 256 of K per block, larger K split across
    // blockIdx.z (deterministic reduce)
    float* ws = t_ctx ? t_ctx->splitk_ws : nullptr;
    int splits = (K + SGT_KMAX - 1) / SGT_KMAX;
    if (splits > 1 && !(ws && size_t(M) * N * splits <= t_ctx->splitk_ws_floats)) SAMRS_FAIL("sgemm: split-K workspace too small");
    const int kps = ((K / splits + 15) / 16) * 16;
    if (kps > SGT_KMAX || kps * splits < K) SAMRS_FAIL("sgemm: K does not split into panels of at most 256");
    // 32 x 32 output tiles, or 32 x 16 when that would leave most of the GPU without a block (N = 128 / 256 with a few hundred rows)
    const bool narrow = SGEMM_NARROW && ((M + 31) / 32) * ((N + 31) / 32) * splits < 148 && N % 16 == 0;
    if (narrow) {
      SAMRS_TRY(opt_in_smem(sgemm_small_kernel<16>, SGT_SMEM));
      sgemm_small_kernel<16><<<dim3((M + 31) / 32, N / 16, splits), 128, SGT_SMEM, st>>>(p, kps, splits > 1 ? ws : nullptr);
    } else {
      SAMRS_TRY(opt_in_smem(sgemm_small_kernel<32>, SGT_SMEM));
      sgemm_small_kernel<32><<<dim3((M + 31) / 32, (N + 31) / 32, splits), 128, SGT_SMEM, st>>>(p, kps, splits > 1 ? ws : nullptr);
    }
    SAMRS_CUDA_OK(cudaGetLastError());
    count_launch();
    if (splits > 1) {
      splitk_reduce_kernel<<<(M * N + 255) / 256, 256, 0, st>>>(p, ws, splits);
      SAMRS_CUDA_OK(cudaGetLastError());
      count_launch();
    }
    return 0;
  }
  dim3 grid((M + 127) / 128, (N + 63) / 64);
  sgemm_tn_kernel<<<grid, 256, 0, st>>>(p);
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

template <typename OutT, int ACT>
static int ln_rows(cudaStream_t st, const float* in, int ld_in, const float* g, const float* b, float eps, OutT* out, int ld_out,
                   int rows, int C) {
  const int nv = C / 4;
  const int threads = 256, rows_per_block = threads / 32;
  const int grid = (rows + rows_per_block - 1) / rows_per_block;
  if (C % 4 != 0) SAMRS_FAIL("layernorm: C must be a multiple of 4");
  if constexpr (sizeof(OutT) == 2 && ACT == 0) {
    if (rows >= 1024 && (C == 1280 || C == 1024 || C == 768)) {
      const int g4 = (rows + 3) / 4;
      __half* o16 = reinterpret_cast<__half*>(out);
      if (C == 1280) SAMRS_CUDA_OK(launch_pdl(ln_rows_stream_kernel<10>, dim3(g4), dim3(64), 0, st, in, ld_in, g, b, eps, o16, ld_out, rows));
      else if (C == 1024) SAMRS_CUDA_OK(launch_pdl(ln_rows_stream_kernel<8>, dim3(g4), dim3(64), 0, st, in, ld_in, g, b, eps, o16, ld_out, rows));
      else SAMRS_CUDA_OK(launch_pdl(ln_rows_stream_kernel<6>, dim3(g4), dim3(64), 0, st, in, ld_in, g, b, eps, o16, ld_out, rows));
      count_launch();
      return 0;
    }
  }
  if (nv <= 32) ln_rows_kernel<OutT, ACT, 1><<<grid, threads, 0, st>>>(in, ld_in, g, b, eps, out, ld_out, rows, C);
  else if (nv <= 64) ln_rows_kernel<OutT, ACT, 2><<<grid, threads, 0, st>>>(in, ld_in, g, b, eps, out, ld_out, rows, C);
  else if (nv <= 192) ln_rows_kernel<OutT, ACT, 6><<<grid, threads, 0, st>>>(in, ld_in, g, b, eps, out, ld_out, rows, C);
  else if (nv <= 320) ln_rows_kernel<OutT, ACT, 10><<<grid, threads, 0, st>>>(in, ld_in, g, b, eps, out, ld_out, rows, C);
  else SAMRS_FAIL("layernorm: C too large");
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

template <int HD, int BX, int QBY, int KBY, int NKT>
static int launch_attn2_inst(const CUtensorMap& tQ, const CUtensorMap& tKV, const CUtensorMap& tT, const AttnParams& p, int num_sms, cudaStream_t st) {
  using C = Attn2Cfg<HD, BX, QBY, KBY, NKT>;
  SAMRS_TRY(opt_in_smem(attn_tc2_kernel<HD, BX, QBY, KBY, NKT>, C::kSmemBytes));
  const int units = p.num_qtiles * p.heads;
  const int grid = units < num_sms ? units : num_sms;
  SAMRS_CUDA_OK(launch_pdl(attn_tc2_kernel<HD, BX, QBY, KBY, NKT>, dim3(grid), dim3(384), C::kSmemBytes, st, tQ, tKV, tT, p));
  count_launch();
  return 0;
}

// encoder attention of one block.  qkv: [4096][3D] fp16.  Windowed blocks form their rel-pos terms inside the attention
// kernel (table in shared memory); global blocks first run the head-batched rel-pos GEMM whose output the kernel gathers.
static int encoder_attention(Engine* e, cudaStream_t st, const __half* qkv, const __half* reltab, bool global, __half* out) {
  const int D = e->D, hd = e->hd;
  if (hd != 64 && hd != 80) SAMRS_FAIL("head_dim must be 64 or 80");
  const float scale_log2e = (1.0f / sqrtf(float(hd))) * 1.4426950408889634f;
  if (global) {
    // decomposed rel-pos terms for every (head, token): G = q . [rel_pos_h ; rel_pos_w]^T as one head-batched
    // tensor-core GEMM (A = the q columns of the qkv activation through a rank-3 tensor map); the epilogue writes
    // fp16(G / scale_log2e), the value the attention kernel feeds to its bias MMA
    ProfScope ps(PC_RELPOS, st);
    const int NP = 256;
    CUtensorMap tA;
    SAMRS_TRY(make_tmap_3d(&tA, qkv, uint64_t(hd), uint64_t(e->heads), 4096, uint64_t(hd) * 2, uint64_t(3 * D) * 2, GEMM_BK, 1, GEMM_BM));
    GemmParams gp;
    gp.M = 4096; gp.N = NP; gp.K = hd; gp.out = e->rel; gp.ldc = NP; gp.bias = nullptr; gp.res = nullptr; gp.ldr = 0; gp.res_mod = 0;
    gp.tiles_m = gp.tiles_n = 0; gp.batch = e->heads; gp.a_rank3 = 1; gp.out_batch_stride = (long long)4096 * NP; gp.dbg = nullptr; gp.dbg_mode = 0; gp.accumulate = 0;
    gp.out_scale = 1.0f / scale_log2e;
    SAMRS_TRY(launch_gemm_tc(nullptr, 8, reltab, hd, gp, true, 0, e->num_sms, st, 256, &tA));
  }
  ProfScope ps2(global ? PC_ATTN_GLOB : PC_ATTN_WIN, st);
  AttnParams p;
  p.rel = e->rel;
  p.rel16 = reinterpret_cast<const __half*>(e->rel);
  p.out = out;
  p.D = D;
  p.heads = e->heads;
  p.scale_log2e = scale_log2e;
  p.rel_scale = 1.0f / scale_log2e;
  p.dbg = static_cast<unsigned long long*>(g_attn_dbg);
  p.pv_split = 0;
  CUtensorMap tQ, tKV, tT;
  const uint64_t pitch_x = uint64_t(3 * D) * 2, pitch_y = pitch_x * 64;
  if (global) {
    p.num_qtiles = 16;                                                // pairs of 128-query tiles
    SAMRS_TRY(make_tmap_3d(&tQ, qkv, uint64_t(3 * D), 64, 64, pitch_x, pitch_y, 64, 64, 2));
    tKV = tQ;
    tT = tQ;                                                          // unused by the global variant
    if (hd == 64) return launch_attn2_inst<64, 64, 2, 2, 32>(tQ, tKV, tT, p, e->num_sms, st);
    return launch_attn2_inst<80, 64, 2, 2, 32>(tQ, tKV, tT, p, e->num_sms, st);
  }
  p.num_qtiles = 25;                                                  // windows (two 7-row halves each)
  SAMRS_TRY(make_tmap_3d(&tQ, qkv, uint64_t(3 * D), 64, 64, pitch_x, pitch_y, 64, 14, 7));
  SAMRS_TRY(make_tmap_3d(&tKV, qkv, uint64_t(3 * D), 64, 64, pitch_x, pitch_y, 64, 14, 14));
  SAMRS_TRY(make_tmap_2d(&tT, reltab, uint64_t(hd), 64, uint64_t(hd) * 2, 64, 64));   // [27 + 27 + 10 zero rows][hd]
  if (hd == 64) return launch_attn2_inst<64, 14, 7, 14, 1>(tQ, tKV, tT, p, e->num_sms, st);
  return launch_attn2_inst<80, 14, 7, 14, 1>(tQ, tKV, tT, p, e->num_sms, st);
}

// ------------------------------------------------------------------ weight ingest
struct Src { const float* p; int64_t n; };
typedef std::unordered_map<std::string, Src> SrcMap;

static int need(const SrcMap& m, const std::string& k, int64_t n, const float** out) {
  auto it = m.find(k);
  if (it == m.end()) return samrs::fail(__FILE__, __LINE__, ("missing key in state_dict: " + k).c_str());
  if (it->second.n != n) return samrs::fail(__FILE__, __LINE__, ("size mismatch for " + k).c_str());
  *out = it->second.p;
  return 0;
}

__global__ void zero_range_kernel(float* p, int from, int to) {
  const int i = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < to) p[i] = 0.f;
}
// neck.2.weight [256][256][3][3] -> [256][tap*256 + ci] fp16
__global__ void repack_neck3x3_kernel(const float* __restrict__ w, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 256 * 2304) return;
  const int co = i / 2304, r = i % 2304, tap = r / 256, ci = r % 256;
  out[i] = __float2half_rn(w[(size_t(co) * 256 + ci) * 9 + tap]);
}
// ConvTranspose2d weight [Cin][Cout][2][2] -> [(dy*2+dx)*Cout + co][ci]
__global__ void repack_convT_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * Cout * Cin) return;
  const int ci = i % Cin, co = (i / Cin) % Cout, d = i / (Cin * Cout);
  out[i] = w[(size_t(ci) * Cout + co) * 4 + d];
}
__global__ void tile_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int C, int reps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * reps) out[i] = b[i % C];
}

// rel-pos table for the batched GEMM: rows [0,2S-1) = log2e*rel_pos_h, [2S-1,4S-2) = log2e*rel_pos_w, zero padded
__global__ void build_reltab_kernel(const float* __restrict__ rph, const float* __restrict__ rpw, int S, int hd, int NP,
                                    __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NP * hd) return;
  const int r = i / hd, c = i % hd, L = 2 * S - 1;
  float v = 0.f;
  if (r < L) v = rph[r * hd + c];
  else if (r < 2 * L) v = rpw[(r - L) * hd + c];
  out[i] = __float2half_rn(v * 1.4426950408889634f);
}
struct Engine;
static int build_reltab(Engine* e, cudaStream_t st, const float* rph, const float* rpw, int S, __half** out);

static int copy_f32(Engine* e, cudaStream_t st, const float* src, int64_t n, float** dst) {
  SAMRS_TRY(e->alloc(dst, size_t(n)));
  SAMRS_CUDA_OK(cudaMemcpyAsync(*dst, src, size_t(n) * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}
static int to_half(Engine* e, cudaStream_t st, const float* src, int64_t n, __half** dst) {
  if (n % 4 != 0) SAMRS_FAIL("to_half: size must be a multiple of 4");
  SAMRS_TRY(e->alloc(dst, size_t(n)));
  const size_t n4 = size_t(n) / 4;
  cast_f32_f16_kernel<<<unsigned((n4 + 255) / 256), 256, 0, st>>>(src, *dst, n4);
  SAMRS_CUDA_OK(cudaGetLastError());
  return 0;
}

static int build_reltab(Engine* e, cudaStream_t st, const float* rph, const float* rpw, int S, __half** out) {
  const int NP = (S == 64) ? 256 : 64;
  SAMRS_TRY(e->alloc(out, size_t(NP) * e->hd));
  build_reltab_kernel<<<(NP * e->hd + 255) / 256, 256, 0, st>>>(rph, rpw, S, e->hd, NP, *out);
  SAMRS_CUDA_OK(cudaGetLastError());
  return 0;
}

static int load_dec_attn(Engine* e, cudaStream_t st, const SrcMap& m, const std::string& pre, int internal, DecAttn* a) {
  // q | k | v projection weights and biases are stored back to back ([3 * internal][256], [3 * internal])
  const float* s;
  a->internal = internal;
  const int64_t nw = int64_t(internal) * 256;
  float *wbase = nullptr, *bbase = nullptr;
  SAMRS_TRY(e->alloc(&wbase, size_t(3 * nw)));
  SAMRS_TRY(e->alloc(&bbase, size_t(3 * internal)));
  a->wq = wbase; a->wk = wbase + nw; a->wv = wbase + 2 * nw;
  a->bq = bbase; a->bk = bbase + internal; a->bv = bbase + 2 * internal;
  const char* names[3] = {".q_proj", ".k_proj", ".v_proj"};
  for (int i = 0; i < 3; ++i) {
    SAMRS_TRY(need(m, pre + names[i] + ".weight", nw, &s));
    SAMRS_CUDA_OK(cudaMemcpyAsync(wbase + i * nw, s, size_t(nw) * 4, cudaMemcpyDeviceToDevice, st));
    SAMRS_TRY(need(m, pre + names[i] + ".bias", internal, &s));
    SAMRS_CUDA_OK(cudaMemcpyAsync(bbase + i * internal, s, size_t(internal) * 4, cudaMemcpyDeviceToDevice, st));
  }
  SAMRS_TRY(need(m, pre + ".out_proj.weight", nw, &s)); SAMRS_TRY(copy_f32(e, st, s, nw, &a->wo));
  SAMRS_TRY(need(m, pre + ".out_proj.bias", 256, &s));  SAMRS_TRY(copy_f32(e, st, s, 256, &a->bo));
  return 0;
}
static int load_pair(Engine* e, cudaStream_t st, const SrcMap& m, const std::string& pre, int64_t n, float** w, float** b) {
  const float* s;
  SAMRS_TRY(need(m, pre + ".weight", n, &s)); SAMRS_TRY(copy_f32(e, st, s, n, w));
  SAMRS_TRY(need(m, pre + ".bias", n, &s));   SAMRS_TRY(copy_f32(e, st, s, n, b));
  return 0;
}
static int load_mlp3(Engine* e, cudaStream_t st, const SrcMap& m, const std::string& pre, int out_dim, Mlp3* mlp) {
  const int dims[4] = {256, 256, 256, out_dim};
  for (int j = 0; j < 3; ++j) {
    const float* s;
    const std::string k = pre + ".layers." + std::to_string(j);
    SAMRS_TRY(need(m, k + ".weight", int64_t(dims[j + 1]) * dims[j], &s)); SAMRS_TRY(copy_f32(e, st, s, int64_t(dims[j + 1]) * dims[j], &mlp->w[j]));
    SAMRS_TRY(need(m, k + ".bias", dims[j + 1], &s));                      SAMRS_TRY(copy_f32(e, st, s, dims[j + 1], &mlp->b[j]));
  }
  return 0;
}

static int load_weights_impl(Engine* e, const SrcMap& m, cudaStream_t st) {
  const int D = e->D;
  const int64_t D64 = D;
  const float* s;
  const std::string ie = "image_encoder.";
  SAMRS_TRY(need(m, ie + "pos_embed", 4096 * D64, &s));               SAMRS_TRY(copy_f32(e, st, s, 4096 * D64, &e->pos_embed));
  SAMRS_TRY(need(m, ie + "patch_embed.proj.weight", D64 * 768, &s));  SAMRS_TRY(to_half(e, st, s, D64 * 768, &e->w_patch));
  SAMRS_TRY(need(m, ie + "patch_embed.proj.bias", D64, &s));          SAMRS_TRY(copy_f32(e, st, s, D64, &e->b_patch));
  e->blocks.resize(e->depth);
  for (int i = 0; i < e->depth; ++i) {
    BlockWeights& b = e->blocks[i];
    const std::string p = ie + "blocks." + std::to_string(i) + ".";
    b.global = false;
    for (int gidx : e->global_idx) b.global |= (gidx == i);
    const int S = b.global ? 64 : 14;
    SAMRS_TRY(load_pair(e, st, m, p + "norm1", D64, &b.ln1w, &b.ln1b));
    SAMRS_TRY(load_pair(e, st, m, p + "norm2", D64, &b.ln2w, &b.ln2b));
    const float *rph, *rpw;
    SAMRS_TRY(need(m, p + "attn.rel_pos_h", int64_t(2 * S - 1) * e->hd, &rph));
    SAMRS_TRY(need(m, p + "attn.rel_pos_w", int64_t(2 * S - 1) * e->hd, &rpw));
    SAMRS_TRY(build_reltab(e, st, rph, rpw, S, &b.reltab));
    SAMRS_TRY(need(m, p + "attn.qkv.weight", 3 * D64 * D64, &s));             SAMRS_TRY(to_half(e, st, s, 3 * D64 * D64, &b.wqkv));
    const float* bqkv;
    SAMRS_TRY(need(m, p + "attn.qkv.bias", 3 * D64, &bqkv));
    // K bias cancels in the softmax, V bias moves into the proj bias (see attn_tc.cuh header)
    SAMRS_TRY(copy_f32(e, st, bqkv, 3 * D64, &b.bqkv_eff));
    zero_range_kernel<<<(2 * D + 255) / 256, 256, 0, st>>>(b.bqkv_eff, D, 3 * D);
    const float *wproj, *bproj;
    SAMRS_TRY(need(m, p + "attn.proj.weight", D64 * D64, &wproj));            SAMRS_TRY(to_half(e, st, wproj, D64 * D64, &b.wproj));
    SAMRS_TRY(need(m, p + "attn.proj.bias", D64, &bproj));
    SAMRS_TRY(e->alloc(&b.bproj_eff, size_t(D)));
    // bproj_eff = bproj + Wproj * bv     (fp32, one-row GEMM)
    SAMRS_TRY(sgemm(st, bqkv + 2 * D, D, wproj, D, b.bproj_eff, D, bproj, nullptr, 0, 0, 1, D, D, 0));
    SAMRS_TRY(need(m, p + "mlp.lin1.weight", 4 * D64 * D64, &s));             SAMRS_TRY(to_half(e, st, s, 4 * D64 * D64, &b.w1));
    SAMRS_TRY(need(m, p + "mlp.lin1.bias", 4 * D64, &s));                     SAMRS_TRY(copy_f32(e, st, s, 4 * D64, &b.b1));
    SAMRS_TRY(need(m, p + "mlp.lin2.weight", 4 * D64 * D64, &s));             SAMRS_TRY(to_half(e, st, s, 4 * D64 * D64, &b.w2));
    SAMRS_TRY(need(m, p + "mlp.lin2.bias", D64, &s));                         SAMRS_TRY(copy_f32(e, st, s, D64, &b.b2));
  }
  SAMRS_TRY(need(m, ie + "neck.0.weight", 256 * D64, &s));       SAMRS_TRY(to_half(e, st, s, 256 * D64, &e->w_neck0));
  SAMRS_TRY(load_pair(e, st, m, ie + "neck.1", 256, &e->neck1w, &e->neck1b));
  SAMRS_TRY(need(m, ie + "neck.2.weight", 256 * 2304, &s));
  SAMRS_TRY(e->alloc(&e->w_neck2, size_t(256) * 2304));
  repack_neck3x3_kernel<<<(256 * 2304 + 255) / 256, 256, 0, st>>>(s, e->w_neck2);
  SAMRS_TRY(load_pair(e, st, m, ie + "neck.3", 256, &e->neck3w, &e->neck3b));

  const std::string pe = "prompt_encoder.";
  SAMRS_TRY(need(m, pe + "pe_layer.positional_encoding_gaussian_matrix", 256, &s)); SAMRS_TRY(copy_f32(e, st, s, 256, &e->gauss));
  SAMRS_TRY(e->alloc(&e->point_emb, 4 * 256));
  for (int i = 0; i < 4; ++i) {
    SAMRS_TRY(need(m, pe + "point_embeddings." + std::to_string(i) + ".weight", 256, &s));
    SAMRS_CUDA_OK(cudaMemcpyAsync(e->point_emb + i * 256, s, 1024, cudaMemcpyDeviceToDevice, st));
  }
  SAMRS_TRY(need(m, pe + "not_a_point_embed.weight", 256, &s)); SAMRS_TRY(copy_f32(e, st, s, 256, &e->not_a_point));
  SAMRS_TRY(need(m, pe + "no_mask_embed.weight", 256, &s));     SAMRS_TRY(copy_f32(e, st, s, 256, &e->no_mask));
  SAMRS_TRY(need(m, pe + "mask_downscaling.0.weight", 16, &s)); SAMRS_TRY(copy_f32(e, st, s, 16, &e->md_w0));
  SAMRS_TRY(need(m, pe + "mask_downscaling.0.bias", 4, &s));    SAMRS_TRY(copy_f32(e, st, s, 4, &e->md_b0));
  SAMRS_TRY(load_pair(e, st, m, pe + "mask_downscaling.1", 4, &e->md_g1, &e->md_be1));
  SAMRS_TRY(need(m, pe + "mask_downscaling.3.weight", 256, &s)); SAMRS_TRY(copy_f32(e, st, s, 256, &e->md_w3));
  SAMRS_TRY(need(m, pe + "mask_downscaling.3.bias", 16, &s));    SAMRS_TRY(copy_f32(e, st, s, 16, &e->md_b3));
  SAMRS_TRY(load_pair(e, st, m, pe + "mask_downscaling.4", 16, &e->md_g4, &e->md_be4));
  SAMRS_TRY(need(m, pe + "mask_downscaling.6.weight", 4096, &s)); SAMRS_TRY(copy_f32(e, st, s, 4096, &e->md_w6));
  SAMRS_TRY(need(m, pe + "mask_downscaling.6.bias", 256, &s));    SAMRS_TRY(copy_f32(e, st, s, 256, &e->md_b6));

  const std::string md = "mask_decoder.";
  for (int i = 0; i < 2; ++i) {
    DecLayer& L = e->dl[i];
    const std::string p = md + "transformer.layers." + std::to_string(i);
    SAMRS_TRY(load_dec_attn(e, st, m, p + ".self_attn", 256, &L.self_attn));
    SAMRS_TRY(load_dec_attn(e, st, m, p + ".cross_attn_token_to_image", 128, &L.t2i));
    SAMRS_TRY(load_dec_attn(e, st, m, p + ".cross_attn_image_to_token", 128, &L.i2t));
    SAMRS_TRY(load_pair(e, st, m, p + ".norm1", 256, &L.n1w, &L.n1b));
    SAMRS_TRY(load_pair(e, st, m, p + ".norm2", 256, &L.n2w, &L.n2b));
    SAMRS_TRY(load_pair(e, st, m, p + ".norm3", 256, &L.n3w, &L.n3b));
    SAMRS_TRY(load_pair(e, st, m, p + ".norm4", 256, &L.n4w, &L.n4b));
    SAMRS_TRY(need(m, p + ".mlp.lin1.weight", 2048 * 256, &s)); SAMRS_TRY(copy_f32(e, st, s, 2048 * 256, &L.m1w));
    SAMRS_TRY(need(m, p + ".mlp.lin1.bias", 2048, &s));         SAMRS_TRY(copy_f32(e, st, s, 2048, &L.m1b));
    SAMRS_TRY(need(m, p + ".mlp.lin2.weight", 2048 * 256, &s)); SAMRS_TRY(copy_f32(e, st, s, 2048 * 256, &L.m2w));
    SAMRS_TRY(need(m, p + ".mlp.lin2.bias", 256, &s));          SAMRS_TRY(copy_f32(e, st, s, 256, &L.m2b));
  }
  SAMRS_TRY(load_dec_attn(e, st, m, md + "transformer.final_attn_token_to_image", 128, &e->final_attn));
  SAMRS_TRY(load_pair(e, st, m, md + "transformer.norm_final_attn", 256, &e->nfw, &e->nfb));
  SAMRS_TRY(need(m, md + "iou_token.weight", 256, &s));    SAMRS_TRY(copy_f32(e, st, s, 256, &e->iou_token));
  SAMRS_TRY(need(m, md + "mask_tokens.weight", 1024, &s)); SAMRS_TRY(copy_f32(e, st, s, 1024, &e->mask_tokens));
  const float *w1, *b1, *w2;
  SAMRS_TRY(need(m, md + "output_upscaling.0.weight", 256 * 64 * 4, &w1));
  SAMRS_TRY(need(m, md + "output_upscaling.0.bias", 64, &b1));
  SAMRS_TRY(e->alloc(&e->up_w1r, 256 * 256));
  SAMRS_TRY(e->alloc(&e->up_b1r, 256));
  repack_convT_kernel<<<(4 * 64 * 256 + 255) / 256, 256, 0, st>>>(w1, e->up_w1r, 256, 64);
  tile_bias_kernel<<<1, 256, 0, st>>>(b1, e->up_b1r, 64, 4);
  SAMRS_TRY(load_pair(e, st, m, md + "output_upscaling.1", 64, &e->up_lnw, &e->up_lnb));
  SAMRS_TRY(need(m, md + "output_upscaling.3.weight", 64 * 32 * 4, &w2));
  SAMRS_TRY(e->alloc(&e->up_w2r, 4 * 32 * 64));
  repack_convT_kernel<<<(4 * 32 * 64 + 255) / 256, 256, 0, st>>>(w2, e->up_w2r, 64, 32);
  SAMRS_TRY(need(m, md + "output_upscaling.3.bias", 32, &s)); SAMRS_TRY(copy_f32(e, st, s, 32, &e->up_b2));
  for (int i = 0; i < 4; ++i) SAMRS_TRY(load_mlp3(e, st, m, md + "output_hypernetworks_mlps." + std::to_string(i), 32, &e->hyper[i]));
  SAMRS_TRY(load_mlp3(e, st, m, md + "iou_prediction_head", 4, &e->iou_head));

  // constants of the model: dense positional encoding and its five 256->128 projections
  SAMRS_TRY(e->alloc(&e->dense_pe, size_t(4096) * 256));
  dense_pe_kernel<<<4096, 128, 0, st>>>(e->gauss, e->dense_pe);
  const float* pw[5] = {e->dl[0].t2i.wk, e->dl[0].i2t.wq, e->dl[1].t2i.wk, e->dl[1].i2t.wq, e->final_attn.wk};
  for (int i = 0; i < 5; ++i) {
    SAMRS_TRY(e->alloc(&e->pek[i], size_t(4096) * 128));
    SAMRS_TRY(sgemm(st, e->dense_pe, 256, pw[i], 256, e->pek[i], 128, nullptr, nullptr, 0, 0, 4096, 128, 256, 0));
  }
  // layer-0 projections of the (prompt-independent) image tokens as one GEMM: [Wk ; Wv ; Wq], [bk ; bv ; bq], [pe Wk | 0 | pe Wq]
  SAMRS_TRY(e->alloc(&e->wd_kvq0, size_t(384) * 768));
  SAMRS_TRY(e->alloc(&e->b_kvq0, 384));
  SAMRS_TRY(e->alloc(&e->R_kvq0, size_t(4096) * 384));
  split_weight_kernel<<<(128 * 256 + 255) / 256, 256, 0, st>>>(e->dl[0].t2i.wk, 128, 256, 256.0f, e->wd_kvq0);
  split_weight_kernel<<<(128 * 256 + 255) / 256, 256, 0, st>>>(e->dl[0].t2i.wv, 128, 256, 256.0f, e->wd_kvq0 + size_t(128) * 768);
  split_weight_kernel<<<(128 * 256 + 255) / 256, 256, 0, st>>>(e->dl[0].i2t.wq, 128, 256, 256.0f, e->wd_kvq0 + size_t(256) * 768);
  SAMRS_CUDA_OK(cudaMemcpyAsync(e->b_kvq0, e->dl[0].t2i.bk, 512, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpyAsync(e->b_kvq0 + 128, e->dl[0].t2i.bv, 512, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpyAsync(e->b_kvq0 + 256, e->dl[0].i2t.bq, 512, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemsetAsync(e->R_kvq0, 0, size_t(4096) * 384 * 4, st));
  SAMRS_CUDA_OK(cudaMemcpy2DAsync(e->R_kvq0, 384 * 4, e->pek[0], 128 * 4, 128 * 4, 4096, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpy2DAsync(e->R_kvq0 + 256, 384 * 4, e->pek[1], 128 * 4, 128 * 4, 4096, cudaMemcpyDeviceToDevice, st));
  // tensor-core decoder GEMMs (3-term split fp16): concatenated along N
  //   P1 = keys(l0) -> [K(l1 t2i) | V(l1 t2i) | Q(l1 i2t)]      N = 384
  //   P2 = keys(l1) -> [K(final)  | V(final)  | ConvT1 (4x64)]  N = 512
  //   O0 / O1 = out_proj of the image->token attention of layer 0 / 1   N = 256, K = 128
  const float WS = 256.0f;
  auto split_into = [&](const float* w, int N, int K, __half* dst) {
    split_weight_kernel<<<(N * K + 255) / 256, 256, 0, st>>>(w, N, K, WS, dst);
  };
  SAMRS_TRY(e->alloc(&e->wd_p1, size_t(384) * 768));
  SAMRS_TRY(e->alloc(&e->wd_p2, size_t(512) * 768));
  SAMRS_TRY(e->alloc(&e->wd_o0, size_t(256) * 384));
  SAMRS_TRY(e->alloc(&e->wd_o1, size_t(256) * 384));
  SAMRS_TRY(e->alloc(&e->wd_up2, size_t(128) * 192));
  SAMRS_TRY(e->alloc(&e->up_b2r, 128));
  split_into(e->dl[1].t2i.wk, 128, 256, e->wd_p1);
  split_into(e->dl[1].t2i.wv, 128, 256, e->wd_p1 + size_t(128) * 768);
  split_into(e->dl[1].i2t.wq, 128, 256, e->wd_p1 + size_t(256) * 768);
  split_into(e->final_attn.wk, 128, 256, e->wd_p2);
  split_into(e->final_attn.wv, 128, 256, e->wd_p2 + size_t(128) * 768);
  split_into(e->up_w1r, 256, 256, e->wd_p2 + size_t(256) * 768);
  split_into(e->dl[0].i2t.wo, 256, 128, e->wd_o0);
  split_into(e->dl[1].i2t.wo, 256, 128, e->wd_o1);
  split_into(e->up_w2r, 128, 64, e->wd_up2);                 // ConvT2 as a GEMM: N = 4 sub-positions x 32 channels, K = 64
  tile_bias_kernel<<<1, 128, 0, st>>>(e->up_b2, e->up_b2r, 32, 4);
  SAMRS_TRY(e->alloc(&e->bias_p1, 384));
  SAMRS_TRY(e->alloc(&e->bias_p2, 512));
  SAMRS_CUDA_OK(cudaMemcpyAsync(e->bias_p1, e->dl[1].t2i.bk, 512, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpyAsync(e->bias_p1 + 128, e->dl[1].t2i.bv, 512, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpyAsync(e->bias_p1 + 256, e->dl[1].i2t.bq, 512, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpyAsync(e->bias_p2, e->final_attn.bk, 512, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpyAsync(e->bias_p2 + 128, e->final_attn.bv, 512, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpyAsync(e->bias_p2 + 256, e->up_b1r, 1024, cudaMemcpyDeviceToDevice, st));
  SAMRS_TRY(e->alloc(&e->R1, size_t(4096) * 384));
  SAMRS_TRY(e->alloc(&e->R2, size_t(4096) * 512));
  SAMRS_CUDA_OK(cudaMemsetAsync(e->R1, 0, size_t(4096) * 384 * 4, st));
  SAMRS_CUDA_OK(cudaMemsetAsync(e->R2, 0, size_t(4096) * 512 * 4, st));
  SAMRS_CUDA_OK(cudaMemcpy2DAsync(e->R1, 384 * 4, e->pek[2], 128 * 4, 128 * 4, 4096, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpy2DAsync(e->R1 + 256, 384 * 4, e->pek[3], 128 * 4, 128 * 4, 4096, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpy2DAsync(e->R2, 512 * 4, e->pek[4], 128 * 4, 128 * 4, 4096, cudaMemcpyDeviceToDevice, st));
  SAMRS_CUDA_OK(cudaGetLastError());
  return 0;
}

static int alloc_activations(Engine* e) {
  const size_t D = e->D, T = 4096;
  SAMRS_TRY(e->alloc(&e->a_pe, T * 768));
  SAMRS_TRY(e->alloc(&e->x, T * D));
  SAMRS_TRY(e->alloc(&e->xn, T * D));
  SAMRS_TRY(e->alloc(&e->qkv, T * 3 * D));
  SAMRS_TRY(e->alloc(&e->rel, size_t(e->heads) * T * 256));
  SAMRS_TRY(e->alloc(&e->attn_o, T * D));
  SAMRS_TRY(e->alloc(&e->hid, T * 4 * D));
  SAMRS_TRY(e->alloc(&e->x16, T * D));
  SAMRS_TRY(e->alloc(&e->neck0, T * 256));
  SAMRS_TRY(e->alloc(&e->neck_ln16, T * 256));
  SAMRS_TRY(e->alloc(&e->neck_col, T * 2304));
  SAMRS_TRY(e->alloc(&e->neck2, T * 256));
  SAMRS_TRY(e->alloc(&e->feat_tok, T * 256));
  SAMRS_TRY(e->alloc(&e->feat_nchw, T * 256));
  SAMRS_TRY(e->alloc(&e->src0, T * 256));
  SAMRS_TRY(e->alloc(&e->KVQ0, T * 384));
  SAMRS_TRY(e->alloc(&e->src0A, T * 512));
  e->ctx.splitk_ws_floats = size_t(8) * 1024 * 2048;
  SAMRS_TRY(e->alloc(&e->ctx.splitk_ws, e->ctx.splitk_ws_floats));
#ifdef SAMRS_EXPERIMENTS
  SAMRS_TRY(e->alloc(&e->ctx.sk_flags, size_t(2) * SK_MAX_TILES));
  SAMRS_CUDA_OK(cudaMemset(e->ctx.sk_flags, 0, size_t(2) * SK_MAX_TILES * sizeof(int)));
#endif
  return 0;
}

static int ensure_decoder_scratch(Engine* e, int B) {
  if (B <= e->dec_cap) return 0;
  // grow geometrically; the superseded buffers are given back first (cudaFree waits for work that still uses them)
  int cap = e->dec_cap ? e->dec_cap : 32;
  while (cap < B) cap *= 2;
  const size_t c = cap, TT = 16;
  e->release(&e->d_tok0); SAMRS_TRY(e->alloc(&e->d_tok0, c * TT * 256));
  e->release(&e->d_q); SAMRS_TRY(e->alloc(&e->d_q, c * TT * 256));
  e->release(&e->d_qkv); SAMRS_TRY(e->alloc(&e->d_qkv, c * TT * 768));
  e->release(&e->d_tmp256d); SAMRS_TRY(e->alloc(&e->d_tmp256d, c * TT * 256));
  e->release(&e->d_tmp128a); SAMRS_TRY(e->alloc(&e->d_tmp128a, c * TT * 128));
  e->release(&e->d_tmp128b); SAMRS_TRY(e->alloc(&e->d_tmp128b, c * TT * 128));
  e->release(&e->d_tmp128c); SAMRS_TRY(e->alloc(&e->d_tmp128c, c * TT * 128));
  e->release(&e->d_mlp); SAMRS_TRY(e->alloc(&e->d_mlp, c * TT * 2048));
  e->release(&e->d_t2i_part); SAMRS_TRY(e->alloc(&e->d_t2i_part, c * 8 * 16 * 8 * 18));
  e->release(&e->d_keys); SAMRS_TRY(e->alloc(&e->d_keys, c * 4096 * 256));
  e->release(&e->d_P); SAMRS_TRY(e->alloc(&e->d_P, c * 4096 * 512));
  e->release(&e->d_keysA); SAMRS_TRY(e->alloc(&e->d_keysA, c * 4096 * 512));
  e->release(&e->d_ioA); SAMRS_TRY(e->alloc(&e->d_ioA, c * 4096 * 256));
  e->release(&e->d_up1); SAMRS_TRY(e->alloc(&e->d_up1, c * 16384 * 128));
  e->release(&e->d_hyper); SAMRS_TRY(e->alloc(&e->d_hyper, c * 4 * 32));
  e->release(&e->d_hy_a); SAMRS_TRY(e->alloc(&e->d_hy_a, c * 256));
  e->release(&e->d_hy_b); SAMRS_TRY(e->alloc(&e->d_hy_b, c * 256));
  e->release(&e->d_iou_all); SAMRS_TRY(e->alloc(&e->d_iou_all, c * 4));
  e->release(&e->d_low); SAMRS_TRY(e->alloc(&e->d_low, c * 3 * 65536));
  e->dec_cap = cap;
  return 0;
}

static int ensure_mask_scratch(Engine* e, int B) {
  if (B <= e->mask_cap) return 0;
  int cap = e->mask_cap ? e->mask_cap : 8;
  while (cap < B) cap *= 2;
  const size_t c = cap;
  e->release(&e->d_Kp); SAMRS_TRY(e->alloc(&e->d_Kp, c * 4096 * 128));
  e->release(&e->d_Vp); SAMRS_TRY(e->alloc(&e->d_Vp, c * 4096 * 128));
  e->release(&e->d_Qp); SAMRS_TRY(e->alloc(&e->d_Qp, c * 4096 * 128));
  e->release(&e->d_src); SAMRS_TRY(e->alloc(&e->d_src, c * 4096 * 256));
  e->mask_cap = cap;
  return 0;
}

// ------------------------------------------------------------------ encoder
static int gemm_enc(Engine* e, cudaStream_t st, const __half* A, int lda, const __half* W, int M, int N, int K, void* out, int ldc,
                    bool out_half, const float* bias, const float* res, int ldr, int res_mod, int act, int accumulate = 0) {
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.out = out; p.ldc = ldc;
  p.bias = bias; p.res = res; p.ldr = ldr; p.res_mod = res_mod;
  p.tiles_m = p.tiles_n = 0;
  p.batch = 1; p.a_rank3 = 0; p.out_batch_stride = 0; p.out_scale = 0.f; p.dbg = nullptr; p.dbg_mode = 0; p.accumulate = accumulate;
  p.sk_flags = e->ctx.sk_flags;
  ProfScope ps(PC_GEMM, st);
  return launch_gemm_tc(A, lda, W, K, p, out_half, act, e->num_sms, st, 0);
}

// per-image decoder cache: src0 = features + no_mask_embed and its three layer-0 projections
static int gemm_dec(Engine* e, cudaStream_t st, const __half* A3, const __half* W3, int M, int N, int K3, float* out, int ldc,
                    const float* bias, const float* res, int ldr, int res_mod, int accumulate);

static int build_image_cache(Engine* e, cudaStream_t st) {
  add_rowvec_split_kernel<<<(4096 * 64 + 255) / 256, 256, 0, st>>>(e->feat_tok, e->no_mask, e->src0, e->src0A, 4096, 256);
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  // one tensor-core GEMM (3-term split fp16) for the three layer-0 projections of the image tokens: [K | V | Q] = src0 [Wk ; Wv ; Wq]^T
  SAMRS_TRY(gemm_dec(e, st, e->src0A, e->wd_kvq0, 4096, 384, 768, e->KVQ0, 384, e->b_kvq0, e->R_kvq0, 384, 4096, 0));
  e->image_set = true;
  return 0;
}

// everything of an encode that only touches engine-owned memory: a_pe (patch operand) -> feat_tok + the per-image decoder cache
static int encode_body(Engine* e, cudaStream_t st) {
  const int D = e->D, T = 4096;
  // patch embedding + absolute position embedding (image_encoder.py:107-109)
  SAMRS_TRY(gemm_enc(e, st, e->a_pe, 768, e->w_patch, T, D, 768, e->x, D, false, e->b_patch, e->pos_embed, D, 0, 0));
  for (int i = 0; i < e->depth; ++i) {
    const BlockWeights& b = e->blocks[i];
    { ProfScope ps(PC_LN, st); SAMRS_TRY((ln_rows<__half, 0>(st, e->x, D, b.ln1w, b.ln1b, 1e-6f, e->xn, D, T, D))); }
    SAMRS_TRY(gemm_enc(e, st, e->xn, D, b.wqkv, T, 3 * D, D, e->qkv, 3 * D, true, b.bqkv_eff, nullptr, 0, 0, 0));
    SAMRS_TRY(encoder_attention(e, st, e->qkv, b.reltab, b.global, e->attn_o));
    // proj / lin2 update the fp32 residual stream in place: x += A W^T + b through TMA reduce-add stores
    SAMRS_TRY(gemm_enc(e, st, e->attn_o, D, b.wproj, T, D, D, e->x, D, false, b.bproj_eff, nullptr, 0, 0, 0, 1));
    { ProfScope ps(PC_LN, st); SAMRS_TRY((ln_rows<__half, 0>(st, e->x, D, b.ln2w, b.ln2b, 1e-6f, e->xn, D, T, D))); }
    SAMRS_TRY(gemm_enc(e, st, e->xn, D, b.w1, T, 4 * D, D, e->hid, 4 * D, true, b.b1, nullptr, 0, 0, 1));
    SAMRS_TRY(gemm_enc(e, st, e->hid, 4 * D, b.w2, T, D, 4 * D, e->x, D, false, b.b2, nullptr, 0, 0, 0, 1));
  }
  // neck (image_encoder.py:88-104)
  cast_f32_f16_kernel<<<unsigned((size_t(T) * D / 4 + 255) / 256), 256, 0, st>>>(e->x, e->x16, size_t(T) * D / 4);
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  SAMRS_TRY(gemm_enc(e, st, e->x16, D, e->w_neck0, T, 256, D, e->neck0, 256, false, nullptr, nullptr, 0, 0, 0));
  SAMRS_TRY((ln_rows<__half, 0>(st, e->neck0, 256, e->neck1w, e->neck1b, 1e-6f, e->neck_ln16, 256, T, 256)));
  neck_im2col3x3_kernel<<<(4096 * 9 * 32 + 255) / 256, 256, 0, st>>>(e->neck_ln16, e->neck_col);
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  SAMRS_TRY(gemm_enc(e, st, e->neck_col, 2304, e->w_neck2, T, 256, 2304, e->neck2, 256, false, nullptr, nullptr, 0, 0, 0));
  SAMRS_TRY((ln_rows<float, 0>(st, e->neck2, 256, e->neck3w, e->neck3b, 1e-6f, e->feat_tok, 256, T, 256)));
  return build_image_cache(e, st);
}

static int encode_impl(Engine* e, const uint8_t* img, int H, int W, int chw, float* features_out, cudaStream_t st) {
  if (!e->weights_loaded) SAMRS_FAIL("encode: weights not loaded");
  if (H < 1 || W < 1 || H > 1024 || W > 1024) SAMRS_FAIL("encode: image must be at most 1024x1024 (resize the long side to 1024 first)");
  preprocess_im2col_kernel<<<(4096 * 48 + 255) / 256, 256, 0, st>>>(img, H, W, chw, e->a_pe);   // reads the caller's image
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  SAMRS_TRY(run_graphed(e, e->enc_graph, st, [&](cudaStream_t s) { return encode_body(e, s); }));
  e->image_set = true;
  float* nchw = features_out ? features_out : e->feat_nchw;                                      // writes the caller's buffer
  transpose_tok_to_nchw_kernel<<<dim3(128, 8), dim3(32, 8), 0, st>>>(e->feat_tok, nchw, 256);
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

// ------------------------------------------------------------------ decoder
// C = A' W'^T / 256 + bias + R : 3-term split-fp16 product on the tcgen05 GEMM (near-fp32 accuracy)
static int gemm_dec(Engine* e, cudaStream_t st, const __half* A3, const __half* W3, int M, int N, int K3, float* out, int ldc,
                    const float* bias, const float* res, int ldr, int res_mod, int accumulate = 0);
static int gemm_dec(Engine* e, cudaStream_t st, const __half* A3, const __half* W3, int M, int N, int K3, float* out, int ldc,
                    const float* bias, const float* res, int ldr, int res_mod, int accumulate) {
  GemmParams p;
  p.M = M; p.N = N; p.K = K3; p.out = out; p.ldc = ldc; p.bias = bias; p.res = res; p.ldr = ldr; p.res_mod = res_mod;
  p.tiles_m = p.tiles_n = 0; p.batch = 1; p.a_rank3 = 0; p.out_batch_stride = 0; p.out_scale = 1.0f / 256.0f; p.dbg = nullptr; p.dbg_mode = 0; p.accumulate = accumulate;
  // A = [hi | lo] with 2K = 2/3 K' columns; the GEMM reads the hi block a second time for its third term
  const int K2 = K3 / 3 * 2;
  if (K3 % 192 != 0) SAMRS_FAIL("split GEMM: K must be a multiple of 64");
  p.a_wrap_kb = K2 / GEMM_BK;
  return launch_gemm_tc(A3, K2, W3, K3, p, false, 0, e->num_sms, st, 0);
}

// LayerNorm of the token rows in place plus `with_pe = normed + pe` for the next attention's q / k input (transformer.py:161-179)
static int ln_tokens(cudaStream_t st, float* x, const float* g, const float* b, const float* pe, float* with_pe, int rows) {
  ln256_tok_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, g, b, 1e-5f, pe, with_pe, rows);
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

static int token_attention(Engine* e, cudaStream_t st, const DecAttn& a, const float* qk_in, const float* v_in,
                           int BT, int T, int B, float* out256 /*projected*/, const float* residual) {
  // self-attention on tokens (internal dim 256); the three projections land side by side as [BT][q | k | v].  They stay three
  // launches of 56 blocks: one GEMM over the concatenated weight rows (168 blocks x 66 KB of shared memory at once) made the
  // decoder 27 us faster in isolation and the two-stream step 1.5 % SLOWER - every SM it touches is closed to the other tile's
  // persistent 220 KB GEMM blocks for its duration (profiles/r02_decoder_ab.txt)
  SAMRS_TRY(sgemm(st, qk_in, 256, a.wq, 256, e->d_qkv, 768, a.bq, nullptr, 0, 0, BT, 256, 256, 0));
  SAMRS_TRY(sgemm(st, qk_in, 256, a.wk, 256, e->d_qkv + 256, 768, a.bk, nullptr, 0, 0, BT, 256, 256, 0));
  SAMRS_TRY(sgemm(st, v_in, 256, a.wv, 256, e->d_qkv + 512, 768, a.bv, nullptr, 0, 0, BT, 256, 256, 0));
  const size_t smem = (size_t(3) * T * 256 + 8 * T * T) * 4;
  tok_self_attn_kernel<<<B, 256, smem, st>>>(e->d_qkv, e->d_qkv + 256, e->d_qkv + 512, 768, e->d_tmp256d, T);
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  return sgemm(st, e->d_tmp256d, 256, a.wo, 256, out256, 256, a.bo, residual, 256, 0, BT, 256, 256, 0);
}

static int decode_body(Engine* e, cudaStream_t st, bool has_mask, int B, int T, int multimask);

static int decode_chunk(Engine* e, cudaStream_t st, const float* boxes, const float* points, const int* labels, int NP,
                        const float* mask_in, int B, int multimask, float* lowres_out, float* iou_out) {
  const int pad = (points && !boxes) ? 1 : 0;
  const int Ns = (points ? NP + pad : 0) + (boxes ? 2 : 0);
  const int T = 5 + Ns;
  if (T > 16) SAMRS_FAIL("decode: at most 11 sparse prompt tokens per prompt are supported");
  SAMRS_TRY(ensure_decoder_scratch(e, B));
  if (mask_in) SAMRS_TRY(ensure_mask_scratch(e, B));
  SAMRS_TRY(opt_in_smem(tok_self_attn_kernel, 64 * 1024));
  // (1) kernels that read the caller's prompt buffers: tokens, and the per-prompt dense embedding of mask prompts
  PromptParams pp;
  pp.gauss = e->gauss; pp.point_emb = e->point_emb; pp.not_a_point = e->not_a_point;
  pp.iou_token = e->iou_token; pp.mask_tokens = e->mask_tokens;
  pp.points = points; pp.labels = labels; pp.boxes = boxes; pp.NP = points ? NP : 0; pp.pad = pad; pp.T = T;
  pp.tokens = e->d_tok0;
  prompt_tokens_kernel<<<B, 128, 0, st>>>(pp);
  SAMRS_CUDA_OK(cudaGetLastError());
  count_launch();
  if (mask_in) {
    MaskEmbedParams mp{mask_in, e->md_w0, e->md_b0, e->md_g1, e->md_be1, e->md_w3, e->md_b3, e->md_g4, e->md_be4, e->md_w6, e->md_b6,
                       e->feat_tok, e->d_src};
    mask_embed_src_kernel<<<(B * 4096 + 127) / 128, 128, 0, st>>>(mp, B);
    SAMRS_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  // (2) the two-way transformer, hyper-network / IoU heads and the first upscaling stage: engine-owned memory only, one graph
  //     per (prompts, tokens per prompt, multimask, mask prompt)
  const uint64_t key = uint64_t(B) | (uint64_t(T) << 16) | (uint64_t(multimask ? 1 : 0) << 24) | (uint64_t(mask_in ? 1 : 0) << 25);
  SAMRS_TRY(run_graphed(e, e->dec_graphs[key], st, [&](cudaStream_t s) { return decode_body(e, s, mask_in != nullptr, B, T, multimask); }));
  // (3) kernels that write the caller's outputs: IoU predictions and the fused ConvT2 + GELU + hyper-network product
  const int NM = multimask ? 3 : 1;
  SAMRS_CUDA_OK(cudaMemcpyAsync(iou_out, e->d_iou_all, size_t(B) * NM * sizeof(float), cudaMemcpyDeviceToDevice, st));
  {
    // ConvT2 on the tensor cores (3-term split fp16, K' = 192) with GELU and the hyper-network product in the epilogue
    GemmParams p;
    p.M = B * 16384; p.N = 128; p.K = 192; p.out = e->d_P; p.ldc = 512; p.bias = e->up_b2r; p.res = nullptr; p.ldr = 0; p.res_mod = 0;
    p.tiles_m = p.tiles_n = 0; p.batch = 1; p.a_rank3 = 0; p.out_batch_stride = 0; p.out_scale = 1.0f / 256.0f; p.dbg = nullptr; p.dbg_mode = 0;
    p.accumulate = 0;
    p.hyper = e->d_hyper; p.low = lowres_out; p.hyper_nm = NM;
    p.a_wrap_kb = 2;                                   // d_up1 rows are [hi(64) | lo(64)]
    SAMRS_TRY(launch_gemm_tc(e->d_up1, 128, e->wd_up2, 192, p, false, 3, e->num_sms, st, 128));
  }
  return 0;
}

#ifndef SAMRS_DEC_FORK
#define SAMRS_DEC_FORK 3             // bit 0: layer 1's token chain beside layer 0's image-side tail; bit 1: up-scaling LayerNorm beside the heads
#endif
constexpr int DEC_FORK = SAMRS_DEC_FORK;
// work issued to `side_stream` after fork_side(i) runs beside what follows on `st` until join_side(i)
static int fork_side(Engine* e, cudaStream_t st, int i) {
  SAMRS_CUDA_OK(cudaEventRecord(e->ev_fork[i], st));
  SAMRS_CUDA_OK(cudaStreamWaitEvent(e->side_stream, e->ev_fork[i], 0));
  return 0;
}
static int join_side(Engine* e, cudaStream_t st, int i) {
  SAMRS_CUDA_OK(cudaEventRecord(e->ev_join[i], e->side_stream));
  SAMRS_CUDA_OK(cudaStreamWaitEvent(st, e->ev_join[i], 0));
  return 0;
}

constexpr int T2I_CHUNKS = 4;     // key chunks of the token -> image attention (blocks = prompts x 8 heads x chunks)

static int decode_body(Engine* e, cudaStream_t st, bool has_mask, int B, int T, int multimask) {
  const int BT = B * T;
  // image-side layer-0 operands: shared across prompts unless a mask prompt makes src per-prompt
  const float *K0 = e->KVQ0, *V0 = e->KVQ0 + 128, *Qi0 = e->KVQ0 + 256, *src = e->src0;
  int ld0 = 384;                       // row pitch of the layer-0 K / V / Q operands
  size_t kv_stride = 0;
  int src_mod = 4096;
  if (has_mask) {
    const DecLayer& L0 = e->dl[0];
    SAMRS_TRY(sgemm(st, e->d_src, 256, L0.t2i.wk, 256, e->d_Kp, 128, L0.t2i.bk, e->pek[0], 128, 4096, B * 4096, 128, 256, 0));
    SAMRS_TRY(sgemm(st, e->d_src, 256, L0.t2i.wv, 256, e->d_Vp, 128, L0.t2i.bv, nullptr, 0, 0, B * 4096, 128, 256, 0));
    SAMRS_TRY(sgemm(st, e->d_src, 256, L0.i2t.wq, 256, e->d_Qp, 128, L0.i2t.bq, e->pek[1], 128, 4096, B * 4096, 128, 256, 0));
    K0 = e->d_Kp; V0 = e->d_Vp; Qi0 = e->d_Qp; src = e->d_src;
    ld0 = 128;
    kv_stride = size_t(4096) * 128;
    src_mod = 0;
  }

  float* queries = e->d_q;
  float* qpl = e->d_tmp256d;          // queries + query_pe, written by the token LayerNorms (also the self-attention's output scratch)
  const int M4 = B * 4096;
  const unsigned ln_blocks = unsigned((size_t(M4) * 32 + 255) / 256);
  for (int layer = 0; layer < 2; ++layer) {
    const DecLayer& L = e->dl[layer];
    // (1) token self-attention (transformer.py:155-161) and (2)'s query projection (:164-165).  Layer 1's were issued to the side
    //     stream at the end of layer 0 (below): they only need the tokens, not the image-side update that was still running
    if (layer == 0) {
      SAMRS_TRY(token_attention(e, st, L.self_attn, e->d_tok0, e->d_tok0, BT, T, B, queries, nullptr));
      SAMRS_TRY(ln_tokens(st, queries, L.n1w, L.n1b, e->d_tok0, qpl, BT));
      SAMRS_TRY(sgemm(st, qpl, 256, L.t2i.wq, 256, e->d_tmp128c, 128, L.t2i.bq, nullptr, 0, 0, BT, 128, 256, 0));
      t2i_attn_kernel<<<dim3(B, 8, T2I_CHUNKS), 32 * T, 0, st>>>(e->d_tmp128c, K0, V0, ld0, kv_stride, e->d_t2i_part, T);
    } else {
      // K | V | Q(i2t) projections of the per-prompt image tokens in one tensor-core GEMM
      SAMRS_TRY(gemm_dec(e, st, e->d_keysA, e->wd_p1, M4, 384, 768, e->d_P, 384, e->bias_p1, e->R1, 384, 4096));
      if ((DEC_FORK & 1) != 0) SAMRS_TRY(join_side(e, st, 0));
      t2i_attn_kernel<<<dim3(B, 8, T2I_CHUNKS), 32 * T, 0, st>>>(e->d_tmp128c, e->d_P, e->d_P + 128, 384, size_t(4096) * 384, e->d_t2i_part, T);
    }
    t2i_combine_kernel<<<(B * 8 * T * 16 + 255) / 256, 256, 0, st>>>(e->d_t2i_part, e->d_tmp128b, B, T, T2I_CHUNKS);
    SAMRS_CUDA_OK(cudaGetLastError());
    count_launch(2);
    SAMRS_TRY(sgemm(st, e->d_tmp128b, 128, L.t2i.wo, 128, queries, 256, L.t2i.bo, queries, 256, 0, BT, 256, 128, 0));
    SAMRS_TRY((ln_rows<float, 0>(st, queries, 256, L.n2w, L.n2b, 1e-5f, queries, 256, BT, 256)));
    // (3) token MLP (transformer.py:171-173)
    SAMRS_TRY(sgemm(st, queries, 256, L.m1w, 256, e->d_mlp, 2048, L.m1b, nullptr, 0, 0, BT, 2048, 256, 1));
    SAMRS_TRY(sgemm(st, e->d_mlp, 2048, L.m2w, 2048, queries, 256, L.m2b, queries, 256, 0, BT, 256, 2048, 0));
    SAMRS_TRY(ln_tokens(st, queries, L.n3w, L.n3b, e->d_tok0, qpl, BT));
    // (4) image -> tokens (transformer.py:176-180)
    SAMRS_TRY(sgemm(st, qpl, 256, L.i2t.wk, 256, e->d_tmp128a, 128, L.i2t.bk, nullptr, 0, 0, BT, 128, 256, 0));
    SAMRS_TRY(sgemm(st, queries, 256, L.i2t.wv, 256, e->d_tmp128b, 128, L.i2t.bv, nullptr, 0, 0, BT, 128, 256, 0));
    if (layer == 0) {
      // the tokens are final for this layer: layer 1's self-attention, norm1 and t2i query projection (6 small, latency-bound
      // kernels) run on the side stream while the image-side kernels below stream their 100+ MB.  q = k = queries + pe is the
      // `qpl` norm3 just wrote (the attention kernel overwrites that scratch only after the q | k projection has read it);
      // the side chain touches queries, qpl, d_qkv and d_tmp128c, none of which the image side reads
      cudaStream_t sd = (DEC_FORK & 1) != 0 ? e->side_stream : st;
      const DecLayer& L1 = e->dl[1];
      if ((DEC_FORK & 1) != 0) SAMRS_TRY(fork_side(e, st, 0));
      SAMRS_TRY(token_attention(e, sd, L1.self_attn, qpl, queries, BT, T, B, queries, queries));
      SAMRS_TRY(ln_tokens(sd, queries, L1.n1w, L1.n1b, e->d_tok0, qpl, BT));
      SAMRS_TRY(sgemm(sd, qpl, 256, L1.t2i.wq, 256, e->d_tmp128c, 128, L1.t2i.bq, nullptr, 0, 0, BT, 128, 256, 0));
    }
    if (layer == 0)
      i2t_attn_kernel<<<dim3(4096 / I2T_TOK_PER_BLOCK, B), 256, size_t(2) * T * 8 * I2T_HP * 4, st>>>(Qi0, ld0, kv_stride, e->d_tmp128a, e->d_tmp128b, e->d_ioA, T);
    else
      i2t_attn_kernel<<<dim3(4096 / I2T_TOK_PER_BLOCK, B), 256, size_t(2) * T * 8 * I2T_HP * 4, st>>>(e->d_P + 256, 384, size_t(4096) * 384, e->d_tmp128a,
                                                                                  e->d_tmp128b, e->d_ioA, T);
    SAMRS_CUDA_OK(cudaGetLastError());
    count_launch();
    // keys = norm4(keys + out_proj(attn)); layer 0's `keys` is src (shared or per prompt)
    if (layer == 0)
      SAMRS_TRY(gemm_dec(e, st, e->d_ioA, e->wd_o0, M4, 256, 384, e->d_keys, 256, L.i2t.bo, src, 256, src_mod));
    else
      SAMRS_TRY(gemm_dec(e, st, e->d_ioA, e->wd_o1, M4, 256, 384, e->d_keys, 256, L.i2t.bo, nullptr, 0, 0, 1));
    ln256_split_kernel<<<ln_blocks, 256, 0, st>>>(e->d_keys, L.n4w, L.n4b, 1e-5f, layer == 0 ? e->d_keys : nullptr, e->d_keysA, M4);
    SAMRS_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  // final tokens -> image attention (transformer.py:99-104); its K | V projections share one GEMM with ConvT1
  {
    const DecAttn& a = e->final_attn;
    // q = queries + pe is the `qpl` layer 1's norm3 left behind
    SAMRS_TRY(sgemm(st, qpl, 256, a.wq, 256, e->d_tmp128c, 128, a.bq, nullptr, 0, 0, BT, 128, 256, 0));
    SAMRS_TRY(gemm_dec(e, st, e->d_keysA, e->wd_p2, M4, 512, 768, e->d_P, 512, e->bias_p2, e->R2, 512, 4096));
    t2i_attn_kernel<<<dim3(B, 8, T2I_CHUNKS), 32 * T, 0, st>>>(e->d_tmp128c, e->d_P, e->d_P + 128, 512, size_t(4096) * 512, e->d_t2i_part, T);
    t2i_combine_kernel<<<(B * 8 * T * 16 + 255) / 256, 256, 0, st>>>(e->d_t2i_part, e->d_tmp128b, B, T, T2I_CHUNKS);
    SAMRS_CUDA_OK(cudaGetLastError());
    count_launch(2);
    // up-scaling: the ConvT1 columns of P2 -> LN2d + GELU per 64-channel group -> split fp16 operand of the ConvT2 GEMM.  It only
    // needs P2: one full-GPU streaming kernel that runs on the side stream beside the ~10 small, latency-bound kernels of the
    // output projection, final norm and the hyper-network / IoU heads
    const bool fork_tail = (DEC_FORK & 2) != 0;
    cudaStream_t s64 = fork_tail ? e->side_stream : st;
    if (fork_tail) SAMRS_TRY(fork_side(e, st, 1));
    ln64_gelu_split_kernel<<<unsigned((size_t(M4) * 4 * 16 + 255) / 256), 256, 0, s64>>>(e->d_P, 512, 256, e->up_lnw, e->up_lnb, M4, e->d_up1);
    SAMRS_CUDA_OK(cudaGetLastError());
    count_launch();
    SAMRS_TRY(sgemm(st, e->d_tmp128b, 128, a.wo, 128, queries, 256, a.bo, queries, 256, 0, BT, 256, 128, 0));
    SAMRS_TRY((ln_rows<float, 0>(st, queries, 256, e->nfw, e->nfb, 1e-5f, queries, 256, BT, 256)));
  }
  // hypernetwork MLPs on the needed mask tokens; IoU head on the iou token (mask_decoder.py:150-172)
  const int NM = multimask ? 3 : 1, m_first = multimask ? 1 : 0;
  for (int j = 0; j < NM; ++j) {
    const Mlp3& h = e->hyper[m_first + j];
    // the mask token of every prompt is read in place: row stride T * 256 selects token 1 + m_first + j
    SAMRS_TRY(sgemm(st, queries + size_t(1 + m_first + j) * 256, T * 256, h.w[0], 256, e->d_hy_a, 256, h.b[0], nullptr, 0, 0, B, 256, 256, 1));
    SAMRS_TRY(sgemm(st, e->d_hy_a, 256, h.w[1], 256, e->d_hy_b, 256, h.b[1], nullptr, 0, 0, B, 256, 256, 1));
    SAMRS_TRY(sgemm(st, e->d_hy_b, 256, h.w[2], 256, e->d_hyper + j * 32, NM * 32, h.b[2], nullptr, 0, 0, B, 32, 256, 0));
  }
  SAMRS_TRY(sgemm(st, queries, T * 256, e->iou_head.w[0], 256, e->d_hy_a, 256, e->iou_head.b[0], nullptr, 0, 0, B, 256, 256, 1));   // iou token = row 0
  SAMRS_TRY(sgemm(st, e->d_hy_a, 256, e->iou_head.w[1], 256, e->d_hy_b, 256, e->iou_head.b[1], nullptr, 0, 0, B, 256, 256, 1));
  // last layer restricted to the returned slice: rows m_first .. m_first+NM-1 of the (4,256) weight
  SAMRS_TRY(sgemm(st, e->d_hy_b, 256, e->iou_head.w[2] + m_first * 256, 256, e->d_iou_all, NM, e->iou_head.b[2] + m_first, nullptr, 0, 0, B, NM,
                  256, 0));
  // the ConvT2 + GELU + hyper-network product (decode_chunk) needs both branches
  if ((DEC_FORK & 2) != 0) SAMRS_TRY(join_side(e, st, 1));
  return 0;
}

}  // namespace samrs

// ====================================================================== C ABI
using namespace samrs;
static void* g_gemm_dbg = nullptr;      // device buffer of 4096 u64 for gemm pipeline traces (tests/tools only)
extern "C" void samrs_test_set_gemm_trace(void* dev_buf) { g_gemm_dbg = dev_buf; }
static int g_gemm_mode = 0;             // GemmParams::dbg_mode for the pipeline experiments in tools/gemm_trace.py
extern "C" void samrs_test_set_gemm_mode(int mode) { g_gemm_mode = mode; }
extern "C" void samrs_test_set_attn_trace(void* dev_buf) { samrs::g_attn_dbg = dev_buf; }

struct LaunchScope {
  explicit LaunchScope(Engine* e) { t_ctx = e ? &e->ctx : nullptr; }
  ~LaunchScope() { t_ctx = nullptr; }
};

extern "C" {

int samrs_create(int device, int embed_dim, int depth, int num_heads, const int* global_idx, int n_global, void** out) {
  if (!out) return 1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return samrs::fail(__FILE__, __LINE__, "no CUDA device: samrs_b200 has no CPU path");
  if (device < 0 || device >= ndev) return samrs::fail(__FILE__, __LINE__, "bad device index");
  if (embed_dim % num_heads != 0) return samrs::fail(__FILE__, __LINE__, "embed_dim must be divisible by num_heads");
  const int hd = embed_dim / num_heads;
  if (hd != 64 && hd != 80) return samrs::fail(__FILE__, __LINE__, "head_dim must be 64 or 80");
  if (embed_dim % 32 != 0 || embed_dim > 1280) return samrs::fail(__FILE__, __LINE__, "embed_dim must be a multiple of 32, at most 1280");
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return samrs::fail(__FILE__, __LINE__, "cudaSetDevice failed");
  if (prop.major != 10) return samrs::fail(__FILE__, __LINE__, "samrs_b200 requires an sm_100 (Blackwell) GPU");
  Engine* e = new Engine();
  e->device = device;
  e->D = embed_dim; e->depth = depth; e->heads = num_heads; e->hd = hd;
  e->global_idx.assign(global_idx, global_idx + n_global);
  e->num_sms = prop.multiProcessorCount;
  bool side_ok = cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 && side_ok; ++i)
    side_ok = cudaEventCreateWithFlags(&e->ev_fork[i], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
  if (!side_ok || alloc_activations(e) != 0 || cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
    samrs_destroy(e);
    return 1;
  }
  *out = e;
  return 0;
}

int samrs_load_weights(void* engine, int n, const char* const* names, const void* const* dev_ptrs, const int64_t* numel, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  SrcMap m;
  for (int i = 0; i < n; ++i) m[names[i]] = Src{static_cast<const float*>(dev_ptrs[i]), numel[i]};
  LaunchScope ls(e);
  e->free_weights();                                   // loading again replaces the previous set instead of leaking it
  e->loading_weights = true;
  int rc = load_weights_impl(e, m, static_cast<cudaStream_t>(stream));
  e->loading_weights = false;
  if (rc == 0) e->weights_loaded = true;
  return set_err(e, rc);
}

int samrs_encode(void* engine, const uint8_t* img, int H, int W, int chw, float* features_out, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  ProfScope ps(PC_ENC_OTHER, static_cast<cudaStream_t>(stream));   // whole encode; "other" = this minus the categories
  return set_err(e, encode_impl(e, img, H, W, chw, features_out, static_cast<cudaStream_t>(stream)));
}

int samrs_set_features(void* engine, const float* features, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!e->weights_loaded) return set_err(e, samrs::fail(__FILE__, __LINE__, "set_features: weights not loaded"));
  transpose_nchw_to_tok_kernel<<<dim3(128, 8), dim3(32, 8), 0, st>>>(features, e->feat_tok, 256);
  count_launch();
  return set_err(e, build_image_cache(e, st));
}

int samrs_decode(void* engine, const float* boxes, const float* points, const int* labels, int NP, const float* mask_in, int B,
                 int multimask, float* lowres_out, float* iou_out, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  if (!e->image_set) return set_err(e, samrs::fail(__FILE__, __LINE__, "An image must be set with .set_image(...) before mask prediction."));
  if (B < 1) return set_err(e, samrs::fail(__FILE__, __LINE__, "decode: B must be >= 1"));
  if (points && !labels) return set_err(e, samrs::fail(__FILE__, __LINE__, "decode: point labels missing"));
  const int C = multimask ? 3 : 1;
  const int CH = 64;     // prompts per pass (bounds scratch at ~1.3 GB)
  ProfScope ps(PC_DECODER, static_cast<cudaStream_t>(stream));
  for (int b0 = 0; b0 < B; b0 += CH) {
    const int nb = (B - b0 < CH) ? (B - b0) : CH;
    int rc = decode_chunk(e, static_cast<cudaStream_t>(stream), boxes ? boxes + size_t(b0) * 4 : nullptr,
                          points ? points + size_t(b0) * NP * 2 : nullptr, labels ? labels + size_t(b0) * NP : nullptr, NP,
                          mask_in ? mask_in + size_t(b0) * 65536 : nullptr, nb, multimask, lowres_out + size_t(b0) * C * 65536,
                          iou_out + size_t(b0) * C);
    if (rc != 0) return set_err(e, rc);
  }
  return 0;
}

int samrs_postprocess(void* engine, const float* lowres, int NB, int in_h, int in_w, int out_h, int out_w, uint8_t* masks_out,
                      float* logits_out, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (NB < 1) return 0;
  ProfScope ps(PC_EPILOGUE, st);
  if (in_h == 1024 && in_w == 1024 && out_h == 1024 && out_w == 1024) {
    for (int b0 = 0; b0 < NB; b0 += 32768) {
      const int nb = NB - b0 < 32768 ? NB - b0 : 32768;
      upsample4_threshold_kernel<<<dim3(1, 257, nb), 256, 0, st>>>(lowres + size_t(b0) * 65536, masks_out ? masks_out + size_t(b0) * 1048576 : nullptr,
                                                                    logits_out ? logits_out + size_t(b0) * 1048576 : nullptr, nb);
    }
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_err(e, samrs::fail(__FILE__, __LINE__, "postprocess launch failed"));
    return 0;
  }
  // general sizes: 256 -> 1024 (fp32 scratch), crop, -> (out_h, out_w)
  const int CH = 16;
  if (!e->pp_full && e->alloc(&e->pp_full, size_t(CH) * 1048576) != 0) return set_err(e, 1);
  float* full = e->pp_full;
  for (int b0 = 0; b0 < NB; b0 += CH) {
    const int nb = NB - b0 < CH ? NB - b0 : CH;
    bilinear_resize_kernel<<<dim3(4, 1024, nb), 256, 0, st>>>(lowres + size_t(b0) * 65536, 256, 65536, 256, 256, full, nullptr, 1024, 1024, nb);
    bilinear_resize_kernel<<<dim3((out_w + 255) / 256, out_h, nb), 256, 0, st>>>(
        full, 1024, 1048576, in_h, in_w, logits_out ? logits_out + size_t(b0) * out_h * out_w : nullptr,
        masks_out ? masks_out + size_t(b0) * out_h * out_w : nullptr, out_h, out_w, nb);
    count_launch(2);
  }
  if (cudaGetLastError() != cudaSuccess) return set_err(e, samrs::fail(__FILE__, __LINE__, "postprocess launch failed"));
  return 0;
}

// Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter, in double as there
static void pil_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk, int& ksize) {
  const double scale = double(in_size) / double(out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  ksize = int(ceil(support)) * 2 + 1;
  bounds.assign(size_t(out_size) * 2, 0);
  kk.assign(size_t(out_size) * ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = int(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = int(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < ksize; ++x) w[x] = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double v = (x + xmin - center + 0.5) * ss;
      if (v < 0.0) v = -v;
      w[x] = v < 1.0 ? 1.0 - v : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) w[x] /= ww;
    for (int x = 0; x < ksize; ++x)
      kk[size_t(xx) * ksize + x] = w[x] < 0 ? int(-0.5 + w[x] * double(1 << 22)) : int(0.5 + w[x] * double(1 << 22));
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

static int resize_table(Engine* e, int in_size, int out_size, cudaStream_t st, const Engine::ResizeTab** out) {
  for (auto& t : e->resize_tabs)
    if (t.in == in_size && t.out == out_size) { t.used = ++e->resize_clock; *out = &t; return 0; }
  if (e->resize_tabs.size() >= 32) {
    // datasets with free image sizes (DOTA, HRSC) would grow the cache without bound: drop the least recently used
    // table (cudaFree waits for the kernels that may still read it; this only happens on the 33rd distinct size pair)
    size_t lru = 0;
    for (size_t i = 1; i < e->resize_tabs.size(); ++i)
      if (e->resize_tabs[i].used < e->resize_tabs[lru].used) lru = i;
    e->release(&e->resize_tabs[lru].bounds);
    e->release(&e->resize_tabs[lru].kk);
    e->resize_tabs.erase(e->resize_tabs.begin() + lru);
  }
  e->resize_tabs.emplace_back();
  Engine::ResizeTab& t = e->resize_tabs.back();
  t.in = in_size; t.out = out_size; t.ksize = 0; t.bounds = nullptr; t.kk = nullptr; t.used = ++e->resize_clock;
  pil_coeffs(in_size, out_size, t.h_bounds, t.h_kk, t.ksize);
  if (e->alloc(&t.bounds, t.h_bounds.size()) != 0 || e->alloc(&t.kk, t.h_kk.size()) != 0) return 1;
  // the host vectors stay alive inside the table entry, so the copies need no synchronisation
  SAMRS_CUDA_OK(cudaMemcpyAsync(t.bounds, t.h_bounds.data(), t.h_bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  SAMRS_CUDA_OK(cudaMemcpyAsync(t.kk, t.h_kk.data(), t.h_kk.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  *out = &t;
  return 0;
}

int samrs_resize_bilinear_u8(void* engine, const uint8_t* src_hwc, int H, int W, uint8_t* dst_hwc, int out_h, int out_w, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (H < 1 || W < 1 || out_h < 1 || out_w < 1 || H > 32768 || W > 32768 || out_h > 32768 || out_w > 32768 || !src_hwc || !dst_hwc)
    return set_err(e, samrs::fail(__FILE__, __LINE__, "resize: bad shape or null pointer"));
  const uint8_t* cur = src_hwc;
  if (out_w != W) {
    const Engine::ResizeTab* th = nullptr;
    if (resize_table(e, W, out_w, st, &th) != 0) return set_err(e, 1);
    uint8_t* hdst = dst_hwc;
    if (out_h != H) {                                   // a vertical pass follows: horizontal result goes to scratch
      const size_t need = size_t(H) * out_w * 3;
      if (need > e->resize_tmp_bytes) {
        e->release(&e->resize_tmp);
        if (e->alloc(&e->resize_tmp, need) != 0) return set_err(e, 1);
        e->resize_tmp_bytes = need;
      }
      hdst = e->resize_tmp;
    }
    pil_resize_h_kernel<<<dim3((out_w + 127) / 128, H), 128, 0, st>>>(cur, W, hdst, out_w, th->bounds, th->kk, th->ksize);
    count_launch();
    cur = hdst;
  }
  if (out_h != H) {
    const Engine::ResizeTab* tv = nullptr;
    if (resize_table(e, H, out_h, st, &tv) != 0) return set_err(e, 1);
    const int row = out_w * 3;
    pil_resize_v_kernel<<<dim3((row + 255) / 256, out_h), 256, 0, st>>>(cur, row, dst_hwc, tv->bounds, tv->kk, tv->ksize);
    count_launch();
  } else if (out_w == W) {
    SAMRS_CUDA_OK(cudaMemcpyAsync(dst_hwc, src_hwc, size_t(H) * W * 3, cudaMemcpyDeviceToDevice, st));
  }
  if (cudaGetLastError() != cudaSuccess) return set_err(e, samrs::fail(__FILE__, __LINE__, "resize launch failed"));
  return 0;
}

int samrs_rle_encode(void* engine, const uint8_t* masks, const float* lowres, int B, int H, int W, uint32_t* counts_out,
                     long long capacity, long long* offsets_out, long long* area_out, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B < 0 || H < 1 || W < 1) return set_err(e, samrs::fail(__FILE__, __LINE__, "rle_encode: bad shape"));
  if (B > 0 && (masks == nullptr) == (lowres == nullptr))
    return set_err(e, samrs::fail(__FILE__, __LINE__, "rle_encode: pass either bool masks or low-res logits"));
  if (lowres != nullptr && (H != 1024 || W != 1024))
    return set_err(e, samrs::fail(__FILE__, __LINE__, "rle_encode: the fused low-res path needs a 1024x1024 tile (pass masks otherwise)"));
  if (!offsets_out || !area_out || (!counts_out && capacity > 0))
    return set_err(e, samrs::fail(__FILE__, __LINE__, "rle_encode: null output"));
  if (size_t(H) * size_t(W) >= (size_t(1) << 31)) return set_err(e, samrs::fail(__FILE__, __LINE__, "rle_encode: mask too large"));
  ProfScope ps(PC_EPILOGUE, st);
  const int HW32 = (H + 31) / 32;
  const size_t words = size_t(B > 0 ? B : 1) * HW32 * W;
  if (words > e->rle_packed_words) {
    e->release(&e->rle_packed);
    if (e->alloc(&e->rle_packed, words) != 0) return set_err(e, 1);
    e->rle_packed_words = words;
  }
  if (B + 1 > e->rle_runs_cap) {
    e->release(&e->rle_runs);
    e->release(&e->rle_tstate);
    if (e->alloc(&e->rle_runs, size_t(B) + 1) != 0) return set_err(e, 1);
    if (e->alloc(&e->rle_tstate, (size_t(B) + 1) * RLE_THREADS * 2) != 0) return set_err(e, 1);
    e->rle_runs_cap = B + 1;
  }
  if (B > 0) {
    for (int b0 = 0; b0 < B; b0 += 32768) {
      const int nb = B - b0 < 32768 ? B - b0 : 32768;
      if (masks) rle_pack_kernel<<<dim3((W + 127) / 128, HW32, nb), 128, 0, st>>>(masks + size_t(b0) * H * W, H, W, HW32, e->rle_packed + size_t(b0) * HW32 * W);
      else rle_pack_lowres_kernel<<<dim3(8, 32, nb), 128, 0, st>>>(lowres + size_t(b0) * 65536, e->rle_packed + size_t(b0) * HW32 * W);
    }
    rle_scan_kernel<false><<<B, RLE_THREADS, 0, st>>>(e->rle_packed, H, W, HW32, e->rle_runs, area_out, nullptr, nullptr, 0, e->rle_tstate);
  }
  rle_offsets_kernel<<<1, 32, 0, st>>>(e->rle_runs, B, offsets_out);
  if (B > 0) rle_scan_kernel<true><<<B, RLE_THREADS, 0, st>>>(e->rle_packed, H, W, HW32, nullptr, nullptr, offsets_out, counts_out, capacity, e->rle_tstate);
  count_launch(B > 0 ? 4 : 1);
  if (cudaGetLastError() != cudaSuccess) return set_err(e, samrs::fail(__FILE__, __LINE__, "rle_encode launch failed"));
  return 0;
}

int samrs_rle_string(void* engine, const uint32_t* counts, const long long* offsets, int B, long long run_capacity,
                     uint8_t* chars_out, long long char_capacity, long long* char_offsets_out, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B < 0 || !offsets || !char_offsets_out || (B > 0 && !counts) || (!chars_out && char_capacity > 0))
    return set_err(e, samrs::fail(__FILE__, __LINE__, "rle_string: bad arguments"));
  ProfScope ps(PC_EPILOGUE, st);
  if (B + 1 > e->rle_runs_cap) {
    e->release(&e->rle_runs);
    e->release(&e->rle_tstate);
    if (e->alloc(&e->rle_runs, size_t(B) + 1) != 0) return set_err(e, 1);
    if (e->alloc(&e->rle_tstate, (size_t(B) + 1) * RLE_THREADS * 2) != 0) return set_err(e, 1);
    e->rle_runs_cap = B + 1;
  }
  if (B > 0) coco_string_kernel<false><<<B, RLE_THREADS, 0, st>>>(counts, offsets, run_capacity, e->rle_runs, nullptr, nullptr, 0);
  rle_offsets_kernel<<<1, 32, 0, st>>>(e->rle_runs, B, char_offsets_out);
  if (B > 0) coco_string_kernel<true><<<B, RLE_THREADS, 0, st>>>(counts, offsets, run_capacity, nullptr, char_offsets_out, chars_out, char_capacity);
  count_launch(B > 0 ? 3 : 1);
  if (cudaGetLastError() != cudaSuccess) return set_err(e, samrs::fail(__FILE__, __LINE__, "rle_string launch failed"));
  return 0;
}

int samrs_rbox_mask_prompts(void* engine, const float* polys, int B, int H, int W, float* out, int* status_out, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B < 0 || H < 1 || W < 1 || H > 16384 || W > 16384 || !out || !status_out || (B > 0 && !polys))
    return set_err(e, samrs::fail(__FILE__, __LINE__, "rbox_mask_prompts: bad shape or null pointer"));
  SAMRS_CUDA_OK(cudaMemsetAsync(status_out, 0, sizeof(int), st));
  if (B == 0) return 0;
  const size_t need = size_t(B) * H * W;
  if (need > e->rbox_mask_bytes) {
    e->release(&e->rbox_mask);
    if (e->alloc(&e->rbox_mask, need) != 0) return set_err(e, 1);
    e->rbox_mask_bytes = need;
  }
  SAMRS_CUDA_OK(cudaMemsetAsync(e->rbox_mask, 0, need, st));
  // ResizeLongestSide.get_preprocess_shape (SA/utils/transforms.py:94-102): the size the driver resizes the +-1000 mask to
  const double scale = 1024.0 / double(H > W ? H : W);
  const int nh = int(H * scale + 0.5), nw = int(W * scale + 0.5);
  rbox_fill_kernel<<<B, 256, 0, st>>>(polys, H, W, e->rbox_mask, status_out);
  rbox_prompt_kernel<<<dim3(1, 256, B), 256, 0, st>>>(e->rbox_mask, H, W, nh, nw, out);
  count_launch(2);
  if (cudaGetLastError() != cudaSuccess) return set_err(e, samrs::fail(__FILE__, __LINE__, "rbox_mask_prompts launch failed"));
  return 0;
}

int samrs_semantic_reduce(void* engine, const float* lowres, const int* class_ids, int B, uint8_t* label_map, int H, int W, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  if (H != 1024 || W != 1024) return set_err(e, samrs::fail(__FILE__, __LINE__, "semantic_reduce: only 1024x1024 tiles are supported"));
  if (B < 1) return 0;
  ProfScope ps(PC_EPILOGUE, static_cast<cudaStream_t>(stream));
  upsample4_paint_kernel<<<dim3(4, 257), 64, 0, static_cast<cudaStream_t>(stream)>>>(lowres, class_ids, B, label_map);
  count_launch();
  if (cudaGetLastError() != cudaSuccess) return set_err(e, samrs::fail(__FILE__, __LINE__, "semantic_reduce launch failed"));
  return 0;
}

int samrs_paint_masks(void* engine, const uint8_t* masks, const int* class_ids, int B, int H, int W, uint8_t* label_map, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  if (H < 1 || W < 1 || !masks || !class_ids || !label_map) return set_err(e, samrs::fail(__FILE__, __LINE__, "paint_masks: bad shape or null pointer"));
  if (B < 1) return 0;
  ProfScope ps(PC_EPILOGUE, static_cast<cudaStream_t>(stream));
  const size_t HW = size_t(H) * W;
  paint_masks_kernel<<<unsigned((HW / 4 + 256) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(masks, class_ids, B, HW, label_map);
  count_launch();
  if (cudaGetLastError() != cudaSuccess) return set_err(e, samrs::fail(__FILE__, __LINE__, "paint_masks launch failed"));
  return 0;
}

int samrs_profile(void* engine, int enable, float* ms_by_category, int* launches_by_category, int ncat) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  if (enable) {
    for (auto& r : e->ctx.prof.recs) { e->ctx.prof.pool.push_back(r.a); e->ctx.prof.pool.push_back(r.b); }
    e->ctx.prof.recs.clear();
    e->ctx.prof.on = true;
    return 0;
  }
  e->ctx.prof.on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) return set_err(e, samrs::fail(__FILE__, __LINE__, "profile: device sync failed"));
  for (int i = 0; i < ncat; ++i) { if (ms_by_category) ms_by_category[i] = 0.f; if (launches_by_category) launches_by_category[i] = 0; }
  for (auto& r : e->ctx.prof.recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (r.cat < ncat) { if (ms_by_category) ms_by_category[r.cat] += ms; if (launches_by_category) launches_by_category[r.cat] += 1; }
    e->ctx.prof.pool.push_back(r.a); e->ctx.prof.pool.push_back(r.b);
  }
  e->ctx.prof.recs.clear();
  return 0;
}

int samrs_set_pdl(void* engine, int enable) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  e->ctx.pdl = enable != 0;
  e->drop_graphs();                                  // captured launches carry the attribute
  return 0;
}

int samrs_set_graphs(void* engine, int enable) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  e->graphs_enabled = enable != 0;
  if (!enable) { cudaDeviceSynchronize(); e->drop_graphs(); }
  return 0;
}

int samrs_launch_count(void* engine, int64_t* out) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e || !out) return 1;
  *out = e->ctx.launches;
  return 0;
}

const char* samrs_last_error(void* engine) {
  Engine* e = static_cast<Engine*>(engine);
  if (e && !e->err.empty()) return e->err.c_str();
  return g_last_error.c_str();
}

void samrs_destroy(void* engine) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  e->drop_graphs();
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->side_stream) cudaStreamDestroy(e->side_stream);
  for (int i = 0; i < 2; ++i) {
    if (e->ev_fork[i]) cudaEventDestroy(e->ev_fork[i]);
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
  }
  for (void* p : e->allocs) cudaFree(p);
  for (void* p : e->weight_allocs) cudaFree(p);
  for (auto& r : e->ctx.prof.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (cudaEvent_t ev : e->ctx.prof.pool) cudaEventDestroy(ev);
  delete e;
}

int samrs_test_gemm(void* engine, const void* A, const void* B, int M, int N, int K, void* out, int out_half, const float* bias,
                    const float* res, int act_gelu, int force_bn, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.out = out; p.ldc = N; p.bias = bias; p.res = res; p.ldr = N; p.res_mod = 0; p.tiles_m = p.tiles_n = 0;
  p.batch = 1; p.a_rank3 = 0; p.out_batch_stride = 0; p.out_scale = 0.f;
  p.dbg = static_cast<unsigned long long*>(g_gemm_dbg);
  p.dbg_mode = g_gemm_mode;
  p.accumulate = 0;
  if (res != nullptr && res == out && !out_half) { p.res = nullptr; p.accumulate = 1; }   // in-place residual -> TMA reduce-add
  p.sk_flags = e->ctx.sk_flags;
  return set_err(e, launch_gemm_tc(static_cast<const __half*>(A), K, static_cast<const __half*>(B), K, p, out_half != 0, act_gelu, e->num_sms,
                                   static_cast<cudaStream_t>(stream), force_bn));
}

int samrs_test_attention(void* engine, const void* qkv, const float* rph, const float* rpw, int global_block, void* out, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  __half* tab = nullptr;
  int rc = build_reltab(e, static_cast<cudaStream_t>(stream), rph, rpw, global_block ? 64 : 14, &tab);
  if (rc == 0) rc = encoder_attention(e, static_cast<cudaStream_t>(stream), static_cast<const __half*>(qkv), tab, global_block != 0,
                                      static_cast<__half*>(out));
  return set_err(e, rc);
}

int samrs_test_sgemm(void* engine, const float* A, const float* W, float* C, const float* bias, int M, int N, int K, int act, void* stream) {
  Engine* e = static_cast<Engine*>(engine);
  if (!e) return 1;
  cudaSetDevice(e->device);
  LaunchScope ls(e);
  return set_err(e, sgemm(static_cast<cudaStream_t>(stream), A, K, W, K, C, N, bias, nullptr, 0, 0, M, N, K, act));
}

}  // extern "C"
