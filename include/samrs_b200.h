/* samrs_b200 -- C ABI of the Blackwell-native SAM box-prompted mask engine.
 *
 * The reference has no FFI on this path: its boundary is the Python class surface of the vendored
 * segment_anything package.  These entry points are what a maintainer binds underneath
 * that surface (ctypes stub in INTEGRATION.md); each cites the reference call it replaces, paths
 * relative to "/root/reference/Generate Dataset/segment_anything/".
 *
 * Conventions: every function returns 0 on success and non-zero on failure (never throws);
 * samrs_last_error() gives the message.  One engine per (GPU, stream); not thread-safe per handle.
 * The caller owns all buffers; pointers are DEVICE pointers unless stated; all work is enqueued
 * asynchronously on `stream` (a cudaStream_t passed as void*), nothing synchronises the device.
 */
#ifndef SAMRS_B200_H
#define SAMRS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* replaces _build_sam(encoder_embed_dim, encoder_depth, encoder_num_heads, encoder_global_attn_indexes)
 * -- build_sam.py:55-101.  Geometry outside these four numbers is fixed exactly as there. */
int samrs_create(int device, int embed_dim, int depth, int num_heads, const int* global_attn_indexes,
                 int n_global, void** engine_out);

/* replaces sam.load_state_dict(state_dict) -- build_sam.py:102-106 (strict: every key of
 * Sam.state_dict() must be present with the right element count).  names[i] are state-dict keys,
 * dev_ptrs[i] fp32 device tensors in the reference layout; the engine repacks into its own storage,
 * so the caller may free them after the stream has drained. */
int samrs_load_weights(void* engine, int n, const char* const* names, const void* const* dev_ptrs,
                       const int64_t* numel, void* stream);

/* replaces Sam.preprocess + ImageEncoderViT.forward as called by SamPredictor.set_torch_image
 * -- predictor.py:88-89, modeling/sam.py:164-174, modeling/image_encoder.py:106-116.
 * img: uint8, HWC (chw=0) or CHW (chw=1), H,W <= 1024 (already resized so the long side is 1024).
 * features_out: (1,256,64,64) fp32 NCHW, may be NULL.  The embedding is also cached in the engine. */
int samrs_encode(void* engine, const uint8_t* img, int H, int W, int chw, float* features_out, void* stream);

/* installs a caller-supplied image embedding (the documented "assign predictor.features by hand" use);
 * features: (1,256,64,64) fp32 NCHW. */
int samrs_set_features(void* engine, const float* features, void* stream);

/* replaces PromptEncoder.forward + MaskDecoder.forward as called by SamPredictor.predict_torch
 * -- predictor.py:222-235, modeling/prompt_encoder.py:128-173, modeling/mask_decoder.py:71-174.
 * boxes (B,4) xyxy | points (B,NP,2) + labels (B,NP) int32 | mask_in (B,1,256,256), each may be NULL,
 * all in the 1024 input frame.  lowres_out (B,C,256,256), iou_out (B,C); C = 3 if multimask else 1.
 * Ordered on `stream` like every other call; internally part of the work runs on an engine-owned stream that forks from and
 * joins back into `stream` through events (parallel branches of the replayed graph). */
int samrs_decode(void* engine, const float* boxes, const float* points, const int* labels, int NP,
                 const float* mask_in, int B, int multimask, float* lowres_out, float* iou_out, void* stream);

/* replaces Sam.postprocess_masks + threshold -- modeling/sam.py:133-162, predictor.py:240-243.
 * lowres (NB,256,256) -> bilinear to 1024, crop to (in_h,in_w), bilinear to (out_h,out_w).
 * masks_out (NB,out_h,out_w) uint8 0/1 (== torch.bool storage) and/or logits_out fp32; either may be NULL. */
int samrs_postprocess(void* engine, const float* lowres, int NB, int in_h, int in_w, int out_h, int out_w,
                      uint8_t* masks_out, float* logits_out, void* stream);

/* fuses postprocess + threshold + the driver's painter reduce
 * -- Generate Dataset/main_sam_hbox_semantic.py:162,195-199: label_map (H,W) uint8 is updated in place;
 * every pixel takes class_ids[j] of the highest j whose mask is true, else keeps its value (initialise to 255).
 * Currently H = W = 1024. */
int samrs_semantic_reduce(void* engine, const float* lowres, const int* class_ids, int B,
                          uint8_t* label_map_inout, int H, int W, void* stream);

/* rotated boxes -> mask prompts on the device: replaces the per-box OpenCV loop of
 * main_sam_rbox_mask_instance.py:125-141 (fillPoly of the int-truncated polygon, +-1000, resize to long side 1024, pad with
 * -1000, resize to 256 x 256).  polys (B,4,2) float in original-image pixels -> out (B,256,256) float = the `mask_input` of
 * predict_torch.  fillPoly's raster is reproduced bit for bit, the two INTER_LINEAR passes to within one float32 ulp
 * (see csrc/rbox.cuh).  *status_out (device int) becomes 1 if a vertex lies outside the image (not supported). */
int samrs_rbox_mask_prompts(void* engine, const float* polys, int B, int H, int W, float* out, int* status_out, void* stream);

/* the same painter reduce over bool masks of any size (tiles whose original size is not 1024 x 1024: the masks come from
 * samrs_postprocess): main_sam_hbox_semantic.py:195-199; highest box index wins, composes across chunks. */
int samrs_paint_masks(void* engine, const uint8_t* masks /*[B][H][W]*/, const int* class_ids, int B, int H, int W,
                      uint8_t* label_map_inout, void* stream);

/* image resize of the predictor on the device: replaces ResizeLongestSide.apply_image
 * -- segment_anything/utils/transforms.py:26-31 (torchvision resize of a PIL image == PIL.Image.resize(BILINEAR)).
 * src (H,W,3) uint8 -> dst (out_h,out_w,3) uint8, both on the device, bit-identical to Pillow's two-pass 8-bit
 * fixed-point resampler (horizontal pass first, 22-bit taps, triangle filter widened when shrinking). */
int samrs_resize_bilinear_u8(void* engine, const uint8_t* src_hwc, int H, int W, uint8_t* dst_hwc, int out_h, int out_w, void* stream);

/* instance payload on the device: uncompressed COCO run-length encoding and area of B masks.
 * Replaces, per mask, the D2H of the bool mask + `maskUtils.encode(np.asfortranarray(mask))` + `np.sum(mask)`
 * -- Generate Dataset/main_sam_hbox_semantic.py:200-203; run semantics as pinned in-repo by
 * segment_anything/utils/amg.py:107-135 (mask_to_rle_pytorch): column-major pixel order, alternating run lengths
 * starting with a run of zeros (a mask whose first pixel is set gets a leading 0), lengths sum to H*W.
 * Exactly one source is non-NULL: `masks` (B,H,W) uint8 0/non-0 (torch.bool storage, e.g. predict_torch's masks), or
 * `lowres` (B,256,256) logits of a 1024x1024 tile (H = W = 1024), thresholded after the same bilinear x4 upsample as
 * samrs_postprocess without materialising the masks.
 * counts_out[offsets_out[b] .. offsets_out[b+1]) are mask b's runs; offsets_out has B+1 entries, area_out B.
 * Runs beyond `capacity` entries are not written: the caller checks offsets_out[B] <= capacity and retries. */
int samrs_rle_encode(void* engine, const uint8_t* masks, const float* lowres, int B, int H, int W,
                     uint32_t* counts_out, long long capacity, long long* offsets_out, long long* area_out, void* stream);

/* compressed COCO string of runs produced by samrs_rle_encode (the `counts` string maskUtils.encode returns,
 * main_sam_hbox_semantic.py:200-201; scheme of pycocotools' rleToString, which is neither vendored nor installed: parity
 * unpinned, host restatement in samrs_b200/rle.py).  Mask b's characters are chars_out[char_offsets_out[b] .. [b+1]);
 * if char_offsets_out[B] > char_capacity the tail was not written. */
int samrs_rle_string(void* engine, const uint32_t* counts, const long long* offsets /*[B+1]*/, int B, long long run_capacity,
                     uint8_t* chars_out, long long char_capacity, long long* char_offsets_out /*[B+1]*/, void* stream);

/* device-side timing by kernel category (bench.py's roofline): enable=1 starts recording CUDA events around the
 * engine's launches on their stream; enable=0 synchronises and returns milliseconds / scope counts per category:
 * 0 tcgen05 GEMM, 1 windowed attention, 2 global attention, 3 rel-pos terms, 4 LayerNorm, 5 whole encode,
 * 6 decode call, 7 post-processing / painter. */
int samrs_profile(void* engine, int enable, float* ms_by_category, int* launches_by_category, int ncat);

/* CUDA graphs: by default the engine captures the encode body and each decode shape on their second call and replays the
 * graphs afterwards (launches that touch caller-owned buffers stay outside; profiling runs eagerly).  enable = 0 drops
 * the graphs and launches every kernel directly. */
int samrs_set_graphs(void* engine, int enable);

/* programmatic dependent launch (default on): the tcgen05 GEMM, attention and LayerNorm kernels are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and wait for their predecessor themselves (griddepcontrol.wait) after
 * their own prologue, so barrier / tensor-memory set-up overlaps the previous kernel's tail.  enable = 0 restores plain
 * stream order for A/B measurements. */
int samrs_set_pdl(void* engine, int enable);

/* kernels launched by this engine since creation (bench.py's gpu_launches). */
int samrs_launch_count(void* engine, int64_t* count_out);

const char* samrs_last_error(void* engine);   /* engine may be NULL: last error of a failed samrs_create */
void samrs_destroy(void* engine);

/* ---- kernel-level test hooks (used by tests/ only; same kernels the engine launches) ---- */
/* C[M,N] = act(A[M,K] B[N,K]^T + bias + res) through the tcgen05 GEMM; A,B fp16; out fp16 or fp32. */
int samrs_test_gemm(void* engine, const void* A, const void* B, int M, int N, int K, void* out, int out_half,
                    const float* bias, const float* res, int act_gelu, int force_bn, void* stream);
/* encoder attention of one block on a (4096, 3*D) fp16 qkv activation; out (4096, D) fp16. */
int samrs_test_attention(void* engine, const void* qkv, const float* rel_pos_h, const float* rel_pos_w,
                         int global_block, void* out, void* stream);
/* fp32 CUDA-core GEMM of the decoder: C = act(A W^T + bias), act 0 none / 1 relu / 2 gelu. */
int samrs_test_sgemm(void* engine, const float* A, const float* W, float* C, const float* bias, int M, int N,
                     int K, int act, void* stream);
/* pipeline instrumentation for tools/gemm_trace.py and tools/attn_trace.py (process-wide, not thread-safe):
 * a device buffer of 4096 uint64 that CTA 0 of the next tensor-core GEMM / attention launches fills with clock64
 * stamps (NULL switches tracing off), and the GEMM experiment mask (bit0 skip TMA loads, bit1 skip MMAs, bit2 skip
 * the epilogue; results are garbage, timing only). */
void samrs_test_set_gemm_trace(void* dev_buf);
void samrs_test_set_gemm_mode(int mode);
void samrs_test_set_attn_trace(void* dev_buf);

#ifdef __cplusplus
}
#endif
#endif
