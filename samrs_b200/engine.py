"""ctypes binding of `libsamrs_b200.so` (C ABI in include/samrs_b200.h).

PyTorch is used here only for device memory, streams and dtype plumbing: every
tensor handed to the library is passed as a raw `data_ptr()` plus sizes.  There
is no CPU path: if the library is missing, or no sm_100 GPU is present,
constructing an `Engine` raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence, Tuple

import torch

from .config import SamGeometry, geometry

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.environ.get("SAMRS_LIB", "libsamrs_b200.so"))   # SAMRS_LIB: A/B builds (tools)
_lib: Optional[ctypes.CDLL] = None

# name -> (restype, argtypes); must list every symbol declared in include/samrs_b200.h
_vp, _i, _i64p = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)
ABI = {
    "samrs_create": (_i, [_i, _i, _i, _i, ctypes.POINTER(_i), _i, ctypes.POINTER(_vp)]),
    "samrs_load_weights": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_vp), _i64p, _vp]),
    "samrs_encode": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "samrs_set_features": (_i, [_vp, _vp, _vp]),
    "samrs_decode": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "samrs_postprocess": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "samrs_semantic_reduce": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    "samrs_rbox_mask_prompts": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "samrs_paint_masks": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "samrs_resize_bilinear_u8": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp]),
    "samrs_rle_encode": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, ctypes.c_longlong, _vp, _vp, _vp]),
    "samrs_rle_string": (_i, [_vp, _vp, _vp, _i, ctypes.c_longlong, _vp, ctypes.c_longlong, _vp, _vp]),
    "samrs_profile": (_i, [_vp, _i, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_i), _i]),
    "samrs_set_graphs": (_i, [_vp, _i]),
    "samrs_set_pdl": (_i, [_vp, _i]),
    "samrs_launch_count": (_i, [_vp, _i64p]),
    "samrs_last_error": (ctypes.c_char_p, [_vp]),
    "samrs_destroy": (None, [_vp]),
    "samrs_test_gemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "samrs_test_attention": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "samrs_test_sgemm": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "samrs_test_set_gemm_trace": (None, [_vp]),
    "samrs_test_set_gemm_mode": (None, [_i]),
    "samrs_test_set_attn_trace": (None, [_vp]),
}


def load_library() -> ctypes.CDLL:
    """dlopen the in-tree library and bind every ABI symbol; raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(samrs_b200 has no CPU or PyTorch fallback)")
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Engine:
    """One engine per (GPU, stream).  All methods enqueue on torch's current stream of `device`."""

    def __init__(self, variant_or_geometry, device="cuda"):
        g = variant_or_geometry if isinstance(variant_or_geometry, SamGeometry) else geometry(variant_or_geometry)
        self.geometry = g
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("samrs_b200 runs on CUDA devices only (no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        gi = (ctypes.c_int * len(g.global_attn_indexes))(*g.global_attn_indexes)
        rc = self._lib.samrs_create(self.device.index, g.embed_dim, g.depth, g.num_heads, gi,
                                    len(g.global_attn_indexes), ctypes.byref(self._h))
        if rc != 0:
            raise RuntimeError("samrs_create failed: " + self._lib.samrs_last_error(None).decode())
        self.weights_loaded = False
        # the image embedding lives in the engine, not in whoever called encode(): every change of it bumps this counter so
        # that a predictor can tell whether the engine still holds ITS image (two predictors may share one Sam)
        self.feature_gen = 0

    # -- helpers -----------------------------------------------------------------
    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed: {self._lib.samrs_last_error(self._h).decode()}")

    def _dev(self, t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        if t.device != self.device:
            raise ValueError(f"tensor on {t.device}, engine on {self.device}")
        return t.to(dtype).contiguous()

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.samrs_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- API ---------------------------------------------------------------------
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """Strict load of a reference-layout `Sam.state_dict()` (any device; staged to the GPU as fp32)."""
        from .weights import check_state_dict
        check_state_dict(self.geometry, state_dict)
        names = list(state_dict.keys())
        staged = [state_dict[k].detach().to(device=self.device, dtype=torch.float32).contiguous() for k in names]
        n = len(names)
        c_names = (ctypes.c_char_p * n)(*[k.encode() for k in names])
        c_ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in staged])
        c_numel = (ctypes.c_int64 * n)(*[t.numel() for t in staged])
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_load_weights(self._h, n, c_names, c_ptrs, c_numel, _stream(self.device)), "load_weights")
            torch.cuda.current_stream(self.device).synchronize()   # staged tensors are freed on return
        self.weights_loaded = True
        self.feature_gen += 1                                      # a reload invalidates the cached image embedding

    def encode(self, image_u8: torch.Tensor, chw: bool = False) -> torch.Tensor:
        """uint8 image on the device, HWC (or CHW) with H,W <= 1024 -> features (1,256,64,64) fp32."""
        img = self._dev(image_u8, torch.uint8)
        H, W = (img.shape[1], img.shape[2]) if chw else (img.shape[0], img.shape[1])
        feats = torch.empty((1, 256, 64, 64), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_encode(self._h, img.data_ptr(), H, W, int(chw), feats.data_ptr(), _stream(self.device)), "encode")
        self.feature_gen += 1
        return feats

    def set_features(self, features: torch.Tensor) -> None:
        f = self._dev(features, torch.float32)
        assert tuple(f.shape) == (1, 256, 64, 64)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_set_features(self._h, f.data_ptr(), _stream(self.device)), "set_features")
        self.feature_gen += 1

    def decode(self, boxes: Optional[torch.Tensor] = None, point_coords: Optional[torch.Tensor] = None,
               point_labels: Optional[torch.Tensor] = None, mask_input: Optional[torch.Tensor] = None,
               multimask_output: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (low_res (B,C,256,256) fp32, iou (B,C) fp32)."""
        b = self._dev(boxes, torch.float32) if boxes is not None else None
        pc = self._dev(point_coords, torch.float32) if point_coords is not None else None
        pl = self._dev(point_labels, torch.int32) if point_labels is not None else None
        mi = self._dev(mask_input, torch.float32) if mask_input is not None else None
        if pc is not None:
            B, NP = pc.shape[0], pc.shape[1]
        elif b is not None:
            B, NP = b.shape[0], 0
        elif mi is not None:
            B, NP = mi.shape[0], 0
        else:
            raise ValueError("decode needs at least one prompt kind")
        C = 3 if multimask_output else 1
        low = torch.empty((B, C, 256, 256), dtype=torch.float32, device=self.device)
        iou = torch.empty((B, C), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_decode(self._h, _ptr(b), _ptr(pc), _ptr(pl), NP, _ptr(mi), B, int(multimask_output),
                                               low.data_ptr(), iou.data_ptr(), _stream(self.device)), "decode")
        return low, iou

    def postprocess(self, low_res: torch.Tensor, input_size: Sequence[int], original_size: Sequence[int],
                    return_logits: bool = False) -> torch.Tensor:
        """(B,C,256,256) -> (B,C,H,W) bool masks (or fp32 logits)."""
        low = self._dev(low_res, torch.float32)
        B, C = low.shape[0], low.shape[1]
        oh, ow = int(original_size[0]), int(original_size[1])
        if return_logits:
            out = torch.empty((B, C, oh, ow), dtype=torch.float32, device=self.device)
            args = (None, out.data_ptr())
        else:
            out = torch.empty((B, C, oh, ow), dtype=torch.bool, device=self.device)
            args = (out.data_ptr(), None)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_postprocess(self._h, low.data_ptr(), B * C, int(input_size[0]), int(input_size[1]), oh, ow,
                                                    args[0], args[1], _stream(self.device)), "postprocess")
        return out

    def semantic_reduce(self, low_res: torch.Tensor, class_ids: torch.Tensor, label_map: torch.Tensor) -> torch.Tensor:
        """In-place painter reduce of (B,[1,]256,256) logits into a (1024,1024) uint8 label map (255-initialised by the caller)."""
        low = self._dev(low_res, torch.float32)
        ids = self._dev(class_ids, torch.int32)
        assert label_map.dtype == torch.uint8 and label_map.is_contiguous() and label_map.device == self.device
        B = low.shape[0]
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_semantic_reduce(self._h, low.data_ptr(), ids.data_ptr(), B, label_map.data_ptr(),
                                                        label_map.shape[0], label_map.shape[1], _stream(self.device)), "semantic_reduce")
        return label_map

    def rbox_mask_prompts(self, polys: torch.Tensor, image_hw: Sequence[int], check: bool = True) -> torch.Tensor:
        """(B,4,2) rotated-box polygons in original-image pixels -> (B,1,256,256) float32 mask prompts, the `mask_input` the
        rbox driver builds per box with OpenCV on the host (main_sam_rbox_mask_instance.py:125-141,159-164).
        `check` synchronises once to verify that every vertex was inside the image (outside vertices are not supported)."""
        p = self._dev(polys, torch.float32).reshape(-1, 4, 2)
        B, (H, W) = p.shape[0], (int(image_hw[0]), int(image_hw[1]))
        out = torch.empty((B, 1, 256, 256), dtype=torch.float32, device=self.device)
        status = torch.empty((1,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_rbox_mask_prompts(self._h, p.data_ptr(), B, H, W, out.data_ptr(), status.data_ptr(),
                                                          _stream(self.device)), "rbox_mask_prompts")
        if check and B > 0 and int(status.item()) != 0:
            raise ValueError("rbox_mask_prompts: a polygon vertex lies outside the image (after truncation to integers)")
        return out

    def paint_masks(self, masks: torch.Tensor, class_ids: torch.Tensor, label_map: torch.Tensor) -> torch.Tensor:
        """In-place painter reduce of (B,[1,]H,W) bool masks into an (H,W) uint8 label map: the general-size companion of
        `semantic_reduce` (tiles whose original size is not 1024 x 1024)."""
        m = masks.view(torch.uint8) if masks.dtype == torch.bool else masks
        m = self._dev(m, torch.uint8)
        ids = self._dev(class_ids, torch.int32)
        H, W = int(label_map.shape[0]), int(label_map.shape[1])
        assert label_map.dtype == torch.uint8 and label_map.is_contiguous() and label_map.device == self.device
        assert m.shape[-2:] == (H, W)
        B = m.numel() // (H * W)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_paint_masks(self._h, m.data_ptr(), ids.data_ptr(), B, H, W, label_map.data_ptr(),
                                                    _stream(self.device)), "paint_masks")
        return label_map

    def resize_image(self, image: torch.Tensor, out_hw: Sequence[int]) -> torch.Tensor:
        """(H,W,3) uint8 CUDA image -> (out_h,out_w,3) uint8, bit-identical to `PIL.Image.resize(..., BILINEAR)`
        (the reference's `ResizeLongestSide.apply_image`, utils/transforms.py:26-31)."""
        img = self._dev(image, torch.uint8)
        if img.dim() != 3 or img.shape[2] != 3:
            raise ValueError("resize_image expects an HWC uint8 image with 3 channels")
        oh, ow = int(out_hw[0]), int(out_hw[1])
        out = torch.empty((oh, ow, 3), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_resize_bilinear_u8(self._h, img.data_ptr(), img.shape[0], img.shape[1], out.data_ptr(), oh, ow,
                                                           _stream(self.device)), "resize_image")
        return out

    def rle_encode(self, masks: Optional[torch.Tensor] = None, low_res: Optional[torch.Tensor] = None, capacity: Optional[int] = None):
        """Uncompressed COCO RLE + area of B masks on the device (the driver's `maskUtils.encode` / `np.sum`,
        main_sam_hbox_semantic.py:200-203, at the run level pinned by amg.py:107-135).

        Pass `masks` (B,[1,]H,W) bool / uint8, or `low_res` (B,[1,]256,256) logits of a 1024x1024 tile (fused upsample +
        threshold, no mask tensor).  Returns `(counts int32[capacity], offsets int64[B+1], area int64[B])` as CUDA
        tensors, asynchronously; mask b's runs are `counts[offsets[b]:offsets[b+1]]`.  If `offsets[B] > capacity` the
        tail was not written: call again with a larger capacity (`samrs_b200.rle.to_rle_dicts` checks this)."""
        if (masks is None) == (low_res is None):
            raise ValueError("rle_encode: pass either masks or low_res")
        if masks is not None:
            m = masks
            if m.dtype == torch.bool:
                m = m.view(torch.uint8)
            m = self._dev(m, torch.uint8)
            if m.dim() == 4:
                m = m.reshape(m.shape[0] * m.shape[1], m.shape[2], m.shape[3])
            B, H, W = m.shape
            src_m, src_l = m.data_ptr(), None
        else:
            low = self._dev(low_res, torch.float32)
            low = low.reshape(-1, 256, 256)
            B, H, W = low.shape[0], 1024, 1024
            src_m, src_l = None, low.data_ptr()
        if capacity is None:
            capacity = max(1, B) * 16384
        counts = torch.empty((max(1, capacity),), dtype=torch.int32, device=self.device)
        offsets = torch.empty((B + 1,), dtype=torch.int64, device=self.device)
        area = torch.empty((max(1, B),), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_rle_encode(self._h, src_m, src_l, B, H, W, counts.data_ptr(), int(capacity),
                                                   offsets.data_ptr(), area.data_ptr(), _stream(self.device)), "rle_encode")
        return counts, offsets, area[:B]

    def rle_strings(self, counts: torch.Tensor, offsets: torch.Tensor, char_capacity: Optional[int] = None):
        """Compressed COCO strings of the runs returned by `rle_encode`, on the device: `(chars uint8[char_capacity],
        char_offsets int64[B+1])`; mask b's string is `bytes(chars[char_offsets[b]:char_offsets[b+1]])`.  A run needs at
        most 7 characters; the default capacity is 2 per run + 64 per mask (real masks: ~1.3), check `char_offsets[-1]`."""
        B = offsets.numel() - 1
        cnt = counts if counts.dtype == torch.int32 else counts.to(torch.int32)
        if char_capacity is None:
            char_capacity = 2 * cnt.numel() + 64 * max(B, 1)
        chars = torch.empty((max(1, char_capacity),), dtype=torch.uint8, device=self.device)
        coff = torch.empty((B + 1,), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_rle_string(self._h, cnt.data_ptr(), offsets.data_ptr(), B, cnt.numel(), chars.data_ptr(),
                                                   int(char_capacity), coff.data_ptr(), _stream(self.device)), "rle_string")
        return chars, coff

    PROFILE_CATEGORIES = ("gemm_tc", "attn_window", "attn_global", "relpos", "layernorm", "encode_total", "decode_total", "epilogue")

    def profile_begin(self) -> None:
        self._check(self._lib.samrs_profile(self._h, 1, None, None, 0), "profile")

    def profile_end(self):
        """-> {category: (milliseconds, scopes)} accumulated since profile_begin (synchronises the device)."""
        n = len(self.PROFILE_CATEGORIES)
        ms, cnt = (ctypes.c_float * n)(), (ctypes.c_int * n)()
        self._check(self._lib.samrs_profile(self._h, 0, ms, cnt, n), "profile")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROFILE_CATEGORIES)}

    def set_graphs(self, enable: bool) -> None:
        """CUDA-graph replay of the encode / decode bodies (default on); off = every kernel is launched directly."""
        self._check(self._lib.samrs_set_graphs(self._h, int(bool(enable))), "set_graphs")

    def set_pdl(self, enable: bool) -> None:
        """Programmatic dependent launch of the GEMM / attention / LayerNorm kernels (default on)."""
        self._check(self._lib.samrs_set_pdl(self._h, int(bool(enable))), "set_pdl")

    def launch_count(self) -> int:
        c = ctypes.c_int64(0)
        self._lib.samrs_launch_count(self._h, ctypes.byref(c))
        return int(c.value)

    # -- kernel-level test hooks ---------------------------------------------------
    def test_gemm(self, A: torch.Tensor, B: torch.Tensor, out_half: bool, bias=None, res=None, gelu=False, force_bn=0, out=None):
        """`out` given and `res is out` exercises the in-place residual path (TMA reduce-add)."""
        M, K = A.shape
        N = B.shape[0]
        if out is None:
            out = torch.empty((M, N), dtype=torch.float16 if out_half else torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_test_gemm(self._h, A.data_ptr(), B.data_ptr(), M, N, K, out.data_ptr(), int(out_half),
                                                  _ptr(bias), _ptr(res), int(gelu), force_bn, _stream(self.device)), "test_gemm")
        return out

    def test_attention(self, qkv: torch.Tensor, rel_pos_h: torch.Tensor, rel_pos_w: torch.Tensor, global_block: bool):
        out = torch.empty((4096, self.geometry.embed_dim), dtype=torch.float16, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_test_attention(self._h, qkv.data_ptr(), rel_pos_h.data_ptr(), rel_pos_w.data_ptr(),
                                                       int(global_block), out.data_ptr(), _stream(self.device)), "test_attention")
        return out

    def test_sgemm(self, A: torch.Tensor, W: torch.Tensor, bias=None, act=0):
        M, K = A.shape
        N = W.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.samrs_test_sgemm(self._h, A.data_ptr(), W.data_ptr(), out.data_ptr(), _ptr(bias), M, N, K, act,
                                                   _stream(self.device)), "test_sgemm")
        return out
