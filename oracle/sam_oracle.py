"""CPU oracle for the SAM box-prompted mask path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product path
(`samrs_b200.engine`, the drop-in `segment_anything`) never does.

This is a functional fp32 restatement, on torch CPU tensors and a flat
`state_dict`, of the arithmetic the reference performs through its nn.Modules.
Each function cites the reference lines it follows (paths relative to
`/root/reference/Generate Dataset/segment_anything/`, "SA/").  The arithmetic
itself lives in third-party PyTorch/ATen (reference pin torch 1.9.0, installed
2.11.0); the restatement therefore calls the same ATen primitives
(`F.linear`, `F.layer_norm`, `softmax`, `F.interpolate`, ...) on CPU.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned by running the reference's own
`segment_anything` in the build container on the same seeded synthetic
checkpoint and committing its outputs as fixtures: see `oracle/make_golden.py`
and `tests/golden/`.  `tests/test_oracle_golden.py` checks this file against
those fixtures on every CPU run.

`round_gemm_inputs` emulates the engine's mixed-precision recipe (fp16 MMA
operands, fp32 accumulate) and exists only for precision-budget analysis in
DESIGN.md; parity tests always use the default fp32 path.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from samrs_b200.config import SamGeometry, geometry

Tensor = torch.Tensor
W = Dict[str, Tensor]

PIXEL_MEAN = (123.675, 116.28, 103.53)   # SA/build_sam.py:99
PIXEL_STD = (58.395, 57.12, 57.375)      # SA/build_sam.py:100


class _Rounder:
    """Optional operand rounding used to emulate tensor-core input precision."""

    def __init__(self, mode: Optional[str]):
        self.mode = mode

    def __call__(self, t: Tensor) -> Tensor:
        if self.mode is None:
            return t
        if self.mode == "fp16":
            return t.half().float()
        if self.mode == "bf16":
            return t.bfloat16().float()
        if self.mode == "tf32":
            i = t.contiguous().view(torch.int32)
            i = ((i + 0x1000) & ~0x1FFF)
            return i.view(torch.float32)
        raise ValueError(self.mode)


# --------------------------------------------------------------------------
# pre-processing                                            SA/modeling/sam.py
# --------------------------------------------------------------------------
def preprocess(img_chw_u8: Tensor, img_size: int = 1024) -> Tensor:
    """`Sam.preprocess` (SA/modeling/sam.py:164-174): normalise then zero-pad."""
    mean = torch.tensor(PIXEL_MEAN).view(3, 1, 1)
    std = torch.tensor(PIXEL_STD).view(3, 1, 1)
    x = (img_chw_u8 - mean) / std          # u8 -> f32 promotion, as in the reference
    h, w = x.shape[-2:]
    return F.pad(x, (0, img_size - w, 0, img_size - h))


# --------------------------------------------------------------------------
# image encoder                                   SA/modeling/image_encoder.py
# --------------------------------------------------------------------------
def _rel_table(rel_pos: Tensor, size: int) -> Tensor:
    """`get_rel_pos` (image_encoder.py:292-322) for q_size == k_size == size.

    The table length is always 2*size-1 on this path, so the interpolation
    branch (:305-313) is dead; R[i, j] = rel_pos[i - j + size - 1].
    """
    assert rel_pos.shape[0] == 2 * size - 1
    idx = torch.arange(size)[:, None] - torch.arange(size)[None, :] + (size - 1)
    return rel_pos[idx]                     # (size, size, hd)


def _attention(x: Tensor, w: W, pre: str, heads: int, rnd: _Rounder) -> Tensor:
    """`Attention.forward` (image_encoder.py:224-240) on (B, S, S, D) tokens."""
    B, S, _, D = x.shape
    hd = D // heads
    qkv = F.linear(rnd(x), rnd(w[pre + "qkv.weight"]), w[pre + "qkv.bias"])
    qkv = qkv.reshape(B, S * S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, B * heads, S * S, hd).unbind(0)
    q, k, v = rnd(q), rnd(k), rnd(v)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)                      # :231
    # add_decomposed_rel_pos (:325-361) -- note the UNSCALED q (:234)
    Rh = _rel_table(w[pre + "rel_pos_h"], S)
    Rw = _rel_table(w[pre + "rel_pos_w"], S)
    rq = q.reshape(B * heads, S, S, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    attn = (attn.view(B * heads, S, S, S, S) + rel_h[..., :, None] + rel_w[..., None, :]).view(
        B * heads, S * S, S * S)
    attn = attn.softmax(dim=-1)                                        # :236
    o = (rnd(attn) @ v).view(B, heads, S, S, hd).permute(0, 2, 3, 1, 4).reshape(B, S, S, D)
    return F.linear(rnd(o), rnd(w[pre + "proj.weight"]), w[pre + "proj.bias"])


def _block(x: Tensor, w: W, pre: str, g: SamGeometry, windowed: bool, rnd: _Rounder) -> Tensor:
    """`Block.forward` (image_encoder.py:166-182)."""
    D, G, ws = g.embed_dim, g.grid, g.window
    y = F.layer_norm(x, (D,), w[pre + "norm1.weight"], w[pre + "norm1.bias"], 1e-6)
    if windowed:
        # window_partition (:243-264): zero-pad AFTER the norm, then regroup
        pad = (ws - G % ws) % ws
        Gp = G + pad
        y = F.pad(y, (0, 0, 0, pad, 0, pad))
        n = Gp // ws
        y = y.view(1, n, ws, n, ws, D).permute(0, 1, 3, 2, 4, 5).reshape(n * n, ws, ws, D)
        y = _attention(y, w, pre + "attn.", g.num_heads, rnd)
        # window_unpartition (:267-289): inverse regroup, crop the padding
        y = y.view(1, n, n, ws, ws, D).permute(0, 1, 3, 2, 4, 5).reshape(1, Gp, Gp, D)
        y = y[:, :G, :G, :]
    else:
        y = _attention(y, w, pre + "attn.", g.num_heads, rnd)
    x = x + y
    z = F.layer_norm(x, (D,), w[pre + "norm2.weight"], w[pre + "norm2.bias"], 1e-6)
    z = F.linear(rnd(z), rnd(w[pre + "mlp.lin1.weight"]), w[pre + "mlp.lin1.bias"])
    z = F.gelu(z)                                                      # nn.GELU default = erf
    z = F.linear(rnd(z), rnd(w[pre + "mlp.lin2.weight"]), w[pre + "mlp.lin2.bias"])
    return x + z


def _layernorm2d(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-6) -> Tensor:
    """`LayerNorm2d.forward` (SA/modeling/common.py:38-43): over channels, biased var."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return weight[:, None, None] * x + bias[:, None, None]


def encode_image(w: W, g: SamGeometry, x: Tensor, round_gemm_inputs: Optional[str] = None,
                 tap: Optional[Callable[[str, Tensor], None]] = None) -> Tensor:
    """`ImageEncoderViT.forward` (image_encoder.py:106-116). x: (1,3,1024,1024) f32."""
    rnd = _Rounder(round_gemm_inputs)
    e = "image_encoder."
    t = F.conv2d(rnd(x), rnd(w[e + "patch_embed.proj.weight"]), w[e + "patch_embed.proj.bias"],
                 stride=g.patch)                                        # :391-395
    t = t.permute(0, 2, 3, 1) + w[e + "pos_embed"]                     # :107-109
    if tap:
        tap("patch_embed", t)
    for i in range(g.depth):
        t = _block(t, w, f"{e}blocks.{i}.", g, i not in g.global_attn_indexes, rnd)
        if tap:
            tap(f"block{i}", t)
    t = t.permute(0, 3, 1, 2)
    t = F.conv2d(rnd(t), rnd(w[e + "neck.0.weight"]))                  # :88-104
    t = _layernorm2d(t, w[e + "neck.1.weight"], w[e + "neck.1.bias"])
    t = F.conv2d(rnd(t), rnd(w[e + "neck.2.weight"]), padding=1)
    t = _layernorm2d(t, w[e + "neck.3.weight"], w[e + "neck.3.bias"])
    return t                                                            # (1,256,64,64)


# --------------------------------------------------------------------------
# prompt encoder                                 SA/modeling/prompt_encoder.py
# --------------------------------------------------------------------------
def _pe_encoding(w: W, coords01: Tensor) -> Tensor:
    """`PositionEmbeddingRandom._pe_encoding` (prompt_encoder.py:190-197)."""
    c = 2 * coords01 - 1
    c = c @ w["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    c = 2 * np.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(w: W, g: SamGeometry) -> Tensor:
    """`PromptEncoder.get_dense_pe` (prompt_encoder.py:62-71,199-210) -> (1,256,64,64)."""
    G = g.grid
    ones = torch.ones((G, G), dtype=torch.float32)
    y = (ones.cumsum(dim=0) - 0.5) / G
    x = (ones.cumsum(dim=1) - 0.5) / G
    return _pe_encoding(w, torch.stack([x, y], dim=-1)).permute(2, 0, 1).unsqueeze(0)


def _pe_coords(w: W, g: SamGeometry, coords: Tensor) -> Tensor:
    """`forward_with_coords` (prompt_encoder.py:212-219)."""
    c = coords.clone()
    c[:, :, 0] = c[:, :, 0] / g.img_size
    c[:, :, 1] = c[:, :, 1] / g.img_size
    return _pe_encoding(w, c.to(torch.float))


def embed_prompts(w: W, g: SamGeometry, points: Optional[Tuple[Tensor, Tensor]],
                  boxes: Optional[Tensor], masks: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """`PromptEncoder.forward` (prompt_encoder.py:128-173) -> sparse (B,Ns,256), dense (B,256,64,64)."""
    p = "prompt_encoder."
    C, G = g.out_chans, g.grid
    if points is not None:
        bs = points[0].shape[0]
    elif boxes is not None:
        bs = boxes.shape[0]
    elif masks is not None:
        bs = masks.shape[0]
    else:
        bs = 1
    sparse = torch.empty((bs, 0, C))
    if points is not None:                                              # _embed_points :73-91
        coords, labels = points
        coords = coords + 0.5
        if boxes is None:
            coords = torch.cat([coords, torch.zeros((bs, 1, 2))], dim=1)
            labels = torch.cat([labels, -torch.ones((bs, 1), dtype=labels.dtype)], dim=1)
        pe = _pe_coords(w, g, coords)
        pe[labels == -1] = 0.0
        pe[labels == -1] += w[p + "not_a_point_embed.weight"]
        pe[labels == 0] += w[p + "point_embeddings.0.weight"]
        pe[labels == 1] += w[p + "point_embeddings.1.weight"]
        sparse = torch.cat([sparse, pe], dim=1)
    if boxes is not None:                                               # _embed_boxes :93-100
        b = (boxes + 0.5).reshape(-1, 2, 2)
        ce = _pe_coords(w, g, b)
        ce[:, 0, :] += w[p + "point_embeddings.2.weight"]
        ce[:, 1, :] += w[p + "point_embeddings.3.weight"]
        sparse = torch.cat([sparse, ce], dim=1)
    if masks is not None:                                               # _embed_masks :102-105, stack :51-59
        d = F.conv2d(masks, w[p + "mask_downscaling.0.weight"], w[p + "mask_downscaling.0.bias"], stride=2)
        d = F.gelu(_layernorm2d(d, w[p + "mask_downscaling.1.weight"], w[p + "mask_downscaling.1.bias"]))
        d = F.conv2d(d, w[p + "mask_downscaling.3.weight"], w[p + "mask_downscaling.3.bias"], stride=2)
        d = F.gelu(_layernorm2d(d, w[p + "mask_downscaling.4.weight"], w[p + "mask_downscaling.4.bias"]))
        dense = F.conv2d(d, w[p + "mask_downscaling.6.weight"], w[p + "mask_downscaling.6.bias"])
    else:
        dense = w[p + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(bs, -1, G, G)
    return sparse, dense


# --------------------------------------------------------------------------
# mask decoder               SA/modeling/mask_decoder.py, SA/modeling/transformer.py
# --------------------------------------------------------------------------
def _dec_attention(w: W, pre: str, q: Tensor, k: Tensor, v: Tensor, heads: int = 8) -> Tensor:
    """decoder `Attention.forward` (transformer.py:218-240)."""
    q = F.linear(q, w[pre + "q_proj.weight"], w[pre + "q_proj.bias"])
    k = F.linear(k, w[pre + "k_proj.weight"], w[pre + "k_proj.bias"])
    v = F.linear(v, w[pre + "v_proj.weight"], w[pre + "v_proj.bias"])

    def split(t: Tensor) -> Tensor:
        b, n, c = t.shape
        return t.reshape(b, n, heads, c // heads).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    a = q @ k.permute(0, 1, 3, 2)
    a = a / math.sqrt(q.shape[-1])
    a = torch.softmax(a, dim=-1)
    o = a @ v
    b, h, n, c = o.shape
    o = o.transpose(1, 2).reshape(b, n, h * c)
    return F.linear(o, w[pre + "out_proj.weight"], w[pre + "out_proj.bias"])


def _ln(w: W, pre: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w[pre + ".weight"], w[pre + ".bias"], 1e-5)


def _two_way(w: W, src: Tensor, pos: Tensor, tokens: Tensor) -> Tuple[Tensor, Tensor]:
    """`TwoWayTransformer.forward` (transformer.py:62-106) + blocks (:151-182)."""
    t = "mask_decoder.transformer."
    keys = src.flatten(2).permute(0, 2, 1)
    kpe = pos.flatten(2).permute(0, 2, 1)
    queries, qpe = tokens, tokens
    for i in range(2):
        L = f"{t}layers.{i}."
        if i == 0:                                                      # skip_first_layer_pe :155-156
            queries = _dec_attention(w, L + "self_attn.", queries, queries, queries)
        else:
            q = queries + qpe
            queries = queries + _dec_attention(w, L + "self_attn.", q, q, queries)
        queries = _ln(w, L + "norm1", queries)
        q, k = queries + qpe, keys + kpe
        queries = queries + _dec_attention(w, L + "cross_attn_token_to_image.", q, k, keys)
        queries = _ln(w, L + "norm2", queries)
        m = F.linear(queries, w[L + "mlp.lin1.weight"], w[L + "mlp.lin1.bias"])
        m = F.linear(F.relu(m), w[L + "mlp.lin2.weight"], w[L + "mlp.lin2.bias"])
        queries = _ln(w, L + "norm3", queries + m)
        q, k = queries + qpe, keys + kpe
        keys = keys + _dec_attention(w, L + "cross_attn_image_to_token.", k, q, queries)
        keys = _ln(w, L + "norm4", keys)
    q, k = queries + qpe, keys + kpe
    queries = queries + _dec_attention(w, t + "final_attn_token_to_image.", q, k, keys)
    queries = _ln(w, t + "norm_final_attn", queries)
    return queries, keys


def _mlp3(w: W, pre: str, x: Tensor) -> Tensor:
    """3-layer ReLU `MLP` (mask_decoder.py:179-201)."""
    for j in range(3):
        x = F.linear(x, w[f"{pre}.layers.{j}.weight"], w[f"{pre}.layers.{j}.bias"])
        if j < 2:
            x = F.relu(x)
    return x


def decode_masks(w: W, g: SamGeometry, features: Tensor, sparse: Tensor, dense: Tensor,
                 multimask_output: bool) -> Tuple[Tensor, Tensor]:
    """`MaskDecoder.forward` / `predict_masks` (mask_decoder.py:71-174)."""
    m = "mask_decoder."
    B = sparse.shape[0]
    out_tokens = torch.cat([w[m + "iou_token.weight"], w[m + "mask_tokens.weight"]], dim=0)
    tokens = torch.cat((out_tokens.unsqueeze(0).expand(B, -1, -1), sparse), dim=1)
    src = torch.repeat_interleave(features, B, dim=0) + dense
    pos = torch.repeat_interleave(dense_pe(w, g), B, dim=0)
    b, c, h, wd = src.shape
    hs, src = _two_way(w, src, pos, tokens)
    iou_tok, mask_toks = hs[:, 0, :], hs[:, 1:5, :]
    src = src.transpose(1, 2).view(b, c, h, wd)
    up = F.conv_transpose2d(src, w[m + "output_upscaling.0.weight"], w[m + "output_upscaling.0.bias"], stride=2)
    up = F.gelu(_layernorm2d(up, w[m + "output_upscaling.1.weight"], w[m + "output_upscaling.1.bias"]))
    up = F.gelu(F.conv_transpose2d(up, w[m + "output_upscaling.3.weight"], w[m + "output_upscaling.3.bias"], stride=2))
    hyper = torch.stack([_mlp3(w, f"{m}output_hypernetworks_mlps.{i}", mask_toks[:, i, :]) for i in range(4)], dim=1)
    b, c, h, wd = up.shape
    masks = (hyper @ up.view(b, c, h * wd)).view(b, -1, h, wd)
    iou = _mlp3(w, m + "iou_prediction_head", iou_tok)
    sl = slice(1, None) if multimask_output else slice(0, 1)            # :101-107
    return masks[:, sl], iou[:, sl]


# --------------------------------------------------------------------------
# post-processing and the driver-side reduce
# --------------------------------------------------------------------------
def postprocess_masks(low_res: Tensor, input_size: Tuple[int, int], original_size: Tuple[int, int],
                      img_size: int = 1024) -> Tensor:
    """`Sam.postprocess_masks` (SA/modeling/sam.py:133-162)."""
    m = F.interpolate(low_res, (img_size, img_size), mode="bilinear", align_corners=False)
    m = m[..., : input_size[0], : input_size[1]]
    return F.interpolate(m, original_size, mode="bilinear", align_corners=False)


def predict_torch(w: W, g: SamGeometry, features: Tensor, point_coords: Optional[Tensor],
                  point_labels: Optional[Tensor], boxes: Optional[Tensor] = None,
                  mask_input: Optional[Tensor] = None, multimask_output: bool = True,
                  return_logits: bool = False, input_size=(1024, 1024), original_size=(1024, 1024)):
    """`SamPredictor.predict_torch` (SA/predictor.py:168-245)."""
    pts = (point_coords, point_labels) if point_coords is not None else None
    sparse, dense = embed_prompts(w, g, pts, boxes, mask_input)
    low, iou = decode_masks(w, g, features, sparse, dense, multimask_output)
    masks = postprocess_masks(low, input_size, original_size, g.img_size)
    if not return_logits:
        masks = masks > 0.0                                             # Sam.mask_threshold, sam.py:19
    return masks, iou, low


def apply_image(image_hwc_u8: np.ndarray, target: int = 1024) -> np.ndarray:
    """`ResizeLongestSide.apply_image` (SA/utils/transforms.py:26-31): torchvision's `resize(to_pil_image(img), size)`
    is PIL's bilinear resampler on the uint8 image (Pillow = the un-vendored dependency; restated in pil_resize_oracle.py)."""
    from PIL import Image
    h, w = image_hwc_u8.shape[:2]
    scale = target * 1.0 / max(h, w)
    nh, nw = int(h * scale + 0.5), int(w * scale + 0.5)
    if (nh, nw) == (h, w):
        return image_hwc_u8
    return np.array(Image.fromarray(image_hwc_u8).resize((nw, nh), Image.BILINEAR))


def set_image(w: W, g: SamGeometry, image_hwc_u8: np.ndarray, round_gemm_inputs: Optional[str] = None) -> Tensor:
    """`SamPredictor.set_image` (SA/predictor.py:34-90): resize the long side to 1024 (identity for 1024x1024 tiles),
    normalise, zero-pad the short side AFTER normalisation (`Sam.preprocess`, SA/modeling/sam.py:164-174), encode.
    The size `predict_torch` needs as `input_size` is `apply_image(img).shape[:2]`."""
    img = apply_image(image_hwc_u8, g.img_size)
    x = torch.as_tensor(np.ascontiguousarray(img)).permute(2, 0, 1).contiguous()
    return encode_image(w, g, preprocess(x, g.img_size)[None], round_gemm_inputs)


def apply_boxes(boxes: Tensor, original_size: Tuple[int, int], target: int = 1024) -> Tensor:
    """`ResizeLongestSide.apply_boxes_torch` (SA/utils/transforms.py:83-102)."""
    oh, ow = original_size
    scale = target * 1.0 / max(oh, ow)
    nh, nw = int(oh * scale + 0.5), int(ow * scale + 0.5)
    c = boxes.reshape(-1, 2, 2).clone().to(torch.float)
    c[..., 0] = c[..., 0] * (nw / ow)
    c[..., 1] = c[..., 1] * (nh / oh)
    return c.reshape(-1, 4)


def painter_reduce(masks_bool: np.ndarray, labels, canvas: Optional[np.ndarray] = None) -> np.ndarray:
    """The driver's semantic reduce (`Generate Dataset/main_sam_hbox_semantic.py:162,195-199`):
    canvas starts at 255; boxes are visited in order and each overwrites its
    true pixels with its class id -- the last box wins on overlap."""
    B, H, Wd = masks_bool.shape
    seg = np.full((H, Wd), 255, dtype=np.uint8) if canvas is None else canvas
    for j in range(B):
        r, c = np.nonzero(masks_bool[j])
        seg[r, c] = labels[j]
    return seg


def upsample_threshold_paint(low_res: np.ndarray, labels, canvas: Optional[np.ndarray] = None) -> np.ndarray:
    """numpy restatement of the fused epilogue (bilinear x4 -> >0 -> painter) for
    1024x1024 tiles, following ATen's align_corners=False rule
    (ATen/native/UpSample.h `area_pixel_compute_source_index`):
    src = 0.25*(dst+0.5)-0.5 clamped at 0, i1 = min(i0+1, 255)."""
    B, h, wd = low_res.shape
    assert h == 256 and wd == 256
    dst = np.arange(1024, dtype=np.float32)
    src = np.maximum(np.float32(0.25) * (dst + np.float32(0.5)) - np.float32(0.5), np.float32(0))
    i0 = src.astype(np.int64)
    i1 = np.minimum(i0 + 1, 255)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    seg = np.full((1024, 1024), 255, dtype=np.uint8) if canvas is None else canvas
    for j in range(B):
        a = low_res[j].astype(np.float32)
        # ATen: w0y*(w0x*a00 + w1x*a01) + w1y*(w0x*a10 + w1x*a11)
        r0 = l0[None, :] * a[:, i0] + l1[None, :] * a[:, i1]            # (256,1024)
        full = l0[:, None] * r0[i0, :] + l1[:, None] * r0[i1, :]
        seg[full > 0.0] = labels[j]
    return seg
