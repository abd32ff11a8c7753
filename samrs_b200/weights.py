"""State-dict layout of `Sam` and a seeded synthetic checkpoint generator.

The engine ingests weights in exactly the key/shape layout of the reference's
`Sam.state_dict()` (constructed by
`Generate Dataset/segment_anything/build_sam.py:55-101`), so a file written by
`save_checkpoint` loads through the reference's own `torch.load` +
`load_state_dict(strict)` (`build_sam.py:102-106`) and through ours.

No SAM checkpoint exists offline, so every parity and bench run uses
`synthetic_state_dict`: each tensor is drawn from its own
`torch.Generator(seed ^ crc32(key))`, which makes the checkpoint reproducible
on any box without shipping 2.5 GB, and independent of construction order.
Distributions follow torch's default module init scale (uniform
+-1/sqrt(fan_in) for linear/conv, N(0,1) embeddings) so activations and logit
magnitudes resemble the reference constructor's; parameters the reference
zero- or one-initialises (rel_pos_*, pos_embed, LayerNorm affine) are
perturbed so those code paths are actually exercised.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Tuple

import torch

from .config import SamGeometry, geometry

Spec = List[Tuple[str, Tuple[int, ...], str]]  # (key, shape, kind)


def state_dict_spec(g: SamGeometry) -> Spec:
    """Ordered (key, shape, kind) list equal to `Sam.state_dict()` of the reference."""
    D, hd, G = g.embed_dim, g.head_dim, g.grid
    s: Spec = []
    e = "image_encoder."
    s.append((e + "pos_embed", (1, G, G, D), "small"))
    s.append((e + "patch_embed.proj.weight", (D, 3, g.patch, g.patch), "fan"))
    s.append((e + "patch_embed.proj.bias", (D,), "fanb:%d" % (3 * g.patch * g.patch)))
    for i in range(g.depth):
        b = f"{e}blocks.{i}."
        S = G if i in g.global_attn_indexes else g.window
        s.append((b + "norm1.weight", (D,), "ln_w"))
        s.append((b + "norm1.bias", (D,), "ln_b"))
        s.append((b + "attn.rel_pos_h", (2 * S - 1, hd), "relpos"))
        s.append((b + "attn.rel_pos_w", (2 * S - 1, hd), "relpos"))
        s.append((b + "attn.qkv.weight", (3 * D, D), "fan"))
        s.append((b + "attn.qkv.bias", (3 * D,), f"fanb:{D}"))
        s.append((b + "attn.proj.weight", (D, D), "fan"))
        s.append((b + "attn.proj.bias", (D,), f"fanb:{D}"))
        s.append((b + "norm2.weight", (D,), "ln_w"))
        s.append((b + "norm2.bias", (D,), "ln_b"))
        s.append((b + "mlp.lin1.weight", (g.mlp_ratio * D, D), "fan"))
        s.append((b + "mlp.lin1.bias", (g.mlp_ratio * D,), f"fanb:{D}"))
        s.append((b + "mlp.lin2.weight", (D, g.mlp_ratio * D), "fan"))
        s.append((b + "mlp.lin2.bias", (D,), f"fanb:{g.mlp_ratio * D}"))
    C = g.out_chans
    s.append((e + "neck.0.weight", (C, D, 1, 1), "fan"))
    s.append((e + "neck.1.weight", (C,), "ln_w"))
    s.append((e + "neck.1.bias", (C,), "ln_b"))
    s.append((e + "neck.2.weight", (C, C, 3, 3), "fan"))
    s.append((e + "neck.3.weight", (C,), "ln_w"))
    s.append((e + "neck.3.bias", (C,), "ln_b"))

    p = "prompt_encoder."
    s.append((p + "pe_layer.positional_encoding_gaussian_matrix", (2, C // 2), "normal"))
    for i in range(4):
        s.append((p + f"point_embeddings.{i}.weight", (1, C), "normal"))
    s.append((p + "not_a_point_embed.weight", (1, C), "normal"))
    s.append((p + "mask_downscaling.0.weight", (4, 1, 2, 2), "fan"))
    s.append((p + "mask_downscaling.0.bias", (4,), "fanb:4"))
    s.append((p + "mask_downscaling.1.weight", (4,), "ln_w"))
    s.append((p + "mask_downscaling.1.bias", (4,), "ln_b"))
    s.append((p + "mask_downscaling.3.weight", (16, 4, 2, 2), "fan"))
    s.append((p + "mask_downscaling.3.bias", (16,), "fanb:16"))
    s.append((p + "mask_downscaling.4.weight", (16,), "ln_w"))
    s.append((p + "mask_downscaling.4.bias", (16,), "ln_b"))
    s.append((p + "mask_downscaling.6.weight", (C, 16, 1, 1), "fan"))
    s.append((p + "mask_downscaling.6.bias", (C,), "fanb:16"))
    s.append((p + "no_mask_embed.weight", (1, C), "normal"))

    m = "mask_decoder."

    def attn(prefix: str, internal: int) -> None:
        for nm in ("q_proj", "k_proj", "v_proj"):
            s.append((f"{prefix}.{nm}.weight", (internal, C), "fan"))
            s.append((f"{prefix}.{nm}.bias", (internal,), f"fanb:{C}"))
        s.append((f"{prefix}.out_proj.weight", (C, internal), "fan"))
        s.append((f"{prefix}.out_proj.bias", (C,), f"fanb:{internal}"))

    for i in range(2):
        L = f"{m}transformer.layers.{i}"
        attn(L + ".self_attn", C)
        s.append((L + ".norm1.weight", (C,), "ln_w"))
        s.append((L + ".norm1.bias", (C,), "ln_b"))
        attn(L + ".cross_attn_token_to_image", C // 2)
        s.append((L + ".norm2.weight", (C,), "ln_w"))
        s.append((L + ".norm2.bias", (C,), "ln_b"))
        s.append((L + ".mlp.lin1.weight", (2048, C), "fan"))
        s.append((L + ".mlp.lin1.bias", (2048,), f"fanb:{C}"))
        s.append((L + ".mlp.lin2.weight", (C, 2048), "fan"))
        s.append((L + ".mlp.lin2.bias", (C,), "fanb:2048"))
        s.append((L + ".norm3.weight", (C,), "ln_w"))
        s.append((L + ".norm3.bias", (C,), "ln_b"))
        s.append((L + ".norm4.weight", (C,), "ln_w"))
        s.append((L + ".norm4.bias", (C,), "ln_b"))
        attn(L + ".cross_attn_image_to_token", C // 2)
    attn(m + "transformer.final_attn_token_to_image", C // 2)
    s.append((m + "transformer.norm_final_attn.weight", (C,), "ln_w"))
    s.append((m + "transformer.norm_final_attn.bias", (C,), "ln_b"))
    s.append((m + "iou_token.weight", (1, C), "normal"))
    s.append((m + "mask_tokens.weight", (4, C), "normal"))
    s.append((m + "output_upscaling.0.weight", (C, C // 4, 2, 2), "fanT"))
    s.append((m + "output_upscaling.0.bias", (C // 4,), f"fanb:{C // 4 * 4}"))
    s.append((m + "output_upscaling.1.weight", (C // 4,), "ln_w"))
    s.append((m + "output_upscaling.1.bias", (C // 4,), "ln_b"))
    s.append((m + "output_upscaling.3.weight", (C // 4, C // 8, 2, 2), "fanT"))
    s.append((m + "output_upscaling.3.bias", (C // 8,), f"fanb:{C // 8 * 4}"))
    for i in range(4):
        H = f"{m}output_hypernetworks_mlps.{i}.layers"
        for j, (o, k) in enumerate(((C, C), (C, C), (C // 8, C))):
            s.append((f"{H}.{j}.weight", (o, k), "fan"))
            s.append((f"{H}.{j}.bias", (o,), f"fanb:{k}"))
    H = f"{m}iou_prediction_head.layers"
    for j, (o, k) in enumerate(((C, C), (C, C), (4, C))):
        s.append((f"{H}.{j}.weight", (o, k), "fan"))
        s.append((f"{H}.{j}.bias", (o,), f"fanb:{k}"))
    return s


def _draw(key: str, shape: Tuple[int, ...], kind: str, seed: int) -> torch.Tensor:
    gen = torch.Generator(device="cpu")
    gen.manual_seed((seed * 0x9E3779B1 ^ zlib.crc32(key.encode())) & 0x7FFFFFFF)
    if kind == "normal":
        return torch.randn(shape, generator=gen, dtype=torch.float32)
    if kind == "small":
        return 0.02 * torch.randn(shape, generator=gen, dtype=torch.float32)
    if kind == "relpos":
        # the reference zero-inits these (image_encoder.py:221-222); a zero table
        # would leave the decomposed rel-pos path untested.
        return 0.05 * torch.randn(shape, generator=gen, dtype=torch.float32)
    if kind == "ln_w":
        return 1.0 + 0.1 * torch.randn(shape, generator=gen, dtype=torch.float32)
    if kind == "ln_b":
        return 0.05 * torch.randn(shape, generator=gen, dtype=torch.float32)
    if kind in ("fan", "fanT") or kind.startswith("fanb:"):
        if kind == "fan":
            fan_in = math.prod(shape[1:])
        elif kind == "fanT":  # ConvTranspose2d weight is (in, out, kh, kw); torch's fan_in uses dim 1
            fan_in = shape[1] * shape[2] * shape[3]
        else:
            fan_in = int(kind.split(":")[1])
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound
    raise ValueError(kind)


def synthetic_state_dict(variant: str, seed: int = 0, logit_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic `Sam.state_dict()` for `variant` (CPU fp32 tensors).

    `logit_scale` != 1 multiplies the last layer (weight and bias) of the four hyper-network MLPs, which scales
    every mask logit by that factor: default-init weights give |logit| <= 0.4, a trained SAM O(10);
    parity cases with `logit_scale = 32` state their error relative to that magnitude."""
    g = geometry(variant)
    sd = {k: _draw(k, shp, kind, seed) for k, shp, kind in state_dict_spec(g)}
    if logit_scale != 1.0:
        scale_logits_(sd, logit_scale)
    return sd


def scale_logits_(sd: Dict[str, torch.Tensor], s: float) -> None:
    """In place: hyper-network output layers x s  =>  low-res mask logits x s (mask_decoder.py:156-167 is linear in them)."""
    for i in range(4):
        for nm in ("weight", "bias"):
            sd[f"mask_decoder.output_hypernetworks_mlps.{i}.layers.2.{nm}"] *= s


def check_state_dict(variant, sd: Dict[str, torch.Tensor]) -> None:
    """Strict key/shape check, same contract as `load_state_dict(strict=True)`."""
    spec = state_dict_spec(variant if isinstance(variant, SamGeometry) else geometry(variant))
    want = {k: shp for k, shp, _ in spec}
    missing = [k for k in want if k not in sd]
    unexpected = [k for k in sd if k not in want]
    if missing or unexpected:
        raise RuntimeError(
            f"Error(s) in loading state_dict for Sam: missing {missing[:5]}"
            f"{'...' if len(missing) > 5 else ''}; unexpected {unexpected[:5]}"
        )
    for k, shp in want.items():
        if tuple(sd[k].shape) != tuple(shp):
            raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(sd[k].shape)} vs model {shp}")


def save_checkpoint(path: str, variant: str, seed: int = 0) -> None:
    """Write a reference-layout checkpoint file (what `sam_vit_h_4b8939.pth` is to the drivers)."""
    torch.save(synthetic_state_dict(variant, seed), path)
