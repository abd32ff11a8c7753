"""Model geometry of the SAM variants on the hot path.

Mirrors the three registry builders of the reference
(`Generate Dataset/segment_anything/build_sam.py:14-44`) plus two tiny
geometries that exist only so parity tests finish in seconds on a CPU oracle.
Everything that is not listed here is fixed by the reference's `_build_sam`
(`build_sam.py:55-101`): 1024x1024 input, 16x16 patches (64x64 token grid),
14x14 attention windows, 256 prompt/neck channels, two-way decoder depth 2,
8 decoder heads, MLP dim 2048, 4 mask tokens.
"""
from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class SamGeometry:
    name: str
    embed_dim: int
    depth: int
    num_heads: int
    global_attn_indexes: Tuple[int, ...]

    # fixed by the reference's _build_sam
    img_size: int = 1024
    patch: int = 16
    window: int = 14
    out_chans: int = 256
    mlp_ratio: int = 4

    @property
    def grid(self) -> int:
        return self.img_size // self.patch

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads


GEOMETRIES = {
    "vit_h": SamGeometry("vit_h", 1280, 32, 16, (7, 15, 23, 31)),
    "vit_l": SamGeometry("vit_l", 1024, 24, 16, (5, 11, 17, 23)),
    "vit_b": SamGeometry("vit_b", 768, 12, 12, (2, 5, 8, 11)),
    # test-only geometries (not in the reference registry): same code paths,
    # head_dim 64 and 80, one windowed + one global block each.
    "vit_t64": SamGeometry("vit_t64", 128, 2, 2, (1,)),
    "vit_t80": SamGeometry("vit_t80", 160, 3, 2, (2,)),
}
GEOMETRIES["default"] = GEOMETRIES["vit_h"]


def geometry(name: str) -> SamGeometry:
    try:
        return GEOMETRIES[name]
    except KeyError:
        raise KeyError(f"unknown SAM variant {name!r}; known: {sorted(GEOMETRIES)}") from None
