"""Generates tests/golden/rle_cases.npz from the REFERENCE's own `mask_to_rle_pytorch` (run in the build container,
where /root/reference exists; the fixture travels, the reference does not)."""
import os, sys
import numpy as np
import torch

sys.path.insert(0, "/root/reference/Generate Dataset")
from segment_anything.utils.amg import mask_to_rle_pytorch, rle_to_mask  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def blobs(rng, h, w, n):
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), bool)
    for _ in range(n):
        cy, cx, ry, rx = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(1, max(1.5, h / 3)), rng.uniform(1, max(1.5, w / 3))
        m |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    return m


def cases():
    rng = np.random.default_rng(7)
    out = {}
    for name, (h, w) in {"a64": (64, 64), "b37x50": (37, 50), "c96x33": (96, 33), "d1x70": (1, 70), "e70x1": (70, 1), "f33x32": (33, 32)}.items():
        ms = [np.zeros((h, w), bool), np.ones((h, w), bool), rng.random((h, w)) < 0.5, rng.random((h, w)) < 0.05, blobs(rng, h, w, 3)]
        first = np.zeros((h, w), bool); first[0, 0] = True
        last = np.zeros((h, w), bool); last[-1, -1] = True
        col = np.zeros((h, w), bool); col[:, w // 2] = True            # a full column: runs touch both column boundaries
        row = np.zeros((h, w), bool); row[h // 2, :] = True
        chk = (np.add.outer(np.arange(h), np.arange(w)) % 2).astype(bool)
        out[name] = np.stack(ms + [first, last, col, row, chk])
    return out


def main():
    blob = {}
    for name, masks in cases().items():
        rles = mask_to_rle_pytorch(torch.from_numpy(masks))
        counts, offsets = [], [0]
        for m, r in zip(masks, rles):
            assert r["size"] == list(m.shape)
            assert np.array_equal(rle_to_mask(r), m)
            counts.extend(r["counts"])
            offsets.append(len(counts))
        blob[name + "_masks"] = np.packbits(masks, axis=None)
        blob[name + "_shape"] = np.asarray(masks.shape, dtype=np.int64)
        blob[name + "_counts"] = np.asarray(counts, dtype=np.int64)
        blob[name + "_offsets"] = np.asarray(offsets, dtype=np.int64)
        blob[name + "_area"] = masks.reshape(len(masks), -1).sum(1).astype(np.int64)
    path = os.path.join(ROOT, "tests", "golden", "rle_cases.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
