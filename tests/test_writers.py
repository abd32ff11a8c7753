"""Host-side output writers: gray / color PNG + instance pickle, against the driver's own
numpy painter semantics restated in the oracle."""
import pickle

import numpy as np
import torch
from PIL import Image

from oracle import rle_oracle, sam_oracle
from samrs_b200 import rle as host_rle
from samrs_b200 import writers

MAPPING = {255: (255, 255, 255), 0: (0, 0, 63), 1: (0, 63, 0), 2: (0, 127, 63), 7: (0, 63, 127)}


def test_save_tile_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    h = w = 96
    masks = np.zeros((3, h, w), bool)
    masks[0, 10:40, 5:50] = True
    masks[1, 30:70, 40:90] = True
    masks[2] = rng.random((h, w)) < 0.02
    labels = [1, 7, 2]
    # the driver's painter (main_sam_hbox_semantic.py:162-163,195-198): later boxes overwrite earlier ones
    seg_mask = np.full((h, w), 255, np.uint8)
    seg_color = np.full((h, w, 3), 255, np.uint8)
    for j in range(3):
        r, c = np.nonzero(masks[j])
        seg_mask[r, c] = labels[j]
        seg_color[r, c] = MAPPING[labels[j]]
    assert np.array_equal(sam_oracle.painter_reduce(masks, np.asarray(labels)), seg_mask)
    counts, offsets, area = rle_oracle.encode_batch(masks)
    boxes = np.asarray([[5, 10, 49, 39], [40, 30, 89, 69], [0, 0, 95, 95]], np.float32)
    recs = host_rle.instance_records(torch.from_numpy(counts), torch.from_numpy(offsets), torch.from_numpy(area), h, w, boxes, labels,
                                     categories=[f"c{i}" for i in range(8)])
    g, c, p = writers.save_tile(str(tmp_path), "P0001", seg_mask, MAPPING, recs)
    assert np.array_equal(np.asarray(Image.open(g)), seg_mask)
    assert np.array_equal(np.asarray(Image.open(c)), seg_color)
    back = pickle.load(open(p, "rb"))
    assert [set(r) for r in back] == [{"mask", "bbox", "category", "label", "size"}] * 3
    for j, r in enumerate(back):
        assert r["label"] == labels[j] and r["category"] == f"c{labels[j]}" and r["size"] == int(masks[j].sum())
        assert np.array_equal(r["bbox"], boxes[j]) and r["mask"]["size"] == [h, w]
        m = host_rle.rle_to_mask({"size": [h, w], "counts": host_rle.coco_string_decode(r["mask"]["counts"])})
        assert np.array_equal(m, masks[j])


def test_colorize_rejects_bad_input():
    import pytest
    with pytest.raises(ValueError):
        writers.colorize(np.zeros((4, 4), np.int32), MAPPING)
    with pytest.raises(ValueError):
        writers.color_lut({300: (1, 2, 3)})
