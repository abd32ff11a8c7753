"""`samrs_b200.stream.run` - the per-rank tile loop that replaces `main_sam_hbox_semantic.py:110-216`: one encode per tile
feeds the label map AND the instance payload; loader / finisher / writer threads around the GPU thread."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from samrs_b200 import rle as host_rle  # noqa: E402
from samrs_b200 import stream, synth  # noqa: E402
from samrs_b200.weights import synthetic_state_dict  # noqa: E402


@pytest.fixture(scope="module")
def predictor():
    import samrs_b200
    sys.path.insert(0, samrs_b200.DROPIN_PATH)
    for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
        del sys.modules[k]
    from segment_anything import SamPredictor
    from segment_anything.modeling import Sam
    from samrs_b200.config import geometry
    sam = Sam(geometry("vit_t80"))
    sam.load_state_dict(synthetic_state_dict("vit_t80", 0))
    sam.to("cuda")
    yield SamPredictor(sam)
    sys.path.remove(samrs_b200.DROPIN_PATH)
    for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
        del sys.modules[k]


def test_stream_run_matches_the_per_tile_helpers_and_writes_the_drivers_files(predictor, tmp_path):
    from PIL import Image
    eng = predictor.model.engine
    mapping = {i: (i, 2 * i, 3 * i) for i in range(18)}
    mapping[255] = (255, 255, 255)
    cats = [f"c{i}" for i in range(18)]
    shapes = [(1024, 1024), (1024, 1024), (600, 800), (1024, 1024), (768, 1024)]
    jobs, want = [], {}
    for t, (H, W) in enumerate(shapes):
        img = synth.tile(40 + t)[:H, :W].copy()
        n = 23 if t != 1 else 5
        boxes = synth.hboxes(40 + t, n, size=min(H, W))
        labels = synth.labels(40 + t, n)
        rb = synth.rbox_polys(40 + t, n, size=min(H, W)) if t == 3 else None
        jobs.append(stream.TileJob(f"tile{t}", (lambda a=img: a) if t % 2 else img, boxes, labels, rb))
        # expectation from the drop-in predictor exactly as a driver would use it (bool masks on the host, numpy painter)
        predictor.set_image(img)
        seg = np.full((H, W), 255, np.uint8)
        areas = []
        for s in range(0, n, 20):
            tb = predictor.transform.apply_boxes_torch(torch.from_numpy(boxes[s:s + 20]).cuda(), (H, W))
            masks, _, _ = predictor.predict_torch(None, None, boxes=tb, mask_input=None, multimask_output=False)
            m = masks[:, 0].cpu().numpy()
            for j in range(m.shape[0]):
                seg[m[j]] = labels[s + j]
                areas.append((int(m[j].sum()), host_rle.mask_to_counts(m[j])))
        want[f"tile{t}"] = (seg, areas)
    seen = {}
    stats = stream.run(predictor, jobs, str(tmp_path), mapping, cats, chunk=20, writer_threads=3, loader_threads=2, depth=2,
                       on_tile=lambda job, lm, recs: seen.__setitem__(job.name, (lm, recs)))
    assert stats["tiles"] == len(shapes) and stats["masks"] == 23 * 4 + 5
    for name, (seg, areas) in want.items():
        lm, recs = seen[name]
        assert np.array_equal(lm, seg), name
        assert np.array_equal(np.array(Image.open(tmp_path / "gray" / (name + ".png"))), seg)
        color = np.array(Image.open(tmp_path / "color" / (name + ".png")))
        assert (color[seg == 255] == 255).all() and color.shape == seg.shape + (3,)
        disk = pickle.load(open(tmp_path / "ins" / (name + ".pkl"), "rb"))
        assert len(disk) == len(areas) == len(recs)
        for r, (a, counts) in zip(disk, areas):
            assert r["size"] == a and host_rle.coco_string_decode(r["mask"]["counts"]) == counts and r["mask"]["size"] == list(seg.shape)
        if name == "tile3":
            assert sorted(disk[0].keys()) == ["category", "label", "mask", "rbox", "rhbox", "size"] and disk[0]["rbox"].shape == (4, 2)
        else:
            assert sorted(disk[0].keys()) == ["bbox", "category", "label", "mask", "size"]
    eng.decode(boxes=torch.zeros(1, 4, device="cuda"), multimask_output=False)       # engine still usable afterwards
