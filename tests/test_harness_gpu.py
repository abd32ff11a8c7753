"""The three SAMRS generation drivers, UNMODIFIED, over the drop-in package on a B200 (north_star: "run unchanged").

`oracle/_ref/GD/main_sam_*.py` are byte-identical copies of the reference's scripts (staged by oracle/stage_ref.py at build
time; they read the hard-coded /root/dataset and /root/dw paths, served here from a temp dir by `harness.PathRedirect`).
Their outputs are compared with the outputs of THE SAME scripts run over the reference's own package on the CPU
(tests/golden/harness/, oracle/make_golden_harness.py): label PNGs by differing pixels, instance areas, COCO-RLE masks."""
import json
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from samrs_b200 import harness  # noqa: E402
from samrs_b200 import rle as host_rle  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GD = os.path.join(ROOT, "oracle", "_ref", "GD")
FLIP_FRAC = 2e-3          # label-map pixels allowed to differ (mask logits within ~1e-4 of 0 flip)


@pytest.fixture(scope="module")
def staged(tmp_path_factory):
    if not os.path.isdir(GD):
        pytest.fail("oracle/_ref/GD missing: run build() in the build container so the driver scripts travel with the snapshot")
    stage = str(tmp_path_factory.mktemp("samrs_stage"))
    return harness.stage_all(stage)


def _png(path):
    from PIL import Image
    return np.array(Image.open(path))


def _semantic(tag, script, save_rel, staged, golden_dir, argv=()):
    harness.run_driver(os.path.join(GD, script), list(argv), redirect=staged)
    save = os.path.join(staged["/root/dataset"], save_rel)
    meta = json.load(open(os.path.join(golden_dir, "harness", "meta.json")))[tag]
    for name, want in meta.items():
        gray, ref = _png(os.path.join(save, "gray", name + ".png")), _png(os.path.join(golden_dir, "harness", f"{tag}_gray_{name}.png"))
        assert gray.shape == ref.shape and gray.dtype == np.uint8
        diff = int((gray != ref).sum())
        color = _png(os.path.join(save, "color", name + ".png"))
        recs = pickle.load(open(os.path.join(save, "ins", name + ".pkl"), "rb"))
        areas = [int(r["size"]) for r in recs]
        rel = max(abs(a - b) / max(b, 1) for a, b in zip(areas, want["size"]))
        print(f"{script} {name}: label-map pixels differing {diff} of {gray.size} ({diff / gray.size:.2e}); "
              f"{len(recs)} instances, max relative area difference {rel:.2e}")
        assert diff <= FLIP_FRAC * gray.size
        assert color.shape == gray.shape + (3,) and (color[gray == 255] == 255).all()
        assert sorted(recs[0].keys()) == want["keys"] and [int(r["label"]) for r in recs] == want["label"]
        assert [r["category"] for r in recs] == want["category"] and rel < 0.02
        m0 = host_rle.rle_to_mask({"size": recs[0]["mask"]["size"], "counts": host_rle.coco_string_decode(recs[0]["mask"]["counts"])})
        assert m0.shape == gray.shape and int(m0.sum()) == areas[0]


def test_hbox_semantic_driver_runs_unchanged(staged, golden_dir):
    _semantic("hbox", "main_sam_hbox_semantic.py", os.path.join("dior", "hbox_segs_test_init"), staged, golden_dir)


def test_rhbox_semantic_driver_runs_unchanged(staged, golden_dir):
    _semantic("rhbox", "main_sam_rhbox_semantic.py", os.path.join("fair1m_1024", "trainval", "rhbox_segs_init"), staged, golden_dir)


def test_rbox_mask_instance_driver_runs_unchanged(staged, golden_dir):
    harness.run_driver(os.path.join(GD, "main_sam_rbox_mask_instance.py"), ["--show", "False"], redirect=staged)
    out = json.load(open(os.path.join(staged["/root/dw"], "samrs", "work_dir", "hrsc", "json", "sam_ins_rbox.json")))
    ref = json.load(open(os.path.join(golden_dir, "harness", "rbox_sam_ins_rbox.json")))
    assert len(out) == len(ref) and [o["image_id"] for o in out] == [r["image_id"] for r in ref]
    tot = flips = 0
    for o, r in zip(out, ref):
        dec = lambda s: host_rle.rle_to_mask({"size": s["size"], "counts": host_rle.coco_string_decode(s["counts"])})
        a, b = dec(o["segmentation"]), dec(r["segmentation"])
        assert a.shape == b.shape
        flips += int((a != b).sum())
        tot += a.size
        assert abs(o["score"] - r["score"]) < 1e-3
    print(f"main_sam_rbox_mask_instance.py: {len(out)} masks, pixels differing {flips} of {tot} ({flips / tot:.2e})")
    assert flips <= FLIP_FRAC * tot
