"""Host logic of the run-unchanged harness (samrs_b200/harness.py): path redirection, stub modules, synthetic dataset
trees in the formats the reference's loaders parse.  The drivers themselves run in tests/test_harness_gpu.py."""
import json
import os
import pickle
import sys

import numpy as np
import pytest

from oracle import rle_oracle
from samrs_b200 import annotations, harness

REF_GD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "GD")


def test_path_redirect_serves_python_level_file_api(tmp_path):
    stage = tmp_path / "stage"
    (stage / "sub").mkdir(parents=True)
    (stage / "sub" / "b.txt").write_text("hello")
    (stage / "sub" / "a.txt").write_text("x")
    with harness.PathRedirect({"/root/dataset": str(stage)}):
        assert os.listdir("/root/dataset/sub/") == ["a.txt", "b.txt"]            # sorted: image ids are listing positions
        assert open("/root/dataset/sub/b.txt").read() == "hello"
        assert os.path.exists("/root/dataset/sub/a.txt") and not os.path.exists("/root/dataset/none")
        os.makedirs("/root/dataset/out/gray", exist_ok=True)
        with open(os.path.join("/root/dataset/out/", "gray", "t.pkl"), "wb") as f:
            pickle.dump([1, 2], f)
        from PIL import Image
        Image.fromarray(np.zeros((4, 4), np.uint8)).save("/root/dataset/out/gray/t.png")
        assert np.array(Image.open("/root/dataset/out/gray/t.png")).shape == (4, 4)
    assert (stage / "out" / "gray" / "t.pkl").exists() and (stage / "out" / "gray" / "t.png").exists()
    assert not os.path.exists("/root/dataset/sub/a.txt")                         # patches are gone outside the context
    import builtins
    assert builtins.open.__module__ in ("io", "_io")


def test_pycocotools_stub_run_lengths_match_the_reference_semantics():
    m = harness._pycocotools_mask()
    rng = np.random.default_rng(0)
    for shape in ((7, 5), (64, 48), (1, 9)):
        mask = rng.random(shape) > 0.5
        rle = m.encode(np.asfortranarray(mask.astype(np.uint8)))
        assert rle["size"] == list(shape) and isinstance(rle["counts"], bytes)
        from samrs_b200 import rle as host_rle
        assert host_rle.coco_string_decode(rle["counts"].decode("ascii")) == rle_oracle.mask_to_rle(mask)
        assert np.array_equal(m.decode(rle).astype(bool), mask)
        assert int(m.area(rle)) == int(mask.sum())
    assert m.encode(np.asfortranarray(np.ones((3, 3), np.uint8)))["counts"] == b"09"     # leading zero run, then 9 ones


def test_matplotlib_stub_absorbs_the_drivers_calls():
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot")}
    try:
        harness.install_stubs(force=True)
        import matplotlib.pyplot as plt
        plt.figure(0, figsize=(10, 10))
        ax = plt.gca()
        ax.add_patch(plt.Rectangle((0, 0), 1, 1, edgecolor=(1, 1, 1, 1), lw=2))
        ax.imshow(np.zeros((2, 2, 4)))
        plt.savefig("/nonexistent/never_written.png", bbox_inches="tight")
        plt.close()
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_synthetic_trees_parse_with_our_loaders_and_the_reference_loaders(tmp_path):
    ds = str(tmp_path / "dataset")
    dior = harness.make_dior(ds, counts=(5, 2), size=96)
    fair = harness.make_fair1m(ds, counts=(4,), size=128)
    hrsc = harness.make_hrsc(ds, counts=(3,), sizes=((200, 320),))
    ann = os.path.join(ds, "dior", "Annotations", "Horizontal Bounding Boxes")
    d = annotations.load_dior(dior[0], ann, harness.DIOR_NAMES)
    assert d.hboxes.shape == (5, 4) and d.hboxes.dtype == np.float32 and len(d.labels) == 5 and d.error == 0
    f = annotations.load_dota(fair[0], os.path.join(ds, "fair1m_1024", "trainval", "rbbtxts"))
    assert f.rboxes.shape == (4, 4, 2) and f.error == 0
    h = annotations.load_hrsc(hrsc[0], os.path.join(ds, "HRSC2016", "Test", "Annotations"))
    assert h.rboxes.shape == (3, 4, 2) and h.colors.shape == (3, 3) and h.error == 0
    if not os.path.isdir(REF_GD):
        pytest.skip("reference loaders not staged (oracle/_ref/GD is created by build() in the build container)")
    sys.path.insert(0, REF_GD)
    try:
        for k in ("loaddata", "mapping", "utils", "utils.transform"):
            sys.modules.pop(k, None)
        import loaddata
        hb, pts, labs, err = loaddata.load_dior(dior[0], ann)
        assert err == 0 and np.array_equal(np.stack(hb), d.hboxes) and list(labs) == list(d.labels)
        _, rb, _, labs2, err2 = loaddata.load_dota(fair[0], os.path.join(ds, "fair1m_1024", "trainval", "rbbtxts"))
        assert err2 == 0 and np.array_equal(np.stack(rb), f.rboxes) and list(labs2) == list(f.labels)
    finally:
        sys.path.remove(REF_GD)
        for k in ("loaddata", "mapping", "utils", "utils.transform"):
            sys.modules.pop(k, None)


def test_harness_fixtures_are_committed(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "harness", "meta.json")))
    assert set(meta["drivers"]) == set(harness.DRIVERS)
    assert meta["hbox"]["11726"]["keys"] == ["bbox", "category", "label", "mask", "size"]
    assert sorted(meta["rhbox"])[0].startswith("0__1024") and meta["rhbox"]["0__1024__0___0"]["keys"] == ["category", "label", "mask", "rbox", "rhbox", "size"]
