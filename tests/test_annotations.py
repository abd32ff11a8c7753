"""Annotation front-end against what the reference's own loaders returned for the same files
(oracle/make_golden_annotations.py -> tests/golden/annotations/)."""
import os

import numpy as np

from samrs_b200 import annotations as A

D = os.path.join(os.path.dirname(__file__), "golden", "annotations")
Z = np.load(os.path.join(D, "expected.npz"))


def test_dota_matches_reference_loader():
    a = A.load_dota("P0001", D)
    assert a.error == int(Z["dota_error"]) == 0 and len(a) == 6
    assert a.hboxes.dtype == np.float32 and a.hboxes.flags["C_CONTIGUOUS"]
    assert np.array_equal(a.hboxes, Z["dota_hboxes"].astype(np.float32))        # the driver casts to f32 in apply_boxes_torch
    assert np.array_equal(a.rboxes, Z["dota_rboxes"]) and np.array_equal(a.points, Z["dota_points"])
    assert np.array_equal(a.labels, Z["dota_labels"]) and a.classes == ["ship"] * 6
    e = A.load_dota("P0002", D)
    assert e.error == int(Z["dota_empty_error"]) == 1 and e.hboxes.shape == (0, 4) and e.rboxes.shape == (0, 4, 2)


def test_dior_matches_reference_loader():
    a = A.load_dior("00011", D, [str(c) for c in Z["dior_classes"]])
    assert a.error == int(Z["dior_error"]) == 0
    assert np.array_equal(a.hboxes, Z["dior_hboxes"]) and np.array_equal(a.points, Z["dior_points"])
    assert np.array_equal(a.labels, Z["dior_labels"])


def test_hrsc_matches_reference_loader():
    a = A.load_hrsc("100000001", D)
    assert a.error == int(Z["hrsc_error"]) == 1                                  # one malformed seg_color
    assert np.array_equal(a.hboxes, Z["hrsc_hboxes"]) and np.array_equal(a.points, Z["hrsc_points"])
    assert np.array_equal(a.colors, Z["hrsc_colors"]) and np.array_equal(a.labels, Z["hrsc_labels"])
    assert a.rboxes.dtype == np.float32 and np.array_equal(a.rboxes, Z["hrsc_rboxes"])   # vertex order incl. best begin point
