"""Annotation front-end: DOTA txt / DIOR xml / HRSC xml -> packed arrays ready for one H2D copy.

Mirrors `Generate Dataset/loaddata.py` (`load_dota` :103-132, `load_dior` :10-39, `load_hrsc` :41-101): same fields, same
values, same `error` flag; but every field comes back as ONE contiguous array per image (boxes `(B,4)` float32, polygons
`(B,4,2)`, points `(B,2)`, labels `(B,)` int64) instead of a Python list of small arrays, so a tile's prompts are a single
pinned buffer.  Class tables are the datasets' own (`mapping.py`) and are passed in.

Reference quirks kept on purpose: DOTA's hbox is `[x1, y1, x3, y3]` (first and third polygon vertex, :122) and its
point their midpoint (:124); DIOR lower-cases the class name (:23) and falls back to `robndbox` when `bndbox` is missing
(:24-26); HRSC is single-class (label 0, :94), converts `(cx, cy, w, h, angle)` with the le90 vertex order of
`utils/transform.py:193-216` followed by the best-begin-point rotation (`:234-262`), and flags a malformed
`seg_color` (:77-79).
"""
from __future__ import annotations

import math
import os.path as osp
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np


@dataclass
class Annotations:
    hboxes: np.ndarray                              # (B,4) float32 xyxy
    labels: np.ndarray                              # (B,) int64
    points: np.ndarray                              # (B,2) centre points
    rboxes: Optional[np.ndarray] = None             # (B,4,2) polygon vertices, DOTA / HRSC
    colors: Optional[np.ndarray] = None             # (B,3) uint8, HRSC
    classes: list = field(default_factory=list)     # class names as written in the file (DOTA)
    error: int = 0                                  # the loaders' flag: 1 = nothing usable (or malformed colour)

    def __len__(self) -> int:
        return int(self.hboxes.shape[0])


def _arr(rows, shape, dtype):
    return np.asarray(rows, dtype=dtype).reshape((len(rows),) + shape) if rows else np.zeros((0,) + shape, dtype=dtype)


def load_dota(img_name: str, ann_path: str) -> Annotations:
    """`<ann_path>/<img_name>.txt`: lines `x1 y1 x2 y2 x3 y3 x4 y4 class_name class_index`."""
    hb, rb, pt, cl, lb = [], [], [], [], []
    with open(osp.join(ann_path, img_name + ".txt"), "r") as f:
        for line in f.readlines():
            x1, y1, x2, y2, x3, y3, x4, y4, name, index = line.strip().split()
            x1, y1, x2, y2, x3, y3, x4, y4 = (float(v) for v in (x1, y1, x2, y2, x3, y3, x4, y4))
            hb.append([x1, y1, x3, y3])
            rb.append([[x1, y1], [x2, y2], [x3, y3], [x4, y4]])
            pt.append([(x1 + x3) / 2, (y1 + y3) / 2])
            cl.append(name)
            lb.append(int(index))
    return Annotations(_arr(hb, (4,), np.float32), _arr(lb, (), np.int64), _arr(pt, (2,), np.float64), _arr(rb, (4, 2), np.float64),
                       None, cl, 1 if not hb else 0)


def load_dior(img_name: str, ann_path: str, classes: Sequence[str]) -> Annotations:
    """`<ann_path>/<img_name>.xml` (VOC style); `classes` = the DIOR class tuple, label = index of the lower-cased name."""
    cls2lbl = {k: v for v, k in enumerate(classes)}
    root = ET.parse(osp.join(ann_path, f"{img_name}.xml")).getroot()
    hb, pt, lb = [], [], []
    for obj in root.findall("object"):
        category = str(obj.find("name").text.lower())
        bnd = obj.find("bndbox")
        if bnd is None or len(bnd) == 0:               # the reference tests Element truthiness: missing or childless
            bnd = obj.find("robndbox")
        xmin, ymin, xmax, ymax = (float(bnd.find(k).text) for k in ("xmin", "ymin", "xmax", "ymax"))
        hb.append([xmin, ymin, xmax, ymax])
        pt.append([(xmin + xmax) / 2, (ymin + ymax) / 2])
        lb.append(int(cls2lbl[category]))
    return Annotations(_arr(hb, (4,), np.float32), _arr(lb, (), np.int64), _arr(pt, (2,), np.float64), error=1 if not hb else 0)


def obb_to_polygon_le90(cx: float, cy: float, w: float, h: float, theta: float) -> np.ndarray:
    """(4,2) float32 vertices of an oriented box in the reference's order (`utils/transform.py:203-216`, float32
    arithmetic as there because the loader builds a float32 row) rotated to the best begin point (`:234-262`)."""
    f = np.float32
    c, s = np.cos(f(theta)), np.sin(f(theta))
    centre = np.array([cx, cy], dtype=f)
    v1 = np.array([f(w) / 2 * c, f(w) / 2 * s], dtype=f)
    v2 = np.array([-f(h) / 2 * s, f(h) / 2 * c], dtype=f)
    pts = [centre - v1 - v2, centre + v1 - v2, centre + v1 + v2, centre - v1 + v2]
    xs, ys = [float(p[0]) for p in pts], [float(p[1]) for p in pts]
    dst = [(min(xs), min(ys)), (max(xs), min(ys)), (max(xs), max(ys)), (min(xs), max(ys))]
    best, best_i = 100000000.0, 0
    for i in range(4):
        force = sum(math.sqrt(math.pow(xs[(i + k) % 4] - dst[k][0], 2) + math.pow(ys[(i + k) % 4] - dst[k][1], 2)) for k in range(4))
        if force < best:
            best, best_i = force, i
    return np.array([[xs[(best_i + k) % 4], ys[(best_i + k) % 4]] for k in range(4)], dtype=np.float32)


def load_hrsc(img_name: str, ann_path: str) -> Annotations:
    """`<ann_path>/<img_name>.xml` with `HRSC_Objects/HRSC_Object` entries; single class."""
    root = ET.parse(osp.join(ann_path, f"{img_name}.xml")).getroot()
    hb, rb, co, pt = [], [], [], []
    error = 0
    for obj in root.findall("HRSC_Objects/HRSC_Object"):
        g = lambda k: float(obj.find(k).text)  # noqa: E731
        hb.append([g("box_xmin"), g("box_ymin"), g("box_xmax"), g("box_ymax")])
        rb.append(obb_to_polygon_le90(g("mbox_cx"), g("mbox_cy"), g("mbox_w"), g("mbox_h"), g("mbox_ang")))
        colour = obj.find("seg_color").text.split(",")
        if len(colour) != 3:
            error = 1
            co.append([0, 0, 0])
        else:
            co.append([int(v) for v in colour])
        pt.append([g("mbox_cx"), g("mbox_cy")])
    if not hb:
        error = 1
    return Annotations(_arr(hb, (4,), np.float32), np.zeros((len(hb),), np.int64), _arr(pt, (2,), np.float32), _arr(rb, (4, 2), np.float32),
                       _arr(co, (3,), np.uint8), [], error)
