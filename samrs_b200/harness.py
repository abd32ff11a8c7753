"""Run the SAMRS generation drivers UNCHANGED against the drop-in `segment_anything` (BASELINE.json north_star).

The drivers (`Generate Dataset/main_sam_hbox_semantic.py`, `main_sam_rhbox_semantic.py`,
`main_sam_rbox_mask_instance.py`) are flat scripts that
  * hard-code dataset, output and checkpoint paths under /root/dataset and /root/dw (`main_sam_hbox_semantic.py:62-83`),
  * build the model at import time (`:87-89`),
  * import `matplotlib.pyplot` and `pycocotools.mask` (`:8,16`), neither of which a generation box needs,
  * put their own directory first on `sys.path`, so the vendored package would shadow any replacement.
This module supplies everything around an unmodified script file:

  `PathRedirect`      maps path prefixes (e.g. /root/dataset -> <stage>/dataset) for the Python-level file API the
                      drivers use (`open`, `os.listdir`, `os.makedirs`, `os.path.exists`, and through `open`: PIL,
                      pickle, json, ElementTree, torch.load) - the script keeps its literal paths;
  `install_stubs`     stand-ins for `matplotlib.pyplot` (draw calls are no-ops) and `pycocotools.mask` (`encode` /
                      `decode` / `area` on `samrs_b200.rle`), only for modules that are not installed;
  `make_*`            seeded synthetic DIOR / FAIR1M / HRSC trees in the exact on-disk formats `loaddata.py:10-132` parses,
                      and the checkpoint file the registry loads;
  `run_driver`        `runpy.run_path(script, run_name="__main__")` with `sys.path = [package dir, script dir, ...]`.

Nothing here touches the engine: the drivers reach it through `sam_model_registry` / `SamPredictor` like any user.
"""
from __future__ import annotations

import builtins
import contextlib
import os
import runpy
import sys
import types
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

DROPIN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
DRIVERS = ("main_sam_hbox_semantic.py", "main_sam_rhbox_semantic.py", "main_sam_rbox_mask_instance.py")
_DRIVER_LOCAL_MODULES = ("segment_anything", "loaddata", "mapping", "instance_to_json", "utils")


# ------------------------------------------------------------------------------------------------ path redirection
class PathRedirect(contextlib.AbstractContextManager):
    """While active, paths starting with a key of `mapping` are served from the mapped directory."""

    def __init__(self, mapping: Dict[str, str]):
        self.mapping = sorted(((os.path.normpath(k), os.path.abspath(v)) for k, v in mapping.items()), key=lambda kv: -len(kv[0]))
        self._saved = []

    def resolve(self, path):
        if not isinstance(path, (str, os.PathLike)):
            return path
        p = os.fspath(path)
        if not isinstance(p, str):
            return path
        n = os.path.normpath(p)
        for src, dst in self.mapping:
            if n == src or n.startswith(src + os.sep):
                out = dst + n[len(src):]
                return out + os.sep if p.endswith(os.sep) and not out.endswith(os.sep) else out
        return path

    def _patch(self, owner, name, make):
        orig = getattr(owner, name)
        self._saved.append((owner, name, orig))
        setattr(owner, name, make(orig))

    def __enter__(self):
        r = self.resolve
        self._patch(builtins, "open", lambda f: (lambda file, *a, **k: f(r(file), *a, **k)))
        self._patch(os, "listdir", lambda f: (lambda path=".": sorted(f(r(path))) if r(path) is not path else f(path)))
        self._patch(os, "makedirs", lambda f: (lambda name, *a, **k: f(r(name), *a, **k)))
        self._patch(os, "mkdir", lambda f: (lambda path, *a, **k: f(r(path), *a, **k)))
        self._patch(os, "stat", lambda f: (lambda path, *a, **k: f(r(path), *a, **k)))
        for fn in ("exists", "isfile", "isdir", "getsize"):
            self._patch(os.path, fn, lambda f: (lambda path: f(r(path))))
        return self

    def __exit__(self, *exc):
        while self._saved:
            owner, name, orig = self._saved.pop()
            setattr(owner, name, orig)
        return False


# ------------------------------------------------------------------------------------------------ stub modules
class _Inert:
    """Absorbs any attribute access / call: every matplotlib call of the drivers is a drawing side effect."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self

    def __call__(self, *a, **k):
        return self


def _pycocotools_mask() -> types.ModuleType:
    from . import rle as host_rle
    m = types.ModuleType("pycocotools.mask")

    def encode(bimask):
        a = np.asarray(bimask)
        if a.ndim == 3:
            return [encode(a[:, :, i]) for i in range(a.shape[2])]
        h, w = a.shape
        return {"size": [int(h), int(w)], "counts": host_rle.coco_string(host_rle.mask_to_counts(a)).encode("ascii")}

    def _counts(r):
        c = r["counts"]
        if isinstance(c, (bytes, str)):
            return host_rle.coco_string_decode(c.decode("ascii") if isinstance(c, bytes) else c)
        return list(c)

    def decode(rle):
        if isinstance(rle, (list, tuple)):
            return np.stack([decode(r) for r in rle], axis=2)
        return np.asfortranarray(host_rle.rle_to_mask({"size": rle["size"], "counts": _counts(rle)}).astype(np.uint8))

    def area(rle):
        if isinstance(rle, (list, tuple)):
            return np.array([area(r) for r in rle], dtype=np.uint32)
        return np.uint32(sum(_counts(rle)[1::2]))

    m.encode, m.decode, m.area = encode, decode, area
    return m


def install_stubs(force: bool = False) -> List[str]:
    """Registers stand-ins for `matplotlib(.pyplot)` and `pycocotools(.mask)` unless the real module imports.
    Returns the names that were stubbed."""
    import importlib.util

    def missing(name):
        mod = sys.modules.get(name)
        if mod is not None:
            return False                                 # the real module, or a stand-in from an earlier call
        try:
            return importlib.util.find_spec(name) is None
        except (ImportError, ValueError):
            return True
    done = []
    if force or missing("matplotlib"):
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        inert = _Inert()

        def _plt_attr(name):                             # module-level __getattr__ (PEP 562); dunders stay absent so that
            if name.startswith("__"):                    # `inspect` and the import system see an ordinary module
                raise AttributeError(name)
            return inert
        plt.__getattr__ = _plt_attr
        mpl.pyplot = plt
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
        done.append("matplotlib.pyplot")
    if force or missing("pycocotools"):
        pkg = types.ModuleType("pycocotools")
        pkg.mask = _pycocotools_mask()
        sys.modules["pycocotools"], sys.modules["pycocotools.mask"] = pkg, pkg.mask
        done.append("pycocotools.mask")
    return done


# ------------------------------------------------------------------------------------------------ synthetic datasets
DIOR_NAMES = ('airplane', 'airport', 'baseballfield', 'basketballcourt', 'bridge', 'chimney', 'expressway-service-area',
              'expressway-toll-station', 'dam', 'golffield', 'groundtrackfield', 'harbor', 'overpass', 'ship', 'stadium',
              'storagetank', 'tenniscourt', 'trainstation', 'vehicle', 'windmill')     # parsed by name: mapping.py:52-56


def _tile(seed: int, h: int, w: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def _boxes(seed: int, n: int, h: int, w: int) -> np.ndarray:
    rng = np.random.default_rng(1_000_003 * (seed + 1))
    c = np.stack([rng.uniform(48, w - 48, n), rng.uniform(48, h - 48, n)], 1)
    wh = rng.uniform(8, 160, (n, 2))
    b = np.concatenate([c - wh / 2, c + wh / 2], 1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, h - 1)
    return np.round(b).astype(np.float32)            # DIOR stores integer pixel coordinates


def make_dior(root: str, counts: Sequence[int] = (32, 7), size: int = 800, seed: int = 100) -> List[str]:
    """<root>/dior/JPEGImages-test/<name>.jpg + Annotations/Horizontal Bounding Boxes/<name>.xml (loaddata.py:10-39)."""
    from PIL import Image
    img_dir = os.path.join(root, "dior", "JPEGImages-test")
    ann_dir = os.path.join(root, "dior", "Annotations", "Horizontal Bounding Boxes")
    os.makedirs(img_dir, exist_ok=True)
    os.makedirs(ann_dir, exist_ok=True)
    names = []
    for i, n in enumerate(counts):
        name = f"{11726 + i:05d}"
        Image.fromarray(_tile(seed + i, size, size)).save(os.path.join(img_dir, name + ".jpg"), quality=95)
        b = _boxes(seed + i, n, size, size)
        lab = np.random.default_rng(7_000_003 * (seed + i + 1)).integers(0, len(DIOR_NAMES), n)
        objs = "".join(
            f"<object><name>{DIOR_NAMES[int(l)]}</name><pose>Unspecified</pose><bndbox><xmin>{int(x0)}</xmin><ymin>{int(y0)}</ymin>"
            f"<xmax>{int(x1)}</xmax><ymax>{int(y1)}</ymax></bndbox></object>" for (x0, y0, x1, y1), l in zip(b, lab))
        with open(os.path.join(ann_dir, name + ".xml"), "w") as f:
            f.write(f"<annotation><filename>{name}.jpg</filename><size><width>{size}</width><height>{size}</height>"
                    f"<depth>3</depth></size>{objs}</annotation>")
        names.append(name)
    return names


def _rpolys(seed: int, n: int, h: int, w: int) -> np.ndarray:
    rng = np.random.default_rng(3_000_017 * (seed + 1))
    c = np.stack([rng.uniform(0.15 * w, 0.85 * w, n), rng.uniform(0.15 * h, 0.85 * h, n)], 1)
    bw, bh = rng.uniform(16, 0.25 * w, n), rng.uniform(12, 0.2 * h, n)
    th = rng.uniform(-np.pi / 2, np.pi / 2, n)
    v1 = np.stack([bw / 2 * np.cos(th), bw / 2 * np.sin(th)], 1)
    v2 = np.stack([-bh / 2 * np.sin(th), bh / 2 * np.cos(th)], 1)
    p = np.stack([c - v1 - v2, c + v1 - v2, c + v1 + v2, c - v1 + v2], 1)
    p[..., 0] = np.clip(p[..., 0], 0, w - 1)
    p[..., 1] = np.clip(p[..., 1], 0, h - 1)
    return p


def make_fair1m(root: str, counts: Sequence[int] = (25, 5), size: int = 1024, seed: int = 200, classes: int = 37) -> List[str]:
    """<root>/fair1m_1024/trainval/images/<name>.png + rbbtxts/<name>.txt, ten fields per line as `load_dota` splits
    them (loaddata.py:104-131): x1 y1 x2 y2 x3 y3 x4 y4 class_name class_index."""
    from PIL import Image
    img_dir = os.path.join(root, "fair1m_1024", "trainval", "images")
    ann_dir = os.path.join(root, "fair1m_1024", "trainval", "rbbtxts")
    os.makedirs(img_dir, exist_ok=True)
    os.makedirs(ann_dir, exist_ok=True)
    names = []
    for i, n in enumerate(counts):
        name = f"{i}__1024__0___{824 * i}"
        Image.fromarray(_tile(seed + i, size, size)).save(os.path.join(img_dir, name + ".png"))
        polys = _rpolys(seed + i, n, size, size)
        lab = np.random.default_rng(7_000_003 * (seed + i + 1)).integers(0, classes, n)
        with open(os.path.join(ann_dir, name + ".txt"), "w") as f:
            for p, l in zip(polys, lab):
                f.write(" ".join(f"{v:.1f}" for v in p.reshape(-1)) + f" cls{int(l)} {int(l)}\n")
        names.append(name)
    return names


def make_hrsc(root: str, counts: Sequence[int] = (3, 2), sizes: Sequence[Sequence[int]] = ((704, 1000), (1024, 768)), seed: int = 300) -> List[str]:
    """<root>/HRSC2016/Test/AllImages/<name>.bmp, Test/Annotations/<name>.xml (`load_hrsc`, loaddata.py:41-102) and
    FullDataSet/LandMask/<name>.png whose colours are the objects' `seg_color` (main_sam_rbox_mask_instance.py:201-208)."""
    from PIL import Image
    img_dir = os.path.join(root, "HRSC2016", "Test", "AllImages")
    ann_dir = os.path.join(root, "HRSC2016", "Test", "Annotations")
    land_dir = os.path.join(root, "HRSC2016", "FullDataSet", "LandMask")
    for d in (img_dir, ann_dir, land_dir):
        os.makedirs(d, exist_ok=True)
    names = []
    for i, (n, (h, w)) in enumerate(zip(counts, sizes)):
        name = f"1000{i:05d}"
        Image.fromarray(_tile(seed + i, h, w)).save(os.path.join(img_dir, name + ".bmp"))
        rng = np.random.default_rng(3_000_017 * (seed + i + 1))
        land = np.zeros((h, w, 3), dtype=np.uint8)
        objs = []
        for j in range(n):
            cx, cy = rng.uniform(0.25 * w, 0.75 * w), rng.uniform(0.25 * h, 0.75 * h)
            bw, bh = rng.uniform(60, 0.3 * w), rng.uniform(20, 0.12 * h)
            ang = rng.uniform(-1.4, 1.4)
            col = (10 + 40 * j, 200 - 30 * j, 90 + 20 * j)
            x0, x1, y0, y1 = int(cx - bw / 2), int(cx + bw / 2), int(cy - bh / 2), int(cy + bh / 2)
            land[max(0, y0):y1, max(0, x0):x1] = col
            objs.append(f"<HRSC_Object><Class_ID>100000001</Class_ID><box_xmin>{x0}</box_xmin><box_ymin>{y0}</box_ymin>"
                        f"<box_xmax>{x1}</box_xmax><box_ymax>{y1}</box_ymax><mbox_cx>{cx:.4f}</mbox_cx><mbox_cy>{cy:.4f}</mbox_cy>"
                        f"<mbox_w>{bw:.4f}</mbox_w><mbox_h>{bh:.4f}</mbox_h><mbox_ang>{ang:.6f}</mbox_ang>"
                        f"<seg_color>{col[0]},{col[1]},{col[2]}</seg_color></HRSC_Object>")
        Image.fromarray(land).save(os.path.join(land_dir, name + ".png"))
        with open(os.path.join(ann_dir, name + ".xml"), "w") as f:
            f.write(f"<HRSC_Image><Img_ID>{name}</Img_ID><Img_SizeWidth>{w}</Img_SizeWidth><Img_SizeHeight>{h}</Img_SizeHeight>"
                    f"<HRSC_Objects>{''.join(objs)}</HRSC_Objects></HRSC_Image>")
        names.append(name)
    return names


def make_checkpoint(root: str, variant: str = "vit_h", seed: int = 0) -> str:
    """<root>/pretrn/sam_vit_h_4b8939.pth: the seeded synthetic checkpoint in the reference's `state_dict` layout."""
    from .weights import save_checkpoint
    d = os.path.join(root, "pretrn")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, "sam_vit_h_4b8939.pth")
    if not os.path.exists(p):
        save_checkpoint(p, variant, seed)
    return p


def stage_all(stage: str, checkpoint: bool = True) -> Dict[str, str]:
    """Everything the three drivers read, under <stage>/dataset and <stage>/dw; returns the PathRedirect mapping."""
    ds, dw = os.path.join(stage, "dataset"), os.path.join(stage, "dw")
    make_dior(ds)
    make_fair1m(ds)
    make_hrsc(ds)
    os.makedirs(os.path.join(dw, "samrs", "work_dir", "hrsc", "json"), exist_ok=True)
    if checkpoint:
        make_checkpoint(dw)
    return {"/root/dataset": ds, "/root/dw": dw}


# ------------------------------------------------------------------------------------------------ running a driver
def run_driver(script: str, argv: Iterable[str] = (), package_dir: Optional[str] = None, redirect: Optional[Dict[str, str]] = None,
               stubs: bool = True) -> dict:
    """Execute an UNMODIFIED driver script as `__main__`.

    `package_dir`: directory that holds the `segment_anything` package to use (default: the samrs_b200 drop-in); it goes
    in front of the script's own directory on `sys.path`, which is the one thing `python script.py` cannot do (F8)."""
    script = os.path.abspath(script)
    package_dir = package_dir or DROPIN_PATH
    if stubs:
        install_stubs()
    saved_argv, saved_path = sys.argv, list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in _DRIVER_LOCAL_MODULES}
    sys.argv = [script] + list(argv)
    sys.path[:] = [package_dir, os.path.dirname(script)] + [p for p in saved_path if p not in (package_dir, os.path.dirname(script))]
    try:
        with (PathRedirect(redirect) if redirect else contextlib.nullcontext()):
            return runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = saved_argv
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in _DRIVER_LOCAL_MODULES]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
