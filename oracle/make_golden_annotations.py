"""Writes synthetic DOTA / DIOR / HRSC annotation files and records what the REFERENCE's loaders
(`Generate Dataset/loaddata.py`) return for them -> tests/golden/annotations/ (files + expected.npz).
Run in the build container, where /root/reference exists."""
import os, sys
import numpy as np

REF = "/root/reference/Generate Dataset"
sys.path.insert(0, REF)
import loaddata  # noqa: E402
from mapping import DIOR  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "annotations")


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(5)
    # ---- DOTA
    lines = []
    for _ in range(6):
        c, w, h, t = rng.uniform(100, 900, 2), rng.uniform(10, 150), rng.uniform(10, 150), rng.uniform(-1.5, 1.5)
        v1, v2 = np.array([w / 2 * np.cos(t), w / 2 * np.sin(t)]), np.array([-h / 2 * np.sin(t), h / 2 * np.cos(t)])
        p = [c - v1 - v2, c + v1 - v2, c + v1 + v2, c - v1 + v2]
        lines.append(" ".join(f"{v:.1f}" for q in p for v in q) + f" ship {int(rng.integers(0, 18))}")
    open(os.path.join(OUT, "P0001.txt"), "w").write("\n".join(lines) + "\n")
    open(os.path.join(OUT, "P0002.txt"), "w").write("")                      # empty image -> error = 1
    # ---- DIOR
    objs = []
    for i in range(5):
        x0, y0 = rng.integers(0, 600, 2); w, h = rng.integers(5, 190, 2)
        name = DIOR[int(rng.integers(0, len(DIOR)))]
        name = name.upper() if i == 1 else name                               # the loader lower-cases
        tag = "robndbox" if i == 3 else "bndbox"                              # some files use robndbox
        objs.append(f"<object><name>{name}</name><{tag}><xmin>{x0}</xmin><ymin>{y0}</ymin><xmax>{x0 + w}</xmax><ymax>{y0 + h}</ymax></{tag}></object>")
    open(os.path.join(OUT, "00011.xml"), "w").write("<annotation>" + "".join(objs) + "</annotation>")
    # ---- HRSC
    objs = []
    for i in range(5):
        cx, cy = rng.uniform(100, 900, 2); w, h = rng.uniform(20, 300), rng.uniform(8, 60); ang = rng.uniform(-1.57, 1.57)
        colour = "12,200,7" if i != 2 else "12,200"                           # malformed colour -> error = 1, colour 0,0,0
        objs.append(f"<HRSC_Object><box_xmin>{cx - w / 2:.3f}</box_xmin><box_ymin>{cy - h / 2:.3f}</box_ymin><box_xmax>{cx + w / 2:.3f}</box_xmax>"
                    f"<box_ymax>{cy + h / 2:.3f}</box_ymax><mbox_cx>{cx:.4f}</mbox_cx><mbox_cy>{cy:.4f}</mbox_cy><mbox_w>{w:.4f}</mbox_w>"
                    f"<mbox_h>{h:.4f}</mbox_h><mbox_ang>{ang:.6f}</mbox_ang><seg_color>{colour}</seg_color></HRSC_Object>")
    open(os.path.join(OUT, "100000001.xml"), "w").write("<HRSC_Image><HRSC_Objects>" + "".join(objs) + "</HRSC_Objects></HRSC_Image>")

    blob = {}
    hb, rb, pt, lb, err = loaddata.load_dota("P0001", OUT)
    blob.update(dota_hboxes=np.stack(hb), dota_rboxes=np.stack(rb), dota_points=np.stack(pt), dota_labels=np.asarray(lb), dota_error=err)
    blob["dota_empty_error"] = loaddata.load_dota("P0002", OUT)[-1]
    hb, pt, lb, err = loaddata.load_dior("00011", OUT)
    blob.update(dior_hboxes=np.stack(hb), dior_points=np.stack(pt), dior_labels=np.asarray(lb), dior_error=err)
    hb, rb, co, pt, lb, err = loaddata.load_hrsc("100000001", OUT)
    blob.update(hrsc_hboxes=np.stack(hb), hrsc_rboxes=np.stack(rb), hrsc_colors=np.stack(co), hrsc_points=np.stack(pt),
                hrsc_labels=np.asarray(lb), hrsc_error=err)
    blob["dior_classes"] = np.asarray(DIOR)
    np.savez(os.path.join(OUT, "expected.npz"), **blob)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in blob.items()})


if __name__ == "__main__":
    main()
