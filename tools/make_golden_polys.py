"""tests/golden/polys.npz: getPoly_core (marie/models/craft/craft_utils.py:101-254) executed from the reference's own file on
synthetic curved words — score maps of thick sine / arc strokes, components and boxes from the reference's getDetBoxes_core.
Container only (/root/reference); the golden travels.

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_polys.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def curved_maps(seed, h=320, w=640, n_words=10):
    """Text map of curved 'words' (thick polylines along sines / arcs of random amplitude, some straight); link map 0."""
    rng = np.random.default_rng(seed)
    canvas = np.zeros((h, w), np.float32)
    rows = max(1, n_words // 2)
    for k in range(n_words):
        x0 = 20 + (k % 2) * (w // 2) + int(rng.integers(0, 30))
        length = int(rng.integers(w // 5, w // 2 - 60))
        y0 = 30 + (k // 2) * ((h - 40) // rows) + int(rng.integers(0, 8))
        amp = float(rng.choice([0.0, 6.0, 10.0, 14.0, 18.0]))
        period = float(rng.uniform(0.8, 1.6)) * length
        xs = np.arange(x0, x0 + length, 2)
        ys = y0 + amp * np.sin((xs - x0) * 2 * np.pi / period + rng.uniform(0, np.pi))
        pts = np.stack([xs, ys], 1).astype(np.int32).reshape(-1, 1, 2)
        cv2.polylines(canvas, [pts], False, 1.0, thickness=int(rng.integers(7, 12)))
    text = np.clip(cv2.GaussianBlur(canvas, (5, 5), 1.2), 0, 1).astype(np.float32)
    return text, np.zeros_like(text)


def main():
    ref = ref_loader.load()
    cu = ref["craft_utils"]
    os.chdir("/tmp")
    out = {}
    n_poly = 0
    for i, seed in enumerate([5, 6, 7, 8]):
        text, link = curved_maps(seed)
        boxes, labels, mapper = cu.getDetBoxes_core(text, link, 0.7, 0.4, 0.4)
        polys = cu.getPoly_core(boxes, labels, mapper, link)
        out[f"boxes{i}"] = np.asarray(boxes, np.float32).reshape(-1, 4, 2)
        out[f"labels{i}"] = labels.astype(np.int32)
        out[f"mapper{i}"] = np.asarray(mapper, np.int32)
        out[f"has{i}"] = np.asarray([p is not None for p in polys])
        out[f"polys{i}"] = np.stack([p for p in polys if p is not None]) if any(p is not None for p in polys) else np.zeros((0, 14, 2))
        n_poly += int(out[f"has{i}"].sum())
        print(f"map {i}: {len(boxes)} boxes, {int(out[f'has{i}'].sum())} polygons")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "polys.npz"), **out)
    print("polygons in the golden:", n_poly)


if __name__ == "__main__":
    main()
