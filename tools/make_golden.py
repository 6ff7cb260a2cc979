"""Generates tests/golden/* by running the REFERENCE's own modules (imported by path from /root/reference, container
only) on seeded synthetic inputs.  The goldens travel to the GPU box; /root/reference does not.

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import craft_net, ref_loader, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    cu, ip = ref["craft_utils"], ref["imgproc"]
    os.chdir("/tmp")   # the reference writes debug PNGs relative to /tmp/fragments

    # --- getDetBoxes + adjustResultCoordinates (craft_utils.py:25-98,257-274) on random blob maps
    presets = [(0.7, 0.45, 0.3), (0.4, 0.2, 0.3), (0.6, 0.3, 0.3)]
    for i, (h, w, seed) in enumerate([(96, 160, 11), (120, 200, 12), (64, 256, 13)]):
        text, link = synth.random_score_maps(seed, h, w, n_blobs=24)
        tt, lt, low = presets[i]
        boxes, labels, mapper = cu.getDetBoxes_core(text.copy(), link.copy(), tt, lt, low)
        ratio = 1.2941176470588236
        adj = cu.adjustResultCoordinates([b.copy() for b in boxes], ratio, ratio)
        np.savez_compressed(os.path.join(OUT, f"craft_post_{i}.npz"), text=text, link=link,
                            thresholds=np.array([tt, lt, low], np.float64), ratio=np.float64(ratio),
                            boxes=np.asarray(boxes, np.float32).reshape(-1, 4, 2), mapper=np.asarray(mapper, np.int32),
                            labels=labels.astype(np.int32), adj=np.asarray(adj, np.float32).reshape(-1, 4, 2))

    # --- resize_aspect_ratio + normalizeMeanVariance (imgproc.py:45-73,26-32)
    rng = np.random.default_rng(21)
    page = rng.integers(0, 256, (132, 102, 3), dtype=np.uint8)
    import cv2
    resized, ratio, heat = ip.resize_aspect_ratio(page, 79, interpolation=cv2.INTER_LINEAR, mag_ratio=1)
    norm = ip.normalizeMeanVariance(resized, mean=(0.5, 0.5, 0.5), variance=(0.5, 0.5, 0.5))
    np.savez_compressed(os.path.join(OUT, "imgproc.npz"), page=page, canvas=np.int32(79), resized=resized,
                        ratio=np.float64(ratio), norm=norm.astype(np.float32))

    # --- CRAFT.forward (craft.py:59-81) with the seeded synthetic state dict
    sd = craft_net.synth_craft_state(3, random_bn=True)
    net = ref["craft"].CRAFT(pretrained=False)
    net.load_state_dict(sd)
    net.eval()
    torch.manual_seed(5)
    x = torch.randn(1, 3, 64, 96).clamp(-1, 1)
    with torch.no_grad():
        y, feat = net(x)
    np.savez_compressed(os.path.join(OUT, "craft_net.npz"), x=x.numpy(), y=y.numpy(), feature=feat.numpy(),
                        seed=np.int32(3))

    # --- RefineNet.forward (refinenet.py:57-66) with a seeded state dict (random BN statistics)
    import importlib
    refinenet = importlib.import_module("refinenet")
    rsd = craft_net.synth_refine_state(4, random_bn=True, round_to=None)
    rnet = refinenet.RefineNet()
    rnet.load_state_dict(rsd)
    rnet.eval()
    torch.manual_seed(6)
    ry, rf = torch.randn(1, 36, 52, 2), torch.randn(1, 32, 36, 52)
    with torch.no_grad():
        rout = rnet(ry, rf)
    np.savez_compressed(os.path.join(OUT, "refine_net.npz"), y=ry.numpy(), feature=rf.numpy(), out=rout.numpy(),
                        seed=np.int32(4))

    # --- line_merge / find_line_number (line_processor.py:15-171), merge_bboxes_as_block (overlap.py:186-204)
    cases = []
    rng = np.random.default_rng(31)
    for n in (1, 7, 40, 120):
        boxes = np.stack([rng.integers(0, 2000, n), rng.integers(0, 900, n), rng.integers(5, 300, n),
                          rng.integers(0, 60, n)], 1).tolist()
        lines = np.asarray(ref["lines"].line_merge(np.zeros((8, 8, 3), np.uint8), boxes)).tolist()
        ids = [int(ref["lines"].find_line_number(lines, b)) for b in boxes]
        block = [int(v) for v in ref["overlap"].merge_bboxes_as_block(boxes)]
        cases.append(dict(boxes=boxes, lines=lines, line_ids=ids, block=block))
    with open(os.path.join(OUT, "lines.json"), "w") as f:
        json.dump(cases, f)
    print("goldens written to", OUT)


if __name__ == "__main__":
    main()


def make_result_golden():
    """OcrProcessor.recognize result assembly (marie/document/ocr_processor.py:87-267) with canned recogniser output."""
    Ref = ref_loader.load_ocr_processor()
    rng = np.random.default_rng(41)
    n = 31
    boxes = np.stack([rng.integers(0, 900, n), rng.integers(0, 500, n), rng.integers(5, 100, n),
                      rng.integers(5, 40, n)], 1).tolist()
    lines = rng.integers(-1, 6, n).tolist()
    canned = [{"confidence": float(rng.random()), "id": f"img-{k}", "text": f"W{k}"} for k in range(n)]

    class P(Ref):
        def __init__(self):
            pass

        def is_available(self):
            return True

        def recognize_from_fragments(self, frags, **kw):
            return canned

    img = np.zeros((600, 1000, 3), np.uint8)
    res, _ = P().recognize("golden", "key", img, boxes, [img[:2, :2]] * n, lines)
    with open(os.path.join(OUT, "ocr_result.json"), "w") as f:
        json.dump(dict(boxes=boxes, lines=lines, canned=canned, result=jsonable(res)), f)


def jsonable(r):
    def cv(v):
        if isinstance(v, (str, int, float)):
            return v
        return np.asarray(v).tolist()
    return {"meta": r["meta"], "words": [{k: cv(v) for k, v in w.items()} for w in r["words"]],
            "lines": [{k: cv(v) for k, v in l.items()} for l in r["lines"]]}


if __name__ == "__main__":
    make_result_golden()
