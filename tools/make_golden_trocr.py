"""Generates the at-scale TrOCR goldens (tests/golden/trocr_scale_*.npz) and the end-to-end page golden
(tests/golden/e2e_page_*.json) by running the CPU ORACLE chain here, in the container (minutes of CPU work that the
GPU box then does not have to repeat).  The crops are the bench's own: page 0 of the synthetic letter stream through
oracle-K1 -> oracle-CRAFT (glyph-path weights) -> oracle getDetBoxes (SPARSE preset) -> rects -> crops -> Pillow-exact
resample, with the bench's seeded TrOCR weights rounded once to the device's 16-bit type.

    python tools/make_golden_trocr.py scale base fp16 512      # ~10 min on 8 cores
    python tools/make_golden_trocr.py scale base bf16 512
    python tools/make_golden_trocr.py scale large fp16 64
    python tools/make_golden_trocr.py e2e                      # small page -> reference OcrProcessor.recognize JSON

Per crop the file holds the oracle's greedy and beam-5 top hypothesis (token ids incl. the final EOS, length-normalised
score) and the margin of the closest call the search made (oracle/trocr.generate(margins=...)).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import craft_net, craft_post, resample, trocr  # noqa: E402
from synthetic import pages as synth  # noqa: E402
from synthetic import weights as sw  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DT = {"fp16": torch.float16, "bf16": torch.bfloat16}
SPARSE = (0.7, 0.45, 0.3)


def oracle_detect(page, dt, preset=SPARSE):
    """oracle K1 -> CRAFT -> getDetBoxes -> adjustResultCoordinates -> rects (craft_box_processor.py:76-146,499-521)."""
    x, ratio = resample.craft_input(page)
    xin = torch.from_numpy(np.ascontiguousarray(x)).to(dt).float().permute(2, 0, 1)[None]
    with torch.no_grad():
        y, _ = craft_net.craft_forward(sw.glyph_craft_state(0), xin)
    det, _, _ = craft_post.det_boxes_cv(y[0, ..., 0].numpy(), y[0, ..., 1].numpy(), *preset)
    adj = craft_post.adjust_result_coordinates([b.copy() for b in det], 1 / ratio, 1 / ratio)
    rects = craft_post.boxes_to_rects(adj, page.shape[0], page.shape[1])
    return np.asarray(rects, np.int32).reshape(-1, 4)


def trocr_weights(model, dt):
    cfg = sw.trocr_base() if model == "base" else sw.trocr_large()
    sd = sw.apply_eos_row(sw.synth_trocr_state(cfg, 0, round_to=dt), f"trocr_{model}_seed0", round_to=dt)
    return sd, cfg


def crops_to_input(page, rects, dt):
    """the exact 16-bit values the device network sees (K9 is bit-exact against this, tests/test_imgproc_gpu.py)"""
    return torch.stack([torch.from_numpy(np.ascontiguousarray(resample.fragment_to_input(craft_post.crop_rect(page, r)))).to(dt).float()
                        for r in rects])


def search(sd, cfg, enc, beam, max_len_b, finalists=None):
    margins = []
    with torch.no_grad():
        hyps = trocr.generate(sd, cfg, enc, beam=beam, max_len_b=max_len_b, margins=margins)
    if finalists is not None:
        finalists.extend(hyps)                # all `beam` finalised hypotheses per crop, best first
    return [h[0] for h in hyps], margins


def pad_tokens(hyps, width):
    out = np.full((len(hyps), width), trocr.PAD, np.int32)
    for i, h in enumerate(hyps):
        t = h["tokens"].tolist()
        out[i, :len(t)] = t[:width]
    return out


def make_scale(model, dtype, n, chunk=64, max_len_b=200):
    dt = DT[dtype]
    t0 = time.time()
    pages, rects, pidx = [], [], []
    while sum(len(r) for r in rects) < n:                    # crops of page 0, then page 1, ... until n
        page, _ = synth.synth_page(len(pages))
        r = oracle_detect(page, dt)[:n - sum(len(x) for x in rects)]
        print(f"page {len(pages)}: {len(r)} rects ({time.time() - t0:.1f} s)", flush=True)
        pidx += [len(pages)] * len(r)
        pages.append(page)
        rects.append(r)
    rects, pidx = np.concatenate(rects), np.asarray(pidx, np.int32)
    sd, cfg = trocr_weights(model, dt)
    res = {}
    for i0 in range(0, len(rects), chunk):
        chw = torch.cat([crops_to_input(pages[p], rects[i0:i0 + chunk][pidx[i0:i0 + chunk] == p], dt)
                         for p in sorted(set(pidx[i0:i0 + chunk].tolist()))])
        with torch.no_grad():
            enc = trocr.encoder_forward(sd, cfg, chw)
        for tag, w in (("", sd),):
            for name, beam in (("greedy", 1), ("beam5", 5)):
                h, m = search(w, cfg, enc, beam, max_len_b, finalists=res.setdefault(tag + name + "_all", []) if beam > 1 else None)
                res.setdefault(tag + name, []).extend(h)
                res.setdefault(tag + name + "_margin", []).extend(m)
        print(f"  {i0 + len(chw)}/{len(rects)} crops ({time.time() - t0:.1f} s)", flush=True)
    out = dict(page_index=pidx, rects=rects, max_len_b=np.int32(max_len_b))
    for key in ("greedy", "beam5"):
        hyps = res[key]
        out[key + "_tokens"] = pad_tokens(hyps, max(len(h["tokens"]) for h in hyps))
        out[key + "_len"] = np.array([len(h["tokens"]) for h in hyps], np.int32)
        out[key + "_score"] = np.array([h["score"] for h in hyps], np.float32)
        out[key + "_margin"] = np.array(res[key + "_margin"], np.float32)
        if key + "_all" in res:
            # every finalised hypothesis of the beam search (the parity test explains rank flips among near-tied finalists)
            alls = res[key + "_all"]
            width = max(len(h["tokens"]) for hs in alls for h in hs)
            ft = np.full((len(alls), 5, width), trocr.PAD, np.int32)
            fs = np.full((len(alls), 5), -np.inf, np.float32)
            for i, hs in enumerate(alls):
                for j, h in enumerate(hs[:5]):
                    ft[i, j, :len(h["tokens"])] = h["tokens"].tolist()
                    fs[i, j] = h["score"]
            out[key + "_finalists"], out[key + "_finalist_scores"] = ft, fs
        print(key, "margin > 0.05:", int((out[key + "_margin"] > 0.05).sum()), "of", len(hyps), "mean length",
              float(out[key + "_len"].mean()))
    path = os.path.join(OUT, f"trocr_scale_{model}_{dtype}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def make_e2e(dtype="fp16", beam=1):
    """A small page through the whole oracle chain, assembled into the reference's page record by the REFERENCE's own
    OcrProcessor.recognize (marie/document/ocr_processor.py:87-267, loaded by path) and finished like
    __process_extract_fullpage (marie/ocr/ocr_engine.py:200-217)."""
    from oracle import ref_loader
    sys.path.insert(0, os.path.join(ROOT, "marie-icr_b200"))
    from bpe import SyntheticDetokenizer       # the deterministic id -> text map both sides use with random weights
    import math
    dt = DT[dtype]
    geom = dict(height=660, width=510, scale=0.8, line_pitch=48, gap=24, margin=30)
    page, _ = synth.synth_page(0, **geom)
    rects = oracle_detect(page, dt)
    sd, cfg = trocr_weights("base", dt)
    chw = crops_to_input(page, rects, dt)
    with torch.no_grad():
        enc = trocr.encoder_forward(sd, cfg, chw)
    hyps, margins = search(sd, cfg, enc, beam, 200)
    detok = SyntheticDetokenizer()
    canned = []
    for k, h in enumerate(hyps):
        toks = h["tokens"].tolist()
        conf = round(round(math.exp(float(np.float32(h["score"]))), 6), 4)          # trocr_ocr_processor.py:159-160,341
        canned.append({"confidence": conf, "id": f"img-{k}", "text": detok.decode(toks).upper()})
    Ref = ref_loader.load_ocr_processor()

    class P(Ref):
        def __init__(self):
            pass

        def is_available(self):
            return True

        def recognize_from_fragments(self, frags, **kw):
            return canned

    boxes = rects.tolist()
    frags = [craft_post.crop_rect(page, r) for r in rects]
    lines = [-1] * len(boxes)                                # find_line_number([], box) (line_processor.py:21-45)
    result, _ = P().recognize("golden", "key", page, boxes, frags, lines)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden import jsonable
    result = jsonable(result)
    result["meta"].update(page=0, lines=lines, lines_bboxes=[], format="xywh")     # ocr_engine.py:200-217
    path = os.path.join(OUT, f"e2e_page_{dtype}_beam{beam}.json")
    with open(path, "w") as f:
        json.dump(dict(page_geometry=geom, dtype=dtype, beam=beam, rects=boxes, margins=[float(m) for m in margins],
                       tokens=[h["tokens"].tolist() for h in hyps], result=result), f)
    print("wrote", path, len(boxes), "words")


def make_regions(dtype="fp16"):
    """Region / field extraction golden: the REFERENCE's own __process_extract_regions (marie/ocr/ocr_engine.py:223-414,
    executed from its source) over oracle-backed processors — detection = oracle chain + the reference's own per-box loop
    (craft_box_processor.py:499-537), recognition = oracle TrOCR behind the reference's OcrProcessor.recognize."""
    import math
    from oracle import ref_loader
    sys.path.insert(0, os.path.join(ROOT, "marie-icr_b200"))
    from bpe import SyntheticDetokenizer
    from plugin_api import PSMode
    dt = DT[dtype]
    geom = dict(height=660, width=510, scale=0.8, line_pitch=48, gap=24, margin=30)
    frames = [synth.synth_page(i, **geom)[0] for i in range(2)]
    rects0 = oracle_detect(frames[0], dt)
    sd, cfg = trocr_weights("base", dt)
    detok = SyntheticDetokenizer()
    box_loop = ref_loader.load_box_loop()
    margins_log = []

    class Box:
        def extract_bounding_boxes(self, _id, key, img, psm=PSMode.SPARSE):
            if psm in (PSMode.WORD, PSMode.RAW_LINE):                          # craft_box_processor.py:453-476
                return [[0, 0, img.shape[1], img.shape[0]]], [img.copy()], [0], dict(), []
            preset = {"sparse": SPARSE, "line": (0.4, 0.2, 0.3), "multiline": (0.6, 0.3, 0.3)}[psm.value]
            x, ratio = resample.craft_input(img)
            xin = torch.from_numpy(np.ascontiguousarray(x)).to(dt).float().permute(2, 0, 1)[None]
            with torch.no_grad():
                y, _ = craft_net.craft_forward(sw.glyph_craft_state(0), xin)
            det, _, _ = craft_post.det_boxes_cv(y[0, ..., 0].numpy(), y[0, ..., 1].numpy(), *preset)
            adj = craft_post.adjust_result_coordinates([b.copy() for b in det], 1 / ratio, 1 / ratio)
            rects, frags, lines = box_loop(img, list(adj), [], "/tmp/fragments/regions")
            return [list(map(int, r)) for r in rects], frags, lines, {"bboxes": adj}, []

    Ref = ref_loader.load_ocr_processor()

    class Icr(Ref):
        def __init__(self):
            pass

        def is_available(self):
            return True

        def recognize_from_fragments(self, frags, **kw):
            chw = torch.stack([torch.from_numpy(np.ascontiguousarray(resample.fragment_to_input(f))).to(dt).float() for f in frags])
            with torch.no_grad():
                hyps, margins = search(sd, cfg, trocr.encoder_forward(sd, cfg, chw), 1, 200)
            margins_log.append([float(m) for m in margins])
            return [{"confidence": round(round(math.exp(float(np.float32(h["score"]))), 6), 4), "id": f"img-{k}",
                     "text": detok.decode(h["tokens"].tolist()).upper()} for k, h in enumerate(hyps)]

    regions = []
    for k in (0, 3, 7, 12, 20, 31):                                        # single words as WORD-mode fields
        x, y, w, h = (int(v) for v in rects0[k])
        regions.append({"id": f"w{k}", "pageIndex": 0, "x": x, "y": y, "w": w, "h": h, "mode": "word"})
    regions.append({"id": "line", "pageIndex": 0, "x": 10, "y": 20, "w": 480, "h": 90, "mode": "raw_line"})
    # page 1: detector-mode regions (two of one shape -> one batched detector pass on the device) and the skipped-region quirk
    regions.append({"id": "s0", "pageIndex": 1, "x": 8, "y": 24, "w": 200, "h": 64, "mode": "sparse"})
    regions.append({"id": "s1", "pageIndex": 1, "x": 8, "y": 120, "w": 200, "h": 64, "mode": "sparse"})
    regions.append({"id": "zero", "pageIndex": 1, "x": 5, "y": 5, "w": 0, "h": 10})
    regions.append({"id": "oob", "pageIndex": 1, "x": 400, "y": 600, "w": 300, "h": 100})
    result = ref_loader.load_region_loop()([f.copy() for f in frames], regions, PSMode.SPARSE, Box(), Icr(), PSMode)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden import jsonable
    out = dict(page_geometry=geom, dtype=dtype, regions=regions, margins=margins_log,
               result={"regions": result["regions"], "extended": [jsonable(e) for e in result["extended"]]})
    path = os.path.join(OUT, f"regions_{dtype}.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, result["regions"])


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    os.chdir("/tmp")
    if sys.argv[1] == "scale":
        make_scale(sys.argv[2], sys.argv[3], int(sys.argv[4]))
    elif sys.argv[1] == "regions":
        make_regions(*(sys.argv[2:3] or ["fp16"]))
    else:
        make_e2e(*(sys.argv[2:3] or ["fp16"]), beam=int(sys.argv[3]) if len(sys.argv) > 3 else 1)
