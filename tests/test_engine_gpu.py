"""The drop-in boundary end to end on the GPU: BoxProcessorCraftB200 + TrOcrProcessorB200 behind OcrEngineB200, the
batched pipeline against the page-by-page plugin loop, and every stage against the oracle fed with the same
intermediate data (score maps -> boxes -> crops -> token ids -> text)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(cuda_ctx):
    from marie_icr_b200.boxes import BoxProcessorCraftB200
    from marie_icr_b200.document import TrOcrProcessorB200
    from marie_icr_b200.engine import OcrEngineB200
    from oracle import craft_net, resample, synth, trocr
    cuda_ctx.set_dtype("fp16")
    pages = [synth.synth_page(i, height=660, width=510, scale=0.8, line_pitch=48, gap=24, margin=30)[0] for i in range(3)]
    sd = craft_net.glyph_craft_state(0)
    cfg = trocr.trocr_tiny()
    tsd = trocr.synth_trocr_state(cfg, 2)
    box = BoxProcessorCraftB200(state_dict=sd)
    frag0 = box.extract_bounding_boxes("t", "k", pages[0])[1][:3]
    cal = torch.stack([torch.from_numpy(resample.fragment_to_input(f)) for f in frag0])
    with torch.no_grad():
        trocr.calibrate_eos(tsd, cfg, eos_step=4, enc=trocr.encoder_forward(tsd, cfg, cal))
    from marie_icr_b200.bpe import SyntheticDetokenizer
    icr = TrOcrProcessorB200(state_dict=tsd, config=cfg, beam=1, max_len_b=16, pipeline=box.pipeline, detokenizer=SyntheticDetokenizer())
    eng = OcrEngineB200(box_processor=box, default_ocr_processor=icr)
    return eng, pages, sd, tsd, cfg


def _plain(results):
    out = []
    for r in results:
        out.append(dict(meta={k: (np.asarray(v).tolist() if not isinstance(v, (str, int, dict)) else v) for k, v in r["meta"].items()},
                        words=[{k: (np.asarray(v).tolist() if not isinstance(v, (str, int, float)) else v) for k, v in w.items()} for w in r["words"]],
                        lines=[{k: (np.asarray(v).tolist() if not isinstance(v, (str, int, float)) else v) for k, v in l.items()} for l in r["lines"]]))
    return out


def test_box_processor_contract_and_oracle(engine):
    from marie_icr_b200 import ops
    from marie_icr_b200.plugin_api import PSMode
    from oracle import craft_post
    eng, pages, sd, _, _ = engine
    box = eng.box_processor
    page = pages[1]
    rects, frags, line_ids, pred, lines_bboxes = box.extract_bounding_boxes("id", "key", page, PSMode.SPARSE)
    assert len(rects) == len(frags) == len(line_ids) > 10 and lines_bboxes == []
    assert all(l == -1 for l in line_ids) and isinstance(frags, list)
    # oracle post-processing on the device's own score maps
    dev = torch.from_numpy(page[None]).cuda()
    x, ratio = ops.page_preprocess(dev)
    scores = ops.craft_forward(x)
    det, _, _ = craft_post.det_boxes_cv(scores[0, 0].cpu().numpy(), scores[1, 0].cpu().numpy(), 0.7, 0.45, 0.3)
    adj = craft_post.adjust_result_coordinates([b.copy() for b in det], 1 / ratio, 1 / ratio)
    want = craft_post.boxes_to_rects(adj, page.shape[0], page.shape[1])
    same = [np.array_equal(a, b) for a, b in zip(np.asarray(det), (pred["bboxes"] / np.float32(2 / ratio)))]
    assert len(want) == len(rects)
    mism = sum(r != w for r, w in zip(rects, want))
    assert mism <= max(1, len(want) // 100), f"{mism}/{len(want)} rects differ from the oracle"
    for r, f in zip(rects, frags):
        assert np.array_equal(f, craft_post.crop_rect(page, r))
    # WORD / RAW_LINE: no detection, the whole image is one box (craft_box_processor.py:453-476)
    r2, f2, l2, p2, _ = box.extract_bounding_boxes("id", "key", page[:40, :200], PSMode.WORD)
    assert r2 == [[0, 0, 200, 40]] and l2 == [0] and p2 == {} and np.array_equal(f2[0], page[:40, :200])
    with pytest.raises(Exception, match="can't be empty"):
        box.extract_bounding_boxes("id", "key", None)
    with pytest.raises(Exception, match="not supported"):
        box.extract_bounding_boxes("id", "key", page, "bogus")


def test_icr_processor_matches_oracle(engine):
    import math
    from marie_icr_b200.bpe import SyntheticDetokenizer
    from oracle import resample, trocr
    eng, pages, _, tsd, cfg = engine
    _, frags, _, _, _ = eng.box_processor.extract_bounding_boxes("id", "key", pages[2])
    frags = frags[:24]
    res = eng.icr_processor.recognize_from_fragments(frags)
    assert [r["id"] for r in res] == [f"img-{k}" for k in range(len(frags))]
    chw = torch.stack([torch.from_numpy(resample.fragment_to_input(f)) for f in frags]).half().float()
    margins = []
    with torch.no_grad():
        hyps = trocr.generate(tsd, cfg, trocr.encoder_forward(tsd, cfg, chw), beam=1, max_len_b=16, margins=margins)
    detok = SyntheticDetokenizer()
    exact = bound = 0
    for r, h, m in zip(res, hyps, margins):
        toks, conf = h[0]["tokens"][:-1].tolist(), math.exp(h[0]["score"])
        same = r["text"] == detok.decode(toks).upper()
        if m > 0.05:                      # margin protocol (tests/test_parity_scale_gpu.py): a confident call must match
            bound += 1
            assert same, f"{r['text']} != {detok.decode(toks).upper()} with margin {m:.3f}"
        if same:
            exact += 1
            assert abs(r["confidence"] - round(round(conf, 6), 4)) <= 2e-3
    print(f"plugin vs oracle: {exact}/{len(frags)} texts identical, {bound} bound by the margin protocol")
    assert eng.icr_processor.recognize_from_fragments([]) == []


def test_engine_batched_equals_pagewise(engine):
    from marie_icr_b200.plugin_api import CoordinateFormat, PSMode
    eng, pages, _, _, _ = engine
    a = _plain(eng.extract(pages, PSMode.SPARSE, CoordinateFormat.XYXY))
    b = _plain(eng._extract_pagewise([p.copy() for p in pages], "q", "0", PSMode.SPARSE, CoordinateFormat.XYXY))
    assert len(a) == len(b) == 3
    for i, (ra, rb) in enumerate(zip(a, b)):
        assert ra["meta"] == rb["meta"] and ra["meta"]["page"] == i and ra["meta"]["format"] == "xyxy"
        assert len(ra["words"]) > 10
        assert ra["words"] == rb["words"]
        assert ra["lines"] == rb["lines"]
    # a blank page yields an empty record, not an error (ocr_processor.py:147-154)
    blank = np.full_like(pages[0], 255)
    r = eng.extract([blank, pages[0]])
    assert r[0]["words"] == [] and len(r[1]["words"]) == len(a[0]["words"])


def test_engine_crop_to_content(engine):
    """`crop_to_content=True` (ocr_engine.py:169-185): every page is cropped to its content and padded with 4 white pixels
    before detection — same result as handing the engine that padded crop directly."""
    from marie_icr_b200 import ingest
    from marie_icr_b200.plugin_api import PSMode
    eng, pages, _, _, _ = engine
    page = pages[0]
    cropped = ingest.crop_to_content(page)
    assert cropped.shape[1] < page.shape[1]
    h, w = cropped.shape[:2]
    padded = np.full((h + 8, w + 8, 3), 255, np.uint8)
    padded[4:h + 4, 4:w + 4] = cropped
    a = _plain(eng.extract([page], PSMode.SPARSE, crop_to_content=True))
    b = _plain(eng._extract_pagewise([padded], "q", "0", PSMode.SPARSE, eng_format(eng)))
    assert a[0]["meta"]["imageSize"] == {"width": w + 8, "height": h + 8}
    assert len(a[0]["words"]) > 10 and a[0]["words"] == b[0]["words"] and a[0]["lines"] == b[0]["lines"]


def eng_format(eng):
    from marie_icr_b200.plugin_api import CoordinateFormat
    return CoordinateFormat.XYWH


def test_engine_regions(engine):
    """Region / field extraction (ocr_engine.py:223-414): per-region PSM, 4 px padding, one recognize() per page, the
    {"regions", "extended"} payload; results equal recognising the padded region directly."""
    from marie_icr_b200.plugin_api import PSMode
    eng, pages, _, _, _ = engine
    page = pages[0]
    rects, _, _, _, _ = eng.box_processor.extract_bounding_boxes("id", "key", page)
    picks = [rects[3], rects[7], rects[12]]
    regions = [{"id": f"r{k}", "pageIndex": 0, "x": int(x), "y": int(y), "w": int(w), "h": int(h), "mode": "raw_line"}
               for k, (x, y, w, h) in enumerate(picks)]
    res = eng.extract([page], PSMode.SPARSE, regions=regions)
    assert set(res) == {"regions", "extended"} and len(res["regions"]) == 3 and len(res["extended"]) == 1
    direct = []
    for (x, y, w, h) in picks:
        ov = np.full((h + 8, w + 8, 3), 255, np.uint8)
        ov[4:h + 4, 4:w + 4] = page[y:y + h, x:x + w]
        direct.append(eng.icr_processor.recognize_from_fragments([ov])[0])
    # all region boxes are [0, 0, w+8, h+8]: the x-sort is a no-op, so ids map in order
    for r, d, reg in zip(res["regions"], direct, regions):
        assert r["id"] == reg["id"] and r["text"] == d["text"] and abs(r["confidence"] - round(d["confidence"], 3)) < 1e-9
    # second call hits the box cache and returns the same payload
    assert eng.extract([page], PSMode.SPARSE, regions=regions)["regions"] == res["regions"]
    # a page whose regions are ALL skipped (zero size / out of bounds) leaves nothing to unpack: the reference raises
    # ValueError at ocr_engine.py:343-351, and so does the mirror
    with pytest.raises(ValueError):
        eng.extract([page], regions=[{"id": "z", "pageIndex": 0, "x": 0, "y": 0, "w": 0, "h": 5}])
    # ... while a skipped region next to a valid one yields empty results for the whole page (ids are recorded before
    # the checks, so the word count no longer matches; ocr_engine.py:262-282,383-394)
    mixed = eng.extract([page], regions=[{"id": "z", "pageIndex": 0, "x": 0, "y": 0, "w": 0, "h": 5}, regions[0]])
    assert [r["id"] for r in mixed["regions"]] == ["z", "z", "r0"] and all(r["text"] == "" for r in mixed["regions"])
    with pytest.raises(Exception, match="Required key missing"):
        eng.extract([page], regions=[{"id": "q", "pageIndex": 0, "x": 1}])


def test_box_processor_poly_option(engine):
    """poly=True (get_prediction's `poly`, craft_box_processor.py:76-135; off in every preset): prediction_result["polys"]
    comes from the polygon refinement over the DEVICE's label map / boxes / mapper and equals the same routine fed with
    the cv2-based oracle's post-processing of the device's score maps; boxes, rects and fragments are unchanged."""
    import cv2
    from marie_icr_b200 import ops
    from marie_icr_b200.boxes import BoxProcessorCraftB200
    from marie_icr_b200.plugin_api import PSMode
    from marie_icr_b200.polys import adjust_polys, get_poly_core
    from oracle import craft_post
    eng, pages, _, _, _ = engine
    page = pages[2].copy()
    xs = np.arange(60, 440, 2)
    for y0, amp in ((250, 16.0), (420, 22.0)):                       # two curved 'words' of ink
        pts = np.stack([xs, y0 + amp * np.sin((xs - 60) * 2 * np.pi / 380.0)], 1).astype(np.int32).reshape(-1, 1, 2)
        cv2.rectangle(page, (40, y0 - 40), (470, y0 + 40), (255, 255, 255), -1)
        cv2.polylines(page, [pts], False, (0, 0, 0), thickness=14)
    plain = eng.box_processor
    curved = BoxProcessorCraftB200(pipeline=plain.pipeline, poly=True)
    r0, f0, l0, p0, _ = plain.extract_bounding_boxes("id", "key", page, PSMode.SPARSE)
    r1, f1, l1, p1, _ = curved.extract_bounding_boxes("id", "key", page, PSMode.SPARSE)
    assert r0 == r1 and l0 == l1 and np.array_equal(p0["bboxes"], p1["bboxes"])
    assert all(np.array_equal(a, b) for a, b in zip(f0, f1))
    assert all(np.array_equal(a, b) for a, b in zip(p0["polys"], p0["bboxes"]))
    # the same refinement on the oracle's labels / boxes of the device's own score maps
    dev = torch.from_numpy(page[None]).cuda()
    x, ratio = ops.page_preprocess(dev)
    scores = ops.craft_forward(x)
    det, labels, mapper = craft_post.det_boxes_cv(scores[0, 0].cpu().numpy(), scores[1, 0].cpu().numpy(), 0.7, 0.45, 0.3)
    want = adjust_polys(get_poly_core([np.asarray(b, np.float32) for b in det], labels.astype(np.int32), mapper),
                        [b for b in p1["bboxes"]], 1 / ratio, 1 / ratio)
    assert len(want) == len(p1["polys"]) == len(r1)
    n14 = 0
    for a, b in zip(p1["polys"], want):
        assert a.shape == b.shape and np.array_equal(a, b)
        n14 += a.shape[0] == 14
    assert n14 >= 1, "no curved component produced a polygon"
