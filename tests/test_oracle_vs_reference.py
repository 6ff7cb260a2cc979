"""Pins the oracle restatements against the REFERENCE's own modules imported by path from /root/reference
(oracle/ref_loader.py).  Runs in the build container only; on the GPU box the reference tree is absent and the
committed goldens (tests/test_golden.py) carry the same evidence."""
import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


@pytest.mark.parametrize("seed,preset", [(0, (0.7, 0.45, 0.3)), (1, (0.4, 0.2, 0.3)), (2, (0.6, 0.3, 0.3))])
def test_get_det_boxes(ref, seed, preset):
    from oracle import craft_post, synth
    text, link = synth.random_score_maps(seed, 110, 180, n_blobs=28)
    boxes, labels, mapper = ref["craft_utils"].getDetBoxes_core(text, link, *preset)
    det, lab, mp = craft_post.det_boxes_cv(text, link, *preset)
    assert np.array_equal(labels, lab) and mapper == mp
    assert np.array_equal(np.asarray(boxes), np.asarray(det))
    a = ref["craft_utils"].adjustResultCoordinates([b.copy() for b in boxes], 1.2941, 1.2941)
    b = craft_post.adjust_result_coordinates([b.copy() for b in det], 1.2941, 1.2941)
    assert np.array_equal(np.asarray(a), np.asarray(b))


def test_glyph_page_maps(ref):
    """A realistic (small) text page: ~40 word components."""
    from oracle import craft_post, synth
    page, _ = synth.synth_page(3, height=600, width=800, scale=0.9, line_pitch=52, gap=30, margin=40)
    text, link = synth.score_maps_from_page(page, 232, 310)
    boxes, labels, mapper = ref["craft_utils"].getDetBoxes_core(text, link, 0.7, 0.45, 0.3)
    det, lab, mp = craft_post.det_boxes_cv(text, link, 0.7, 0.45, 0.3)
    assert len(boxes) > 20 and mapper == mp and np.array_equal(labels, lab)
    assert np.array_equal(np.asarray(boxes), np.asarray(det))


def test_imgproc(ref):
    import cv2
    from oracle import resample
    rng = np.random.default_rng(9)
    page = rng.integers(0, 256, (165, 128, 3), dtype=np.uint8)
    resized, ratio, _ = ref["imgproc"].resize_aspect_ratio(page, 128, interpolation=cv2.INTER_LINEAR, mag_ratio=1)
    norm = ref["imgproc"].normalizeMeanVariance(resized, mean=(0.5, 0.5, 0.5), variance=(0.5, 0.5, 0.5))
    out, r = resample.craft_input(page)
    assert r == ratio and np.array_equal(out, norm)


def test_craft_forward(ref):
    from oracle import craft_net
    sd = craft_net.synth_craft_state(1, random_bn=True, bf16_round=False)
    net = ref["craft"].CRAFT(pretrained=False)
    net.load_state_dict(sd)
    net.eval()
    torch.manual_seed(2)
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        y, f = net(x)
        y2, f2 = craft_net.craft_forward(sd, x)
    assert torch.equal(y, y2) and torch.equal(f, f2)


def test_line_merge(ref):
    from oracle import lines
    rng = np.random.default_rng(4)
    for n in (1, 5, 33, 90):
        boxes = np.stack([rng.integers(0, 2000, n), rng.integers(0, 700, n), rng.integers(5, 300, n),
                          rng.integers(0, 60, n)], 1).tolist()
        a = np.asarray(ref["lines"].line_merge(np.zeros((8, 8, 3), np.uint8), boxes))
        b = np.asarray(lines.line_merge(boxes))
        assert np.array_equal(a, b)
        for box in boxes[:10]:
            assert ref["lines"].find_line_number(a.tolist(), box) == lines.find_line_number(b.tolist(), box)
    assert ref["lines"].find_line_number([], [1, 2, 3, 4]) == -1 == lines.find_line_number([], [1, 2, 3, 4])


def test_refine_forward(ref):
    """RefineNet.forward (marie/models/craft/refinenet.py:57-66) vs oracle/craft_net.refine_forward, same state dict."""
    import importlib
    from oracle import craft_net
    refinenet = importlib.import_module("refinenet")            # marie/models/craft is on sys.path (ref_loader.load)
    sd = craft_net.synth_refine_state(1, random_bn=True, round_to=None)
    net = refinenet.RefineNet()
    net.load_state_dict(sd)
    net.eval()
    torch.manual_seed(3)
    y, f = torch.randn(2, 40, 56, 2), torch.randn(2, 32, 40, 56)
    with torch.no_grad():
        a = net(y, f)
        b = craft_net.refine_forward(sd, y, f)
    assert a.shape == (2, 40, 56, 1) and torch.equal(a, b)


def test_line_closing_restatement():
    """The bit-level closing used to check the device kernels equals cv2.morphologyEx(MORPH_CLOSE, 3x3) — the call the
    reference's line branch makes (marie/boxes/craft_box_processor.py:170-174)."""
    import cv2
    from oracle import craft_post
    rng = np.random.default_rng(0)
    for shape, p in (((50, 70), 0.6), ((33, 31), 0.3), ((7, 100), 0.8), ((1, 5), 0.5)):
        m = rng.random(shape) > p
        cvc = cv2.morphologyEx((m * 255).astype(np.float32), cv2.MORPH_CLOSE,
                               cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3)), iterations=1)
        assert np.array_equal(cvc > 0, craft_post.close3x3_restated(m))


def test_crop_to_content():
    """ingest.crop_to_content (host mirror of marie/utils/image_utils.py:190-251, the engine's optional pre-step) against
    the reference function itself."""
    from marie_icr_b200 import ingest
    from synthetic import pages as synth
    ref_iu = ref_loader.load_image_utils()
    page, _ = synth.synth_page(3, height=600, width=800, scale=0.9, line_pitch=52, gap=30, margin=120)
    gray = page[..., 0].copy()
    blank = np.full((50, 60, 3), 255, np.uint8)
    for frame in (page, gray, blank):
        for aware in (True, False):
            a = ref_iu.crop_to_content(frame.copy(), content_aware=aware)
            b = ingest.crop_to_content(frame.copy(), content_aware=aware)
            assert a.shape == b.shape and np.array_equal(a, b)
    assert ingest.crop_to_content(page).shape[1] < page.shape[1]


# ------------------------------------------------------------------------------------------- round 2 pins
def _vit_state(sd):
    return {k[len("encoder.deit."):]: v for k, v in sd.items() if k.startswith("encoder.deit.")}


def test_trocr_encoder_against_reference_forward_features():
    """oracle.encoder_forward == the reference's own AdaptedVisionTransformer.forward_features
    (marie/models/unilm/trocr/deit.py:105-146) running on the reference-held ViT blocks
    (marie/boxes/dit/ditod/deit.py:44-167) — tiny widths, strict state-dict load (pins the key names too)."""
    from functools import partial
    from oracle import trocr
    deit = ref_loader.load_trocr_deit()
    cfg = trocr.trocr_tiny()
    sd = trocr.synth_trocr_state(cfg, 5, round_to=None)
    net = deit.AdaptedVisionTransformer(img_size=384, patch_size=16, embed_dim=cfg.enc_dim, depth=cfg.enc_layers,
                                        num_heads=cfg.enc_heads, mlp_ratio=cfg.enc_ffn / cfg.enc_dim, qkv_bias=False,
                                        norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), ape=0, mask_ratio=0.0)
    net.load_state_dict(_vit_state(sd), strict=True)
    net.eval()
    torch.manual_seed(1)
    imgs = torch.rand(3, 3, 384, 384) * 2 - 1
    with torch.no_grad():
        want, _ = net.forward_features(imgs)
        got = trocr.encoder_forward(sd, cfg, imgs)
    assert want.shape == got.shape == (3, 577, cfg.enc_dim)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), float((got - want).abs().max())


def test_trocr_encoder_base_factory_geometry():
    """beit_base_patch16_384 (deit.py:323-329) as the reference builds it: one crop at full TrOCR-base geometry."""
    from oracle import trocr
    deit = ref_loader.load_trocr_deit()
    cfg = trocr.trocr_base()
    sd = trocr.synth_trocr_state(cfg, 1, round_to=None)
    net = deit.beit_base_patch16_384(ape=0, mask_ratio=0.0)
    net.load_state_dict(_vit_state(sd), strict=True)
    net.eval()
    torch.manual_seed(2)
    imgs = torch.rand(1, 3, 384, 384) * 2 - 1
    with torch.no_grad():
        want, _ = net.forward_features(imgs)
        got = trocr.encoder_forward(sd, cfg, imgs)
    rel = float((got - want).norm() / want.norm())
    assert rel < 1e-5, rel


def test_box_loop_against_reference_source():
    """oracle boxes_to_rects / crop_rect == the reference's own per-box loop and crop_poly_low
    (marie/boxes/craft_box_processor.py:42-73,499-537, executed from its source): random quadrilaterals incl. boxes
    clipped by every page border, degenerate and negative coordinates."""
    import tempfile
    from oracle import craft_post
    run = ref_loader.load_box_loop()
    rng = np.random.default_rng(5)
    H, W = 300, 420
    image = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    boxes = []
    for _ in range(300):
        cx, cy = rng.uniform(-5, W + 5), rng.uniform(-5, H + 5)
        w, h, a = rng.uniform(1, 90), rng.uniform(1, 40), rng.uniform(0, np.pi)
        c, s = np.cos(a), np.sin(a)
        pts = np.array([[-w, -h], [w, -h], [w, h], [-w, h]], np.float32) / 2 @ np.array([[c, s], [-s, c]], np.float32)
        boxes.append((pts + np.array([cx, cy], np.float32)).astype(np.float32))
    for b in ([[0, 0], [W, 0], [W, H], [0, H]], [[W - 1, H - 1]] * 4, [[3.9, 3.9], [4.1, 3.9], [4.1, 4.1], [3.9, 4.1]]):
        boxes.append(np.array(b, np.float32))
    boxes = [b for b in boxes if b[:, 0].max() >= 0 and b[:, 1].max() >= 0 and b[:, 0].min() < W and b[:, 1].min() < H]
    with tempfile.TemporaryDirectory() as d:
        rects, frags, line_ids = run(image, boxes, [], d)
    want_rects = craft_post.boxes_to_rects(boxes, H, W)
    assert [list(map(int, r)) for r in rects] == want_rects
    assert line_ids == [-1] * len(boxes)
    n_clipped = 0
    for r, f in zip(want_rects, frags):
        mine = craft_post.crop_rect(image, r)
        assert mine.shape == f.shape and np.array_equal(mine, f), r
        n_clipped += (r[0] + r[2] + 1 > W) or (r[1] + r[3] + 1 > H) or r[0] == 0 or r[1] == 0
    assert n_clipped > 20


@pytest.mark.parametrize("beam,max_len_b", [(1, 40), (3, 40), (5, 40), (2, 4)])
def test_search_against_reference_generate_loop(beam, max_len_b):
    """oracle.generate == the reference's own TextRecognitionGenerator._generate (generator.py:11-374, executed from the
    file; fairseq's BeamSearch.step / finalize_hypos restated in oracle/ref_loader.load_generator) on the same
    incremental decoder: tokens, scores and positional scores of EVERY finalised hypothesis, hypotheses of different
    lengths (so the loop's batch compaction :262-297 runs) and the max_len EOS forcing (:165-167)."""
    from oracle import trocr
    Gen, make_model = ref_loader.load_generator()
    cfg = trocr.trocr_tiny()
    sd = trocr.synth_trocr_state(cfg, 8, round_to=None)
    torch.manual_seed(3)
    imgs = torch.stack([(torch.rand(3, 384, 384) * 2 - 1) * a + b for a, b in
                        [(1, 0), (0.2, 0.7), (0.5, -0.5), (0.1, -0.9), (1, 0.0), (0.3, 0.3), (0.05, 0.95)]]).clamp(-1, 1)
    with torch.no_grad():
        trocr.calibrate_eos(sd, cfg, eos_step=6, round_to=None, margin=0.25, enc=trocr.encoder_forward(sd, cfg, imgs[:2]))
        enc = trocr.encoder_forward(sd, cfg, imgs)
        mine = trocr.generate(sd, cfg, enc, beam=beam, max_len_b=max_len_b)
        gen = Gen(make_model(sd, cfg), cfg.vocab, beam_size=beam, max_len_b=max_len_b)
        ref = gen._generate({"net_input": {"imgs": imgs}})
    lens = set()
    for a, b in zip(mine, ref):
        assert len(a) == len(b) == beam
        for ha, hb in zip(a, b):
            assert ha["tokens"].tolist() == hb["tokens"].tolist()
            assert abs(ha["score"] - float(hb["score"])) <= 1e-5
            assert torch.allclose(ha["positional_scores"], hb["positional_scores"], atol=1e-5)
        lens.add(max(len(h["tokens"]) for h in a))            # the step the sentence left the batch
    if max_len_b > 10:
        assert len(lens) > 1, "sentences should finish at different steps"
    else:
        assert all(len(h["tokens"]) <= max_len_b + 1 for hyps in mine for h in hyps)


def test_detokenizer_against_reference_bpe_decode(tmp_path):
    """Gpt2Detokenizer.decode(ids) == GPT2BPEEnhancedSpace.decode(Dictionary.string(ids)) with the reference's own
    decode (marie/models/unilm/trocr/bpe.py:59-67, executed from the file) on a miniature vocabulary."""
    import os
    from marie_icr_b200.bpe import Gpt2Detokenizer
    from test_host_logic import _tiny_gpt2_files
    order = _tiny_gpt2_files(str(tmp_path))
    ref = ref_loader.load_bpe()(os.path.join(str(tmp_path), "encoder.json"))
    detok = Gpt2Detokenizer.locate(str(tmp_path))
    rng = np.random.default_rng(7)
    n_sym = len(detok.symbols)
    for _ in range(200):
        ids = rng.integers(3, n_sym, int(rng.integers(1, 9))).tolist() + [2]
        # fairseq Dictionary.string with extra_symbols_to_ignore = {eos}: bos / eos dropped, everything else joined by ' '
        hypo_str = " ".join(detok.symbols[t] for t in ids if t not in (0, 2))
        assert detok.decode(ids) == ref.decode(hypo_str), ids


def test_convert_frames_and_json_store_against_reference(tmp_path):
    """ingest.convert_frames / load_image's TIFF branch and executor.store_json_object against the reference's own
    functions (marie/utils/docs.py:183-256 extracted from the source — the module imports PyPDF4 / docarray;
    marie/utils/json.py:19-30 loaded behind a stub for its numpy encoder import)."""
    import ast
    import json
    import os
    import cv2
    from marie_icr_b200 import ingest
    from marie_icr_b200.executor import store_json_object
    path = os.path.join(ref_loader.REF_ROOT, "marie", "utils", "docs.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("convert_frames", "load_image")]
    for fn in fns:
        fn.returns = None
    from PIL import Image
    ns = dict(cv2=cv2, np=np, Image=Image, List=list, get_document_type=lambda p: ingest.document_type(p), load_pdf_frames=None)
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    rng = np.random.default_rng(8)
    pages = [rng.integers(0, 256, (21, 33, 3), dtype=np.uint8), rng.integers(0, 256, (21, 33), dtype=np.uint8)]
    src = str(tmp_path / "d.tif")
    cv2.imwritemulti(src, pages)
    ok_ref, ref_frames = ns["load_image"](src)
    ok, frames = ingest.load_image(src)
    assert ok == ok_ref and len(frames) == len(ref_frames) and all(np.array_equal(a, b) for a, b in zip(frames, ref_frames))
    png = str(tmp_path / "p.png")
    cv2.imwrite(png, pages[0])
    assert np.array_equal(ingest.load_image(png)[1][0], ns["load_image"](png)[1][0])
    # store_json_object: same bytes as the reference writes for a page record holding numpy values
    spec_path = os.path.join(ref_loader.REF_ROOT, "marie", "utils", "json.py")
    with open(spec_path) as f:
        jtree = ast.parse(f.read())
    keep = [n for n in jtree.body if isinstance(n, ast.FunctionDef) and n.name == "store_json_object"]
    enc_path = os.path.join(ref_loader.REF_ROOT, "marie", "numpyencoder.py")
    enc_ns = {}
    src_enc = open(enc_path).read().replace("np.float_, ", "").replace("np.complex_, ", "")   # aliases removed in NumPy 2
    exec(compile(src_enc, enc_path, "exec"), enc_ns)
    jns = dict(json=json, os=os, EnhancedJSONEncoder=enc_ns["NumpyEncoder"])
    exec(compile(ast.Module(body=keep, type_ignores=[]), spec_path, "exec"), jns)
    record = [{"meta": {"imageSize": {"width": np.int64(3), "height": 4}, "page": 0}, "words": [{"id": np.int32(0), "text": "É", "confidence": np.float32(0.5), "box": np.array([1, 2, 3, 4])}]}]
    a, b = str(tmp_path / "a.json"), str(tmp_path / "b.json")
    jns["store_json_object"](record, a)
    store_json_object(record, b)
    assert open(a).read() == open(b).read()
