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
