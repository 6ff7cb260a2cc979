"""SURVEY.md section 8f rank 2 on the device: RefineNet (marie/models/craft/refinenet.py) and the line branch of
get_prediction (marie/boxes/craft_box_processor.py:150-217) against the oracle restatements (oracle/craft_net.py
refine_forward — pinned against the reference module and tests/golden/refine_net.npz; oracle/craft_post.py
line_components_cv / line_boxes — cv2 calls of the reference + its line_merge).
Tolerances: refined link map relative L2 <= 1e-2 (16-bit activations vs fp32); components, labels and line boxes exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,h,w,seed", [(2, 64, 96, 0), (1, 40, 200, 1)])
def test_refine_forward_matches_oracle(cuda_ctx, dtype16, n, h, w, seed):
    from marie_icr_b200 import ops, weights
    from oracle import craft_net
    sd = craft_net.synth_refine_state(seed, random_bn=True, round_to=dtype16, out_gain=4.0, out_bias=0.2)
    ops.load_refine(weights.pack_refine(sd, dtype16))
    g = torch.Generator().manual_seed(seed + 10)
    feat32 = torch.rand(n, 32, h, w, generator=g).to(dtype16).float()              # post-ReLU features
    y = (torch.rand(n, h, w, 2, generator=g) * 1.2 - 0.1)
    y16 = y.to(dtype16).float()                                                     # the device reads the maps as 16-bit
    with torch.no_grad():
        ref = craft_net.refine_forward(sd, y16, feat32)[..., 0]
    feat = torch.zeros(n, h, w, 64, dtype=dtype16)
    feat[..., :32] = feat32.permute(0, 2, 3, 1).to(dtype16)
    scores = y.permute(3, 0, 1, 2).contiguous().cuda()
    out = ops.refine_forward(scores, feat.cuda()).cpu()
    rel = ((out - ref).norm() / ref.norm()).item()
    print("refined link rel L2", rel)
    assert torch.isfinite(out).all() and rel <= (1e-2 if dtype16 == torch.float16 else 3e-2)


def _line_maps(seed, h, w):
    """Link maps that look like the refiner's output on text: horizontal bands with gaps and speckle."""
    from oracle import synth
    rng = np.random.default_rng(seed)
    text, _ = synth.random_score_maps(seed, h, w, n_blobs=max(6, h * w // 900), rotated=False)
    import cv2
    link = cv2.GaussianBlur(text, (21, 3), 6) * 1.6 + (rng.random((h, w)) > 0.995) * 0.9
    return link.astype(np.float32)


@pytest.mark.parametrize("h,w", [(96, 160), (120, 200), (65, 33), (40, 31), (200, 321)])
def test_line_components_exact(cuda_ctx, h, w):
    from marie_icr_b200 import ops
    from oracle import craft_post
    maps = np.stack([_line_maps(s, h, w) for s in (3, 4, 5)])
    out = ops.line_components(torch.from_numpy(maps).cuda(), 0.4, max_labels=4096, want_labels=True)
    for i in range(3):
        labels, boxes = craft_post.line_components_cv(maps[i], 0.4)
        nl = int(out["n_labels"][i])
        assert nl == len(boxes) + 1
        assert np.array_equal(out["labels"][i].cpu().numpy(), labels)
        assert out["stats"][i, 1:nl, :4].cpu().tolist() == boxes
        # the closing itself, bit for bit
        closed = craft_post.close3x3_restated(maps[i] > np.float32(0.4))
        assert np.array_equal(out["labels"][i].cpu().numpy() > 0, closed)


def test_line_boxes_and_plugin(cuda_ctx, dtype16):
    """Line boxes in page coordinates and the plugin's lines / line numbers, against the oracle applied to the device's own
    refined link map."""
    from marie_icr_b200 import ops
    from marie_icr_b200.boxes import BoxProcessorCraftB200
    from marie_icr_b200.plugin_api import PSMode
    from oracle import craft_net, craft_post, lines as olines
    from synthetic import pages as synth, weights as sw
    page, _ = synth.synth_page(3, height=600, width=800, scale=0.9, line_pitch=52, gap=30, margin=40)
    rsd = craft_net.synth_refine_state(2, random_bn=False, round_to=dtype16, out_gain=6.0, out_bias=0.3)
    proc = BoxProcessorCraftB200(state_dict=sw.glyph_craft_state(0), line_refiner_state_dict=rsd)
    pages = torch.from_numpy(page[None]).cuda()
    det = proc.pipeline.detect(pages, keep_maps=True, line_refiner=True)
    refined = det["refined_link"][0].cpu().numpy()
    _, ratio = ops.page_preprocess(pages)
    want = craft_post.line_boxes(refined, 0.45, 1.0 / ratio, 1.0 / ratio)
    assert det["lines"][0] == want
    rects, frags, line_ids, pred, lines_bboxes = proc.extract_bounding_boxes("t", "k", page, PSMode.SPARSE)
    assert lines_bboxes == want and len(rects) == len(frags) == len(line_ids) > 10
    if want:
        assert line_ids == [olines.find_line_number(want, r) for r in rects]
    else:
        assert set(line_ids) == {-1}
