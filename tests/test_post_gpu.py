"""K5-K7 parity: CUDA labelling / boxes / rects vs the CPU oracle (oracle/craft_post.py, cv2 flavour, itself pinned
against the reference's getDetBoxes_core).  Integer outputs must be bit-exact; float boxes are compared bit-exact
too, with the documented exception of exact area ties inside cv2's rotating calipers."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PRESETS = [(0.7, 0.45, 0.3), (0.4, 0.2, 0.3), (0.6, 0.3, 0.3)]   # sparse / line / multiline (craft_box_processor.py:340-428)


def _run(maps, preset, ratio=(2.588, 2.588), page=(3300, 2550)):
    from marie_icr_b200 import ops
    from oracle import craft_post
    tt, lt, low = preset
    text = torch.from_numpy(np.stack([m[0] for m in maps])).cuda()
    link = torch.from_numpy(np.stack([m[1] for m in maps])).cuda()
    n = len(maps)
    out = ops.craft_post(text, link, tt, lt, low, ratios=[ratio] * n, page_hw=[page] * n)
    torch.cuda.synchronize()
    tot = mism = 0
    for i, (t, l) in enumerate(maps):
        det, labels, mapper = craft_post.det_boxes_cv(t, l, tt, lt, low)
        assert np.array_equal(out["labels"][i].cpu().numpy(), labels), f"labels differ on image {i}"
        nl, _, stats, _, _ = craft_post.label_maps(t, l, lt, low)
        assert int(out["n_labels"][i]) == nl
        assert np.array_equal(out["stats"][i, 1:nl].cpu().numpy(), stats[1:nl])
        nb = int(out["n_boxes"][i])
        assert nb == len(det)
        assert out["mapper"][i, :nb].cpu().tolist() == mapper
        if nb == 0:
            continue
        got = out["det"][i, :nb].cpu().numpy()
        ref = np.stack(det)
        same = (got == ref).reshape(nb, -1).all(1)
        tot += nb
        mism += int((~same).sum())
        assert np.abs(got - ref).max() < 1e-3, "box differs beyond a calipers tie"
        adj = craft_post.adjust_result_coordinates([b.copy() for b in ref], ratio[0] / 2, ratio[1] / 2)
        rects = np.array(craft_post.boxes_to_rects(adj, page[0], page[1]))
        got_rects = out["rects"][i, :nb].cpu().numpy()
        ok = same
        assert np.array_equal(got_rects[ok], rects[ok])
        assert np.array_equal(out["adj"][i, :nb].cpu().numpy()[ok], np.asarray(adj, dtype=np.float32)[ok])
    return tot, mism


@pytest.mark.parametrize("preset", PRESETS)
def test_random_blobs(cuda_ctx, preset):
    from oracle import synth
    maps = [synth.random_score_maps(s, 200 + 7 * s, 320 + 5 * s if False else 320, n_blobs=50) for s in range(4)]
    maps = [synth.random_score_maps(s, 200, 320, n_blobs=50) for s in range(6)]
    tot, mism = _run(maps, preset)
    assert tot > 50
    assert mism <= max(1, tot // 200), f"{mism}/{tot} boxes not bit-exact"


def test_glyph_maps_letter_heatmap(cuda_ctx):
    """Realistic ~500-component maps at the letter-page heat-map size 1280x992 (config 1/2 shapes)."""
    from oracle import synth
    maps = []
    for s in range(2):
        page, _ = synth.synth_page(s)
        maps.append(synth.score_maps_from_page(page, 1280, 992))
    tot, mism = _run(maps, PRESETS[0])
    assert tot > 600
    assert mism <= max(1, tot // 200), f"{mism}/{tot} boxes not bit-exact"


def test_edge_cases(cuda_ctx):
    h, w = 64, 96
    empty = (np.zeros((h, w), np.float32), np.zeros((h, w), np.float32))
    full = (np.ones((h, w), np.float32), np.zeros((h, w), np.float32))
    specks = np.zeros((h, w), np.float32)
    specks[::4, ::4] = 1.0                      # many 1-px components, all below the area filter
    border = np.zeros((h, w), np.float32)
    border[0:5, 0:30] = 0.9
    border[h - 4:h, w - 20:w] = 0.8
    border[20:40, w - 3:w] = 0.75
    linkonly = np.zeros((h, w), np.float32)
    linkonly[10:20, 10:60] = 0.9               # link-only component: no text pixel -> filtered by max(text)
    maps = [empty, full, (specks, np.zeros_like(specks)), (border, np.zeros_like(border)),
            (np.zeros_like(linkonly), linkonly)]
    tot, mism = _run(maps, PRESETS[0], ratio=(2.0, 2.0), page=(128, 192))
    assert mism == 0
