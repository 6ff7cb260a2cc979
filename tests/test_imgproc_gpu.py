"""K1 / K9 parity: the two exact resamplers vs the CPU oracle (oracle/resample.py, pinned against cv2 / PIL)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _q(x, dt):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dt)


@pytest.mark.parametrize("ph,pw", [(330, 255), (825, 640), (200, 300), (3300, 2550), (1400, 300), (301, 997)])
def test_page_preprocess(cuda_ctx, dtype16, ph, pw):
    from marie_icr_b200 import ops
    from oracle import resample
    rng = np.random.default_rng(ph)
    pages = rng.integers(0, 256, (2, ph, pw, 3), dtype=np.uint8)
    out, ratio = ops.page_preprocess(torch.from_numpy(pages).cuda())
    torch.cuda.synchronize()
    for i in range(2):
        ref, r = resample.craft_input(pages[i])
        assert r == ratio
        got = out[i].cpu()
        assert got.shape[:2] == ref.shape[:2]
        assert torch.equal(got[..., :3], _q(ref, dtype16)), "preprocessed page differs from cv2 fixed-point resize"
        assert torch.all(got[..., 3] == 0)


def test_pack_fragments_exact(cuda_ctx, dtype16):
    from marie_icr_b200 import ops
    from oracle import resample
    rng = np.random.default_rng(3)
    shapes = [(40, 120), (55, 300), (384, 384), (384, 200), (100, 384), (70, 900), (500, 40), (3, 5), (1, 1),
              (61, 2500), (47, 383), (33, 77)]
    frags = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    out = ops.pack_fragments(frags)
    torch.cuda.synchronize()
    for i, f in enumerate(frags):
        ref = resample.fragment_to_input(f)
        assert torch.equal(out[i].cpu(), _q(ref, dtype16)), f"fragment {shapes[i]} differs from PIL bicubic"
    # patch layout carries the same values
    outp = ops.pack_fragments(frags[:3], layout=1).reshape(3, 24, 24, 3, 16, 16)
    chw = outp.permute(0, 3, 1, 4, 2, 5).reshape(3, 3, 384, 384)
    assert torch.equal(chw, out[:3])


def test_pack_crops_from_page(cuda_ctx, dtype16):
    from marie_icr_b200 import ops
    from oracle import resample, craft_post
    rng = np.random.default_rng(4)
    pages = rng.integers(0, 256, (2, 400, 600, 3), dtype=np.uint8)
    rects = np.array([[10, 20, 100, 30], [0, 0, 50, 50], [500, 350, 100, 50], [590, 390, 30, 30], [100, 100, 300, 200]], np.int32)
    pidx = np.array([0, 1, 1, 0, 1], np.int32)
    out = ops.pack_crops(torch.from_numpy(pages).cuda(), torch.from_numpy(rects).cuda(), torch.from_numpy(pidx).cuda())
    torch.cuda.synchronize()
    for i in range(len(rects)):
        frag = craft_post.crop_rect(pages[pidx[i]], rects[i])
        assert torch.equal(out[i].cpu(), _q(resample.fragment_to_input(frag), dtype16)), f"crop {i}"
