"""Parity at BASELINE.json's full sizes (letter page 2550x3300 -> net input 2560x1984 -> heat map 1280x992) and the
size-independent properties the domain offers for the 64-page batch: determinism, batch-order invariance, duplicate
pages giving identical records, and decode == teacher-forced replay."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def letter(cuda_ctx):
    from marie_icr_b200 import weights
    from marie_icr_b200.pipeline import PagePipeline
    from synthetic import pages as synth, weights as sw
    cuda_ctx.set_dtype("fp16")
    cfg = sw.trocr_tiny()
    tsd = sw.synth_trocr_state(cfg, 3)
    craft_sd = sw.glyph_craft_state(0)
    pipe = PagePipeline(craft_blob=weights.pack_craft(craft_sd), trocr_blob=weights.pack_trocr(tsd, cfg), micro_batch=4,
                        crop_chunk=1024)
    pages = np.stack([synth.synth_page(i)[0] for i in range(3)])
    return pipe, pages, craft_sd


def test_craft_letter_page_vs_fp32_oracle(letter):
    """The whole CRAFT network on one full letter page against the fp32 CPU oracle (~10 s of host time), then the
    oracle's post-processing on the device's own maps: identical labels, box list and rects at the full heat-map size."""
    from marie_icr_b200 import ops
    from oracle import craft_net, craft_post, resample
    pipe, pages, craft_sd = letter
    page = pages[0]
    dev = torch.from_numpy(page[None]).cuda()
    x, ratio = ops.page_preprocess(dev)
    ref_x, ref_ratio = resample.craft_input(page)
    assert ratio == ref_ratio and tuple(x.shape) == (1, 2560, 1984, 4)
    assert torch.equal(x[0, ..., :3].cpu(), torch.from_numpy(ref_x).half())
    scores = ops.craft_forward(x)
    with torch.no_grad():
        y, _ = craft_net.craft_forward(craft_sd, torch.from_numpy(ref_x).half().float().permute(2, 0, 1)[None])
    for ch in range(2):
        d = scores[ch, 0].cpu() - y[0, ..., ch]
        rel = (d.norm() / y[0, ..., ch].norm()).item()
        print(f"letter page score map {ch}: rel L2 {rel:.2e}")
        assert rel <= 1e-2
    r2 = 2 / ratio
    out = ops.craft_post(scores[0].contiguous(), scores[1].contiguous(), 0.7, 0.45, 0.3, ratios=[(r2, r2)], page_hw=[(3300, 2550)])
    t, l = scores[0, 0].cpu().numpy(), scores[1, 0].cpu().numpy()
    det, labels, mapper = craft_post.det_boxes_cv(t, l, 0.7, 0.45, 0.3)
    nb = int(out["n_boxes"][0])
    assert np.array_equal(out["labels"][0].cpu().numpy(), labels) and nb == len(det) > 450
    assert out["mapper"][0, :nb].cpu().tolist() == mapper
    adj = craft_post.adjust_result_coordinates([b.copy() for b in det], 1 / ratio, 1 / ratio)
    want = np.array(craft_post.boxes_to_rects(adj, 3300, 2550))
    got = out["rects"][0, :nb].cpu().numpy()
    assert (got != want).any(1).sum() <= max(1, nb // 200)


def test_batch_properties(letter):
    """Records are a pure function of the page: duplicates agree, order does not matter, runs are reproducible, and
    every hypothesis ends with EOS within max_len."""
    from marie_icr_b200.pipeline import RECORD_HEAD
    pipe, pages, _ = letter
    batch = torch.from_numpy(np.stack([pages[0], pages[1], pages[0], pages[2]])).cuda()
    kw = dict(beam=1, max_len_b=12, out_ld=16)
    rec, counts = pipe.run_device(batch, **kw)
    rec2, counts2 = pipe.run_device(batch, **kw)
    assert counts == counts2 and torch.equal(rec, rec2)                      # deterministic
    rec = rec.cpu().numpy()
    by_page = [rec[rec[:, 0] == p][:, 1:] for p in range(4)]
    assert counts[0] == counts[2] and np.array_equal(by_page[0], by_page[2])  # duplicate pages -> identical records
    perm = [3, 2, 1, 0]
    recp, countsp = pipe.run_device(batch[perm].contiguous(), **kw)
    recp = recp.cpu().numpy()
    for new, old in enumerate(perm):                                          # batch-order invariance
        assert countsp[new] == counts[old]
        assert np.array_equal(recp[recp[:, 0] == new][:, 1:], by_page[old])
    lens = rec[:, 6]
    assert (lens >= 2).all() and (lens <= 13).all()
    for r in rec[:200]:
        n = int(r[6])
        assert r[RECORD_HEAD + n - 1] == 2 and 1 not in r[RECORD_HEAD:RECORD_HEAD + n].tolist()
    assert sum(counts) == len(rec) and all(c > 450 for c in counts)


def test_chunk_invariance(letter):
    """The recogniser's chunking is a memory knob, not part of the result: word records are identical whether the crops of
    a batch go through K9 / encoder in passes of 256 ... all crops and through the decoder in batches of 100, 777 or all at once."""
    pipe, pages, _ = letter
    batch = torch.from_numpy(np.stack([pages[0], pages[1]])).cuda()
    kw = dict(beam=1, max_len_b=12, out_ld=16)
    old = pipe.crop_chunk, pipe.encode_chunk
    try:
        recs = []
        for chunk, enc_chunk in ((100, 2048), (777, 300), (1 << 20, 256), (1 << 20, 1 << 20)):
            pipe.crop_chunk, pipe.encode_chunk = chunk, enc_chunk
            rec, counts = pipe.run_device(batch, **kw)
            recs.append(rec.clone())
        assert all(torch.equal(recs[0], r) for r in recs[1:]) and sum(counts) == len(recs[0]) > 900
    finally:
        pipe.crop_chunk, pipe.encode_chunk = old
