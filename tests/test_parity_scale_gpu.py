"""Token-id / record parity AT THE BENCHMARK'S OWN SCALE: >= 512 TrOCR-base crops cut from the bench's synthetic letter
pages (greedy and beam 5, fp16 and bf16), 64 TrOCR-large crops, and one page end to end through OcrEngineB200.extract
against the full oracle chain assembled by the reference's own OcrProcessor.recognize.

The oracle side was run in the build container (tools/make_golden_trocr.py -> tests/golden/trocr_scale_*.npz,
e2e_page_*.json): oracle-K1 -> oracle-CRAFT -> cv2 getDetBoxes -> rects -> crops -> Pillow-exact resample -> fp32 ViT ->
fp32 fairseq decoder -> the fairseq search (pinned against the reference's generator.py in
tests/test_oracle_vs_reference.py).  Here the device runs its own detection, K9, encoder and search.

Margin protocol (SURVEY.md hard part 5).  The oracle records, per crop, the smallest gap between neighbouring
candidates its search ever had to order (greedy: top-1 vs top-2 log-prob; beam 5: the top 11 cumulative scores, every
step).  A hypothesis whose smallest gap exceeds MARGIN MUST be bit-identical; closer calls may legitimately be ordered
differently under 16-bit rounding and are counted, not excused silently.  MARGIN is set per mode from measurement (fp16
greedy: every flip ever observed sits below 0.005 nat -> bound at 0.02; bf16 carries 8x the rounding noise -> 0.08; beam 5
compares cumulative scores of ten candidates out of 5 x 50265, whose neighbours are naturally ~1e-3 nat apart -> 0.002 /
0.005).  Random-init logits are nearly flat (confidence ~0.001), so these weights are the WORST case for flips: a trained
model's decisions are orders of magnitude further apart.  For beam 5 the mismatches are additionally classified: rank
flip among the oracle's own finalists whose scores are within SCORE_TOL, or a hypothesis scoring at least the oracle's best
minus SCORE_TOL (a near-tie at a pruning boundary).  Exact-match floors guard against regressions.  The rates are printed
and written to gpurun_out/parity_scale.json (committed under profiles/).
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MARGIN = {("fp16", 1): 0.02, ("bf16", 1): 0.08, ("fp16", 5): 0.002, ("bf16", 5): 0.005}     # nat, per (dtype, beam)
SCORE_TOL = {"fp16": 2e-2, "bf16": 8e-2}
EXACT_FLOOR = {("fp16", 1): 0.93, ("bf16", 1): 0.70, ("fp16", 5): 0.90, ("bf16", 5): 0.55}
_summary = {}


def _golden(name):
    path = os.path.join(HERE, "golden", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated (tools/make_golden_trocr.py)")
    return np.load(path)


def _write_summary():
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_scale.json"), "w") as f:
        json.dump(_summary, f, indent=1)


def _pipeline(ctx, dtype, model):
    from marie_icr_b200 import weights
    from marie_icr_b200.pipeline import PagePipeline
    from synthetic import weights as sw
    ctx.set_dtype(dtype)
    dt = ctx.torch_dtype
    cfg = sw.trocr_base() if model == "base" else sw.trocr_large()
    tsd = sw.apply_eos_row(sw.synth_trocr_state(cfg, 0, round_to=dt), f"trocr_{model}_seed0", round_to=dt)
    return PagePipeline(device=0, craft_blob=weights.pack_craft(sw.glyph_craft_state(0), dt),
                        trocr_blob=weights.pack_trocr(tsd, cfg, dt), micro_batch=2, crop_chunk=1024, encode_chunk=512)


def _compare(g, key, beam, tokens, lengths, scores, dtype):
    want_t, want_l, want_s, margin = g[key + "_tokens"], g[key + "_len"], g[key + "_score"], g[key + "_margin"]
    n = len(want_l)
    thr, tol = MARGIN[(dtype, beam)], SCORE_TOL[dtype]
    tokens, lengths, scores = tokens.cpu().numpy(), lengths.cpu().numpy(), scores.cpu().numpy()
    exact = np.array([lengths[i] == want_l[i] and np.array_equal(tokens[i, :lengths[i]], want_t[i, :want_l[i]]) for i in range(n)])
    bound = margin > thr
    bad = np.nonzero(bound & ~exact)[0]
    score_err = float(np.abs(scores[exact] - want_s[exact]).max()) if exact.any() else 0.0
    stats = dict(crops=int(n), exact=int(exact.sum()), exact_rate=float(exact.mean()), margin=thr,
                 bound_by_margin=int(bound.sum()), bound_and_exact=int((bound & exact).sum()),
                 bound_at_0p05=int((margin > 0.05).sum()), bound_at_0p05_and_exact=int(((margin > 0.05) & exact).sum()),
                 smallest_margin_of_a_mismatch=float(margin[~exact].min()) if (~exact).any() else None,
                 largest_margin_of_a_mismatch=float(margin[~exact].max()) if (~exact).any() else None,
                 max_score_err_on_exact=score_err,
                 all_end_with_eos=bool(all(tokens[i, lengths[i] - 1] == 2 for i in range(n))))
    if key + "_finalists" in g:                      # beam search: classify the mismatches
        ft, fs = g[key + "_finalists"], g[key + "_finalist_scores"]
        flip = better = 0
        for i in np.nonzero(~exact)[0]:
            got = tokens[i, :lengths[i]]
            for j in range(ft.shape[1]):
                L = int((ft[i, j] != 1).sum())
                if L == len(got) and np.array_equal(ft[i, j, :L], got):
                    flip += int(fs[i, 0] - fs[i, j] <= tol)
                    break
            else:
                better += int(scores[i] >= want_s[i] - tol)
        stats.update(mismatch_is_near_tied_finalist=flip, mismatch_scores_as_well_as_oracle_best=better,
                     mismatch_other=int((~exact).sum()) - flip - better)
    return stats, bad, score_err


def _scale_case(ctx, dtype, model):
    from marie_icr_b200.pipeline import PSM_PRESETS
    from synthetic import pages as synth
    g = _golden(f"trocr_scale_{model}_{dtype}.npz")
    pidx = np.broadcast_to(np.asarray(g["page_index"], np.int32), (len(g["rects"]),)).copy()
    pages = np.stack([synth.synth_page(p)[0] for p in range(int(pidx.max()) + 1)])
    pages_dev = torch.from_numpy(pages).cuda()
    rects = torch.from_numpy(np.ascontiguousarray(g["rects"])).cuda()
    page_idx = torch.from_numpy(pidx).cuda()
    case = {}
    failures = []
    pipe = _pipeline(ctx, dtype, model)
    # detection parity against the oracle chain: rects of the device's own K1 -> CRAFT -> K5-K7 on these pages
    det = pipe.detect(pages_dev, PSM_PRESETS["sparse"])
    same = total = 0
    dr, dp = det["rects"].cpu().numpy(), det["page_idx"].cpu().numpy()
    for p in range(pages.shape[0]):
        want = g["rects"][pidx == p]
        got = dr[dp == p][:len(want)]
        total += len(want)
        same += int(sum(np.array_equal(a, b) for a, b in zip(got, want)))
    case["detection"] = dict(rects=int(total), identical_to_oracle_chain=int(same))
    print(f"[{model} {dtype}] detection: {same}/{total} rects identical to the oracle chain")
    if same < (1.0 if dtype == "fp16" else 0.95) * total:
        failures.append(f"detection: only {same}/{total} rects identical")
    for name, beam in (("greedy", 1), ("beam5", 5)):
        tokens, lengths, scores = pipe.recognize_crops(pages_dev, rects, page_idx, beam=beam, max_len_b=int(g["max_len_b"]), out_ld=32)
        torch.cuda.synchronize()
        stats, bad, score_err = _compare(g, name, beam, tokens, lengths, scores, dtype)
        case[name] = stats
        extra = (f", mismatches: {stats['mismatch_is_near_tied_finalist']} near-tied finalists, {stats['mismatch_scores_as_well_as_oracle_best']} "
                 f"as good as the oracle's best, {stats['mismatch_other']} other") if "mismatch_other" in stats else ""
        print(f"[{model} {dtype}] {name}: exact {stats['exact']}/{stats['crops']} ({100 * stats['exact_rate']:.1f} %), bound by margin > "
              f"{stats['margin']}: {stats['bound_and_exact']}/{stats['bound_by_margin']} exact, largest margin of a mismatch "
              f"{stats['largest_margin_of_a_mismatch']}, score err {score_err:.2e}{extra}")
        if len(bad):
            failures.append(f"{name}: {len(bad)} crops with margin > {stats['margin']} differ, e.g. crop {int(bad[0])} "
                            f"(margin {float(g[name + '_margin'][bad[0]]):.4f})")
        if stats["exact_rate"] < EXACT_FLOOR[(dtype, beam)]:
            failures.append(f"{name}: exact rate {stats['exact_rate']:.3f} below {EXACT_FLOOR[(dtype, beam)]}")
        if score_err > SCORE_TOL[dtype]:
            failures.append(f"{name}: score error {score_err}")
        if not stats["all_end_with_eos"]:
            failures.append(f"{name}: hypothesis without EOS")
    _summary[f"{model}_{dtype}"] = case
    _write_summary()
    ctx.set_dtype("fp16")
    assert not failures, failures


def test_bench_crops_base_fp16(cuda_ctx):
    _scale_case(cuda_ctx, "fp16", "base")


def test_bench_crops_base_bf16(cuda_ctx):
    _scale_case(cuda_ctx, "bf16", "base")


def test_bench_crops_large_fp16(cuda_ctx):
    _scale_case(cuda_ctx, "fp16", "large")


# ------------------------------------------------------------------------------------------ end-to-end page record
def _engine(ctx, g, beam):
    from marie_icr_b200.boxes import BoxProcessorCraftB200
    from marie_icr_b200.document import TrOcrProcessorB200
    from marie_icr_b200.engine import OcrEngineB200
    from marie_icr_b200.bpe import SyntheticDetokenizer
    from synthetic import weights as sw
    ctx.set_dtype(g["dtype"])
    dt = ctx.torch_dtype
    cfg = sw.trocr_base()
    tsd = sw.apply_eos_row(sw.synth_trocr_state(cfg, 0, round_to=dt), "trocr_base_seed0", round_to=dt)
    box = BoxProcessorCraftB200(state_dict=sw.glyph_craft_state(0))
    icr = TrOcrProcessorB200(state_dict=tsd, config=cfg, beam=beam, pipeline=box.pipeline, detokenizer=SyntheticDetokenizer())
    return OcrEngineB200(box_processor=box, default_ocr_processor=icr)


def _jsonable(o):
    if isinstance(o, dict):
        return {k: _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return _jsonable(o.tolist())
    if isinstance(o, np.generic):
        return o.item()
    return o


@pytest.mark.parametrize("beam", [1, 3])
@pytest.mark.parametrize("path", ["batched", "pagewise"])
def test_engine_page_record_equals_oracle_chain(cuda_ctx, beam, path):
    """OcrEngineB200.extract(page) == oracle-K1 -> oracle-CRAFT -> oracle post -> oracle crops -> oracle TrOCR ->
    the REFERENCE's OcrProcessor.recognize + the meta fields of __process_extract_fullpage (ocr_engine.py:200-217), as
    JSON.  Words whose search had a call closer than MARGIN may differ in text / confidence (counted); everything
    else — meta, boxes, ids, word order, line structure, line bbox — must be equal."""
    from synthetic import pages as synth
    name = os.path.join(HERE, "golden", f"e2e_page_fp16_beam{beam}.json")
    if not os.path.exists(name):
        pytest.skip("e2e golden not generated")
    with open(name) as f:
        g = json.load(f)
    eng = _engine(cuda_ctx, g, beam)
    page, _ = synth.synth_page(0, **g["page_geometry"])
    if path == "pagewise":            # the reference's page-by-page loop through the two plugin calls
        from marie_icr_b200.plugin_api import CoordinateFormat, PSMode
        got = eng._extract_pagewise([page.copy()], "q", "0", PSMode.SPARSE, CoordinateFormat.XYWH)
    else:
        got = eng.extract([page.copy()])
    cuda_ctx.set_dtype("fp16")
    assert len(got) == 1
    got, want = _jsonable(got[0]), g["result"]
    assert got["meta"] == want["meta"]
    assert len(got["words"]) == len(want["words"])
    # golden word k (x-sorted) -> detector index through its box
    det_index = {tuple(r): i for i, r in enumerate(g["rects"])}
    free = 0
    for a, b in zip(got["words"], want["words"]):
        assert a["box"] == b["box"] and a["id"] == b["id"] and a["line"] == b["line"] and a["word_index"] == b["word_index"]
        margin = g["margins"][det_index[tuple(b["box"])]]
        if margin > MARGIN[("fp16", 1 if beam == 1 else 5)]:
            assert a["text"] == b["text"], (a, b, margin)
            assert abs(a["confidence"] - b["confidence"]) <= 1e-3 + 2e-2 * b["confidence"], (a, b)
        else:
            free += a["text"] != b["text"]
    assert len(got["lines"]) == len(want["lines"])
    for a, b in zip(got["lines"], want["lines"]):
        assert a["line"] == b["line"] and a["wordids"] == b["wordids"] and a["bbox"] == b["bbox"]
        if free == 0:
            assert a["text"] == b["text"]
    bound = sum(m > MARGIN[("fp16", 1 if beam == 1 else 5)] for m in g["margins"])
    print(f"e2e page ({path}, beam {beam}): {len(want['words'])} words, {bound} bound by the margin protocol and equal, "
          f"{free} of the unbound ones differ")
    _summary[f"e2e_{path}_beam{beam}"] = dict(words=len(want["words"]), bound=bound, unbound_differing=int(free))
    _write_summary()


def test_engine_regions_equal_reference_region_loop(cuda_ctx):
    """OcrEngineB200.extract(frames, regions=...) against a golden produced by the REFERENCE's own
    __process_extract_regions (marie/ocr/ocr_engine.py:223-414, executed from its source over oracle-backed processors;
    tools/make_golden_trocr.py regions): WORD / RAW_LINE fields, two SPARSE regions of one shape (one batched detector
    pass on the device), a zero-size and an out-of-bounds region (the page then falls back to empty results, as there)."""
    from synthetic import pages as synth
    name = os.path.join(HERE, "golden", "regions_fp16.json")
    if not os.path.exists(name):
        pytest.skip("regions golden not generated")
    with open(name) as f:
        g = json.load(f)
    eng = _engine(cuda_ctx, g, 1)
    frames = [synth.synth_page(i, **g["page_geometry"])[0] for i in range(2)]
    launches0 = cuda_ctx.launches
    got = eng.extract(frames, regions=[dict(r) for r in g["regions"]])
    assert cuda_ctx.launches > launches0
    cuda_ctx.set_dtype("fp16")
    want = g["result"]
    assert [r["id"] for r in got["regions"]] == [r["id"] for r in want["regions"]]
    assert len(got["extended"]) == len(want["extended"]) == 2
    free = 0
    for page, (ge, we) in enumerate(zip(got["extended"], want["extended"])):
        ge = _jsonable(ge)
        assert ge["meta"] == we["meta"] and len(ge["words"]) == len(we["words"]) and len(ge["lines"]) == len(we["lines"])
        # words are x-sorted; map back to fragment order through (box, occurrence) to find each word's margin
        for k, (a, b) in enumerate(zip(ge["words"], we["words"])):
            assert a["box"] == b["box"] and a["id"] == b["id"] and a["line"] == b["line"]
        margins = g["margins"][page]
        if page == 0:                                       # all boxes start at x = 0: x-sorted order == fragment order
            for a, b, m in zip(ge["words"], we["words"], margins):
                if m > MARGIN[("fp16", 1)]:
                    assert a["text"] == b["text"], (a, b, m)
                else:
                    free += a["text"] != b["text"]
    wtext = {r["id"]: r for r in want["regions"][:7]}
    for r, m in zip(got["regions"][:7], g["margins"][0]):
        if m > MARGIN[("fp16", 1)]:
            assert r["text"] == wtext[r["id"]]["text"] and abs(r["confidence"] - wtext[r["id"]]["confidence"]) <= 1e-3
    assert got["regions"][7:] == want["regions"][7:]         # the skipped-region fall-back: ids with empty results
    bound = sum(m > MARGIN[("fp16", 1)] for m in g["margins"][0])
    print(f"regions: {len(want['regions'])} region results, {bound} of 7 recognised fields bound by the margin protocol, {free} unbound differ")
    _summary["regions"] = dict(results=len(want["regions"]), bound=bound, unbound_differing=int(free))
    _write_summary()
