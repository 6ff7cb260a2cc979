"""Batched device pipeline: pages -> CRAFT boxes -> crops -> TrOCR token ids, everything resident in HBM between the
stages (the reference round-trips through the host after every stage: score maps `.cpu()` at
marie/boxes/craft_box_processor.py:113-114, one H2D per crop at marie/document/trocr_ocr_processor.py:124, two D2H
per hypothesis at :160-163).

Stage map: K1 mb_page_preprocess -> K2-K4 mb_craft_forward -> K5-K7 mb_craft_post ->
K9 mb_pack_crops (patch-row layout) -> K10 mb_trocr_encode -> K11/K12 mb_trocr_decode.
"""
import numpy as np
import torch

from . import ops
from ._lib import Context

# (text_threshold, link_threshold, low_text) per page-segmentation mode — the hard-coded presets of
# marie/boxes/craft_box_processor.py:317-428
PSM_PRESETS = {
    "word": (0.6, 0.8, 0.3),
    "sparse": (0.7, 0.45, 0.3),
    "line": (0.4, 0.2, 0.3),
    "raw_line": (0.4, 0.2, 0.5),
    "multiline": (0.6, 0.3, 0.3),
}

RECORD_HEAD = 8   # page, x, y, w, h, line, length, score(bits)


class _StageTimer:
    """CUDA-event stage timers on the current stream (bench.py per-stage breakdown); a no-op unless enabled."""

    def __init__(self):
        self.enabled = False
        self.pending, self.ms, self.units = [], {}, {}

    def reset(self, enabled):
        self.enabled = enabled
        self.pending, self.ms, self.units = [], {}, {}

    def start(self):
        if not self.enabled:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def stop(self, name, e0, units=0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.pending.append((name, e0, e1, units))

    def collect(self):
        torch.cuda.synchronize()
        for name, e0, e1, units in self.pending:
            self.ms[name] = self.ms.get(name, 0.0) + e0.elapsed_time(e1)
            self.units[name] = self.units.get(name, 0) + units
        self.pending = []
        return dict(ms=dict(self.ms), units=dict(self.units))


def decode_batches(n, chunk):
    """Sizes of the decode batches for n crops, none above `chunk`.  One batch when it fits; otherwise k - 1 equal batches
    and a LAST one of a third of their size: the host post-processing of a batch (detokenise + record assembly, ~10 us a
    word) runs under the next batch's device work, so only the last batch's share is exposed at the end of a call.  A
    remainder batch much shorter than that would cost a whole decode loop for few rows."""
    if n <= chunk:
        return [n] if n > 0 else []
    k = 2
    while -(-3 * n // (3 * k - 2)) > chunk:
        k += 1
    a = -(-3 * n // (3 * k - 2))
    if n - a * (k - 1) < 1:                      # tiny chunks: plain equal batches
        a = -(-n // k)
        return [a] * (n // a) + ([n % a] if n % a else [])
    return [a] * (k - 1) + [n - a * (k - 1)]


def default_crop_chunk(beam, enc_dim=768):
    """Crops per decode batch that the bench / engine use by default.  No cross-attention K/V cache is kept (xattn_tc.cu
    attends over the encoder states themselves, for every beam width), so the batch is bounded by what grows with the rows
    = crops x beam: the fp32 logits (0.2 MB per row) and the self-attention K / V caches (1.5 MB per row at 32 steps).
    16384 crops greedy; beams: about 41 k rows."""
    if beam <= 1:
        return 16384
    return max(1024, min(8192, (40960 // beam) // 512 * 512))


class PagePipeline:
    """Owns the per-device context and the loaded models."""

    def __init__(self, device=0, craft_blob=None, trocr_blob=None, micro_batch=8, crop_chunk=1024, max_labels=8192,
                 max_boxes=4096, encode_chunk=2048):
        self.device = int(device)
        self.ctx = Context.get(self.device)
        self.micro_batch = micro_batch
        self.crop_chunk = crop_chunk          # crops per decode batch (larger = fewer, better filled decoder steps)
        self.patch_budget_bytes = 24 << 30    # K9 output of one decode batch kept whole up to this size
        self.encode_chunk = encode_chunk      # crops per K9 + encoder pass inside a decode batch (measured: the encoder is
                                              # fastest around 2048 crops, the decoder keeps gaining up to 8192)
        self.max_labels, self.max_boxes = max_labels, max_boxes
        self.has_craft = self.has_trocr = False
        self.timer = _StageTimer()
        if craft_blob is not None:
            self.load_craft(craft_blob)
        if trocr_blob is not None:
            self.load_trocr(trocr_blob)

    @property
    def dtype(self):
        return self.ctx.torch_dtype

    def load_craft(self, blob):
        ops.load_craft(blob, self.device)
        self.has_craft = True

    def load_trocr(self, blob):
        ops.load_trocr(blob, self.device)
        self.has_trocr = True

    # ------------------------------------------------------------------------------------------ detection
    def detect(self, pages_dev, preset=PSM_PRESETS["sparse"], keep_maps=False, line_refiner=False, ready=None,
               keep_post=False):
        """pages_dev [n,H,W,3] u8 (BGR) on the device -> dict with per-crop `rects` [N,4] i32 (x,y,w,h), `page_idx`
        [N] i32, `boxes` [N,4,2] f32 (page coordinates, adjustResultCoordinates output), `counts` (host list).
        ready(i0, i1): optional hook called before pages [i0, i1) are first read (run_frames: the compute stream waits
        there for the copy stream's H2D of that micro-batch)."""
        if not self.has_craft:
            raise RuntimeError("CRAFT weights are not loaded")
        n, ph, pw, _ = pages_dev.shape
        tt, lt, low = preset
        rects, boxes, pidx, counts = [], [], [], []
        scores_all, ratio = None, None
        refined_all = None                      # line branch (craft_box_processor.py:150-217), off in the reference
        # K1 + CRAFT in micro-batches (activation memory), score maps of the whole batch kept in HBM ...
        for i0 in range(0, n, self.micro_batch):
            chunk = pages_dev[i0:i0 + self.micro_batch]
            m = chunk.shape[0]
            if ready is not None:
                ready(i0, i0 + m)
            t = self.timer.start()
            x, ratio = ops.page_preprocess(chunk)
            self.timer.stop("k1_preprocess", t, m)
            t = self.timer.start()
            if line_refiner:
                scores, feat = ops.craft_forward(x, want_feature=True)
                refined = ops.refine_forward(scores, feat)
                del feat
                if refined_all is None:
                    refined_all = torch.empty((n,) + tuple(refined.shape[1:]), dtype=torch.float32, device=pages_dev.device)
                refined_all[i0:i0 + m] = refined
            else:
                scores = ops.craft_forward(x)
            self.timer.stop("k2_4_craft", t, m)
            del x
            if scores_all is None:
                scores_all = torch.empty((2, n) + tuple(scores.shape[2:]), dtype=torch.float32, device=pages_dev.device)
            scores_all[:, i0:i0 + m] = scores
        # ... then ONE post-processing pass over all pages (launch latency amortised over the batch)
        r2 = (1.0 / ratio) * 2
        t = self.timer.start()
        out = ops.craft_post(scores_all[0], scores_all[1], tt, lt, low, ratios=[(r2, r2)] * n, page_hw=[(ph, pw)] * n,
                             max_labels=self.max_labels, max_boxes=self.max_boxes)
        self.timer.stop("k5_7_post", t, n)
        nb = out["n_boxes"].cpu().tolist()          # mb_craft_post has synchronised the stream already
        for j in range(n):
            rects.append(out["rects"][j, :nb[j]])
            boxes.append(out["adj"][j, :nb[j]])
            pidx.append(torch.full((nb[j],), j, dtype=torch.int32, device=pages_dev.device))
        counts = nb
        res = dict(rects=torch.cat(rects).contiguous(), boxes=torch.cat(boxes).contiguous(),
                   page_idx=torch.cat(pidx).contiguous(), counts=counts)
        if line_refiner:
            res["lines"] = ops.line_boxes(refined_all, lt, 1.0 / ratio, 1.0 / ratio)
            if keep_maps:
                res["refined_link"] = refined_all
        if keep_maps:
            res["scores"] = scores_all
        if keep_post:       # K5-K7 outputs at heat-map scale (polygon refinement, polys.py) and the resize ratio of K1
            res["post"] = dict(labels=out["labels"], det=out["det"], mapper=out["mapper"])
            res["ratio"] = ratio
        return res

    # ------------------------------------------------------------------------------------------ recognition
    def recognize_crops(self, pages_dev, rects, page_idx, beam=1, max_len_b=200, out_ld=32, on_batch=None):
        """Crops addressed by (page, rect) -> (tokens [N,out_ld] i32, lengths [N] i32, scores [N] f32) on the device.
        on_batch(i0, m): optional hook called after every decode batch, when rows [i0, i0+m) of the outputs are final."""
        if not self.has_trocr:
            raise RuntimeError("TrOCR weights are not loaded")
        n = rects.shape[0]
        dev = pages_dev.device
        tokens = torch.full((n, out_ld), 1, dtype=torch.int32, device=dev)
        lengths = torch.zeros((n,), dtype=torch.int32, device=dev)
        scores = torch.zeros((n,), dtype=torch.float32, device=dev)
        if n == 0:
            return tokens, lengths, scores
        dims = ops.trocr_dims(self.device)
        i0 = 0
        for m in decode_batches(n, self.crop_chunk):
            enc = torch.empty((m, dims["tokens"], dims["enc_dim"]), dtype=self.dtype, device=dev)
            # K9 for the whole decode batch in one launch set when its patch rows fit the budget (0.88 MB per crop: 14.5 GB
            # for 16384 crops): one classification / one error check instead of one per encoder pass, and the kernel is
            # not a 0.6 ms island between two power-capped encoder passes
            whole = m * 576 * 768 * 2 <= self.patch_budget_bytes
            patches_all = None
            if whole:
                e = self.timer.start()
                patches_all = ops.pack_crops(pages_dev, rects[i0:i0 + m].contiguous(), page_idx[i0:i0 + m].contiguous(), layout=1)
                self.timer.stop("k9_crops", e, m)
            for j0 in range(0, m, self.encode_chunk):
                j1 = min(j0 + self.encode_chunk, m)
                if whole:
                    patches = patches_all[j0 * 576:j1 * 576]
                else:
                    r = rects[i0 + j0:i0 + j1].contiguous()
                    p = page_idx[i0 + j0:i0 + j1].contiguous()
                    e = self.timer.start()
                    patches = ops.pack_crops(pages_dev, r, p, layout=1)
                    self.timer.stop("k9_crops", e, j1 - j0)
                e = self.timer.start()
                ops.trocr_encode(patches, out=enc[j0:j1])
                self.timer.stop("k10_encoder", e, j1 - j0)
                del patches
            del patches_all
            e = self.timer.start()
            t, l, s, _ = ops.trocr_decode(enc, beam=beam, max_len_b=max_len_b, out_ld=out_ld)
            self.timer.stop("k11_12_decode", e, m)
            del enc
            tokens[i0:i0 + m] = t
            lengths[i0:i0 + m] = l
            scores[i0:i0 + m] = s
            if on_batch is not None:
                on_batch(i0, m, tokens, lengths, scores)
            i0 += m
        return tokens, lengths, scores

    def recognize_fragments(self, fragments, beam=1, max_len_b=200, out_ld=32):
        """Host fragments (list of [h,w,3] u8 BGR) -> (tokens, lengths, scores) on the device; one H2D for all."""
        if not self.has_trocr:
            raise RuntimeError("TrOCR weights are not loaded")
        n = len(fragments)
        dev = f"cuda:{self.device}"
        tokens = torch.full((n, out_ld), 1, dtype=torch.int32, device=dev)
        lengths = torch.zeros((n,), dtype=torch.int32, device=dev)
        scores = torch.zeros((n,), dtype=torch.float32, device=dev)
        for i0 in range(0, n, self.crop_chunk):
            part = fragments[i0:i0 + self.crop_chunk]
            patches = ops.pack_fragments(part, device=dev, layout=1)
            t, l, s = ops.trocr_recognize(patches, beam=beam, max_len_b=max_len_b, chunk=self.crop_chunk, out_ld=out_ld)
            tokens[i0:i0 + len(part)] = t
            lengths[i0:i0 + len(part)] = l
            scores[i0:i0 + len(part)] = s
        return tokens, lengths, scores

    # ------------------------------------------------------------------------------------------ whole path
    def run_device(self, pages_dev, preset=PSM_PRESETS["sparse"], beam=1, max_len_b=200, out_ld=32, line_refiner=False,
                   ready=None, want_lines=None, sink=None):
        """Pages already in HBM -> packed per-word records [N, RECORD_HEAD + out_ld] i32 on the device
        (page, x, y, w, h, line, length, score bits, tokens...) and the per-page counts (host).  line = -1
        (find_line_number over an empty line list, as in the reference) unless `line_refiner` runs the refiner's line
        branch (craft_box_processor.py:150-217): then the merged line boxes are appended per page to `want_lines` (a
        list) and every word gets its line number (line_processor.py:15-45).  sink(block, counts): optional consumer of
        the finished records, called with a host copy of every decode batch's rows (in order) as soon as they are final."""
        det = self.detect(pages_dev, preset, line_refiner=line_refiner, ready=ready)
        n = det["rects"].shape[0]
        rec = torch.empty((n, RECORD_HEAD + out_ld), dtype=torch.int32, device=pages_dev.device)
        if n:
            rec[:, 0] = det["page_idx"]
            rec[:, 1:5] = det["rects"]
            rec[:, 5] = -1                       # find_line_number(lines_bboxes=[], box) == -1 (line_processor.py:21-45)
            if line_refiner:
                from . import lines as _lines
                rects_h = det["rects"].cpu().numpy()
                ids, k = np.full((n,), -1, np.int32), 0
                for j, c in enumerate(det["counts"]):
                    if c and len(det["lines"][j]):
                        ids[k:k + c] = _lines.find_line_numbers(det["lines"][j], rects_h[k:k + c])
                    k += c
                rec[:, 5] = torch.from_numpy(ids).to(rec.device)

            def on_batch(i0, m, tokens, lengths, scores):
                rec[i0:i0 + m, 6] = lengths[i0:i0 + m]
                rec[i0:i0 + m, 7] = scores[i0:i0 + m].view(torch.int32)
                rec[i0:i0 + m, RECORD_HEAD:] = tokens[i0:i0 + m]
                if sink is not None:             # records of a finished decode batch leave for the host while the next one runs
                    sink(rec[i0:i0 + m].cpu().numpy(), det["counts"])

            self.recognize_crops(pages_dev, det["rects"], det["page_idx"], beam, max_len_b, out_ld, on_batch=on_batch)
        if want_lines is not None and line_refiner:
            want_lines.extend(det["lines"])
        return rec, det["counts"]

    def run_frames(self, frames, **kw):
        """Host frames (list of equally sized [H,W,3] u8 BGR arrays — what OcrEngine.extract receives) -> records on the
        host.  The frames are copied ONCE, into a pinned staging buffer kept by the pipeline (this is also the deep copy
        the reference makes, ocr_engine.py:118,416-433), by a helper thread one micro-batch ahead; each micro-batch
        goes to the device on a copy stream while the previous one is in K1 / CRAFT on the compute stream."""
        import threading
        n = len(frames)
        shape = (n,) + tuple(frames[0].shape)
        dev = torch.device("cuda", self.device)
        numel = int(np.prod(shape))
        if getattr(self, "_stage", None) is None or self._stage.numel() < numel:
            self._stage = torch.empty((numel,), dtype=torch.uint8).pin_memory()
            self._copy_stream = torch.cuda.Stream(device=dev)
        stage = self._stage[:numel].view(shape)
        stage_np = stage.numpy()
        pages_dev = torch.empty(shape, dtype=torch.uint8, device=dev)
        mb = self.micro_batch
        staged = [threading.Event() for _ in range(0, n, mb)]

        def stager():
            # the pages of a micro-batch are copied by a few threads at once (np.copyto releases the GIL): only the first
            # micro-batch's staging is exposed in front of the device work, the others run one batch ahead of it
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=4) as pool:
                for k, i0 in enumerate(range(0, n, mb)):
                    list(pool.map(lambda i: np.copyto(stage_np[i], frames[i]), range(i0, min(i0 + mb, n))))
                    staged[k].set()

        th = threading.Thread(target=stager, daemon=True)
        th.start()
        compute = torch.cuda.current_stream(dev)
        copied = {}

        def issue(k):
            if k < len(staged) and k not in copied:
                staged[k].wait()
                i0 = k * mb
                with torch.cuda.stream(self._copy_stream):
                    pages_dev[i0:i0 + mb].copy_(stage[i0:i0 + mb], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                copied[k] = ev

        def ready(i0, i1):
            k = i0 // mb
            issue(k)
            issue(k + 1)                      # next micro-batch's H2D runs under this one's CRAFT
            compute.wait_event(copied[k])

        self._copy_stream.wait_stream(compute)           # pages_dev allocation ordering
        rec, counts = self.run_device(pages_dev, ready=ready, **kw)
        th.join()
        return (rec.cpu() if kw.get("sink") is None else None), counts

    def run_host(self, pages_pinned, **kw):
        """Host pages ([n,H,W,3] u8, ideally pinned) -> records on the host; H2D and D2H inside."""
        pages_dev = pages_pinned.to(f"cuda:{self.device}", non_blocking=True)
        rec, counts = self.run_device(pages_dev, **kw)
        return rec.cpu(), counts


def records_to_words(rec, detok, page=None):
    """Host records -> list of dicts {page, box [x,y,w,h], line, text, confidence, tokens} in detector order.
    Text is upper-cased and the confidence is round(round(exp(score), 6), 4) as in
    marie/document/trocr_ocr_processor.py:159-160,338-341 (exp in float32 like torch.exp on the fp32 score).
    A hypothesis longer than the record's token field (only possible with an explicit, small out_ld) is flagged
    `truncated` instead of silently losing its tail."""
    rec = rec.numpy() if hasattr(rec, "numpy") else np.asarray(rec)
    if page is not None:
        rec = rec[rec[:, 0] == page]
    n = rec.shape[0]
    if n == 0:
        return []
    out_ld = rec.shape[1] - RECORD_HEAD
    lens = rec[:, 6]
    conf = np.exp(np.ascontiguousarray(rec[:, 7]).view(np.float32)).tolist()
    head = rec[:, :RECORD_HEAD].tolist()
    toks = rec[:, RECORD_HEAD:RECORD_HEAD + max(1, min(out_ld, int(lens.max())))].tolist()    # only the columns in use
    out = []
    for h, t, c in zip(head, toks, conf):
        ln = h[6]
        t = t[:ln]
        w = dict(page=h[0], box=h[1:5], line=h[5], tokens=t, text=detok.decode(t).upper(),
                 confidence=round(round(c, 6), 4) if ln else 0.0)
        if ln > out_ld:
            w["truncated"] = True
        out.append(w)
    return out
