"""OcrEngineB200 — the per-page driver of the path: mirror of OcrEngine / DefaultOcrEngine
(marie/ocr/ocr_engine.py:28-221, marie/ocr/default_ocr_engine.py:15-98) with the processors injected the same way
(`box_processor=`, `default_ocr_processor=`).  `extract` returns the reference's result records
({meta, words, lines} per page).

When both processors are the B200 ones and all frames share one size, the pages go through the batched device
pipeline (detection in micro-batches, crops from all pages pooled into recogniser batches); the records are the same
as those of the page-by-page loop, which is kept for every other case.
"""
import numpy as np

from itertools import chain

from .boxes import BoxProcessorCraftB200
from .document import TrOcrProcessorB200
from .pipeline import PSM_PRESETS, records_to_words
from .ingest import hash_frames_fast
from .plugin_api import MODEL_PATH, CoordinateFormat, PSMode, assemble_result


def copy_frames(frames):
    """ocr_engine.py:416-433: PIL -> BGR ndarray, deep copy."""
    out = []
    for f in frames:
        if not isinstance(f, np.ndarray):
            f = np.array(f)[:, :, ::-1]
        out.append(np.ascontiguousarray(f).copy())
    return out


class OcrEngineB200:
    def __init__(self, models_dir=MODEL_PATH, cuda=True, *, box_processor=None, default_ocr_processor=None, **kwargs):
        """models_dir is the model zoo root (marie/ocr/ocr_engine.py:35-37): CRAFT weights under `<models_dir>/craft`,
        the TrOCR checkpoint at `<models_dir>/trocr/trocr-large-printed.pt`, BPE assets under `<models_dir>/assets`."""
        import os
        if box_processor is None:
            box_processor = BoxProcessorCraftB200(models_dir=os.path.join(models_dir, "craft"), cuda=cuda)
        if default_ocr_processor is None:
            default_ocr_processor = TrOcrProcessorB200(models_dir=models_dir, cuda=cuda,
                                                       pipeline=getattr(box_processor, "pipeline", None))
        self.box_processor = box_processor
        self.icr_processor = default_ocr_processor
        self.has_cuda = cuda
        self.bbox_cache = {}      # region overlay hash -> extract_bounding_boxes result (ocr_engine.py:20-25,324-336)

    def extract(self, frames, pms_mode=PSMode.SPARSE, coordinate_format=CoordinateFormat.XYWH, regions=None,
                queue_id=None, **kwargs):
        queue_id = "0000-0000-0000-0000" if queue_id is None else queue_id
        regions = [] if regions is None else regions
        if isinstance(frames, np.ndarray) and frames.ndim == 3:
            frames = [frames]
        batched = (not len(regions) and isinstance(self.box_processor, BoxProcessorCraftB200)
                   and isinstance(self.icr_processor, TrOcrProcessorB200)
                   and self.box_processor.pipeline is self.icr_processor.pipeline
                   and pms_mode in (PSMode.SPARSE, PSMode.LINE, PSMode.MULTI_LINE)
                   and all(isinstance(f, np.ndarray) and f.ndim == 3 and f.dtype == np.uint8 for f in frames)
                   and len({f.shape for f in frames}) == 1 and not kwargs.get("crop_to_content", False))
        if batched:
            # the batched path copies the caller's frames once, straight into its pinned staging buffer (the deep copy of
            # ocr_engine.py:118 and the H2D source in one) and never writes to them
            return self._extract_batched(frames, pms_mode, coordinate_format)
        ro_frames = copy_frames(frames)
        if len(regions):
            return self._extract_regions(ro_frames, queue_id, hash_frames_fast(ro_frames), pms_mode, regions)
        return self._extract_pagewise(ro_frames, queue_id, "0", pms_mode, coordinate_format,
                                      crop=bool(kwargs.get("crop_to_content", False)))

    # page-by-page loop of __process_extract_fullpage (ocr_engine.py:154-221); `crop_to_content=True` crops every page to its
    # content and pads it with 4 white pixels first (:169-185)
    def _extract_pagewise(self, frames, queue_id, checksum, pms_mode, coordinate_format, crop=False):
        from .ingest import crop_to_content
        results = []
        for i, img in enumerate(frames):
            if crop:
                img = crop_to_content(img)
                h, w = img.shape[:2]
                overlay = np.full((h + 8, w + 8, 3), 255, np.uint8)
                overlay[4:h + 4, 4:w + 4] = img
                img = overlay
            boxes, fragments, lines, _, line_bboxes = self.box_processor.extract_bounding_boxes(queue_id, checksum, img, pms_mode)
            result, _ = self.icr_processor.recognize(queue_id, checksum, img, boxes, fragments, lines)
            self._finish(result, i, lines, line_bboxes, coordinate_format)
            results.append(result)
        return results

    def _extract_batched(self, frames, pms_mode, coordinate_format):
        import queue
        import threading
        pipe = self.box_processor.pipeline
        icr = self.icr_processor
        refiner = bool(getattr(self.box_processor, "line_refiner", False))
        line_boxes = []
        # Host post-processing (records -> words -> the reference's page records) runs on a helper thread, one decode
        # batch behind the GPU: only the last batch's share is left once the device is done.
        blocks = queue.Queue()
        state = {"words": [], "counts": None, "results": [], "done": 0, "error": None}

        def assemble_ready(final=False):
            counts = state["counts"]
            while counts is not None and len(state["results"]) < len(counts):
                i = len(state["results"])
                c = counts[i]
                if state["done"] + c > len(state["words"]) or (refiner and not final):
                    return                               # page i is not complete yet (line boxes arrive at the end)
                page_words = state["words"][state["done"]:state["done"] + c]
                state["done"] += c
                img = frames[i]
                meta = {"imageSize": {"width": img.shape[1], "height": img.shape[0]}, "page": 0, "lang": "en"}
                if not page_words:
                    result, lines = {"meta": meta, "words": [], "lines": []}, []
                else:
                    boxes = [w["box"] for w in page_words]
                    lines = [w["line"] for w in page_words]
                    res = [{"confidence": w["confidence"], "id": f"img-{j}", "text": w["text"]} for j, w in enumerate(page_words)]
                    result = assemble_result(meta, boxes, lines, res)
                self._finish(result, i, lines, line_boxes[i] if refiner else [], coordinate_format)
                state["results"].append(result)

        def worker():
            try:
                while True:
                    item = blocks.get()
                    if item is None:
                        return
                    block, counts = item
                    state["counts"] = counts
                    state["words"].extend(records_to_words(block, icr.detok))
                    assemble_ready()
            except Exception as ex:                      # surfaced on the caller's thread below
                state["error"] = ex

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        try:
            _, counts = pipe.run_frames(frames, preset=PSM_PRESETS[pms_mode.value], beam=icr.beam, max_len_b=icr.max_len_b,
                                        out_ld=getattr(self, "record_tokens", None) or icr.max_len_b + 1,
                                        line_refiner=refiner, want_lines=line_boxes, sink=lambda b, c: blocks.put((b, c)))
        finally:
            blocks.put(None)
            th.join()
        if state["error"] is not None:
            raise state["error"]
        state["counts"] = counts
        assemble_ready(final=True)
        return state["results"]

    # region / field extraction (__process_extract_regions, ocr_engine.py:223-414): every region is cropped, padded
    # with 4 white pixels, run through the box processor in its own PSM (cached by content), and the fragments of a
    # page go through ONE recognize() call.  Quirks kept as they are in the reference: region ids are recorded before
    # the zero-size / out-of-bounds checks (so a skipped region makes the page fall back to empty results), the
    # box-result list is not reset between pages, and words are matched to region ids in their x-sorted order.
    def _extract_regions(self, frames, queue_id, checksum, pms_mode, regions):
        output, extended = [], []
        for region in regions:
            if not all(k in region for k in ("id", "pageIndex", "x", "y", "w", "h")):
                raise Exception(f"Required key missing in region : {region}")
        pages = {}
        for region in regions:
            pages.setdefault(region["pageIndex"], []).append(region)

        def overlay_of(img, region):
            x, y, w, h = region["x"], region["y"], region["w"], region["h"]
            if w == 0 or h == 0 or y + h > img.shape[0] or x + w > img.shape[1]:
                return None
            overlay = np.full((h + 8, w + 8, 3), 255, np.uint8)
            overlay[4:h + 4, 4:w + 4] = img[y:y + h, x:x + w]
            mode = PSMode.from_value(region["mode"]) if "mode" in region else pms_mode
            return overlay, mode, (mode.value, overlay.shape, hash_frames_fast([overlay]))

        # pass 1: every overlay that still needs detection, grouped by (mode, shape) -> one batched detector pass per group
        todo = {}
        for page_index, page_regions in pages.items():
            for region in page_regions:
                ov = overlay_of(frames[page_index], region)
                if ov is not None and ov[2] not in self.bbox_cache and ov[1] not in (PSMode.WORD, PSMode.RAW_LINE):
                    todo.setdefault((ov[1], ov[0].shape), {})[ov[2]] = ov[0]
        batch_fn = getattr(self.box_processor, "extract_bounding_boxes_batch", None)
        if batch_fn is not None:
            for (mode, _), group in todo.items():
                keys = list(group)
                for k, res in zip(keys, batch_fn(queue_id, checksum, [group[k] for k in keys], psm=mode)):
                    self.bbox_cache[k] = res

        # pass 2: the reference's loop (results of pass 1 are found in the cache)
        bbox_results_batch = []
        for page_index, page_regions in pages.items():
            img = frames[page_index]
            xb, yb, wb, hb = img.shape[1], img.shape[0], 0, 0
            region_ids = []
            for region in page_regions:
                rid = region["id"]
                region_ids.append(rid)
                x, y, w, h = region["x"], region["y"], region["w"], region["h"]
                ov = overlay_of(img, region)
                if ov is None:
                    output.append({"id": rid, "text": "", "confidence": 0.0})
                    continue
                xb, yb = min(x, xb), min(y, yb)
                wb = max(x + w, xb + wb) - xb
                hb = max(y + h, hb + yb) - yb
                overlay, mode, key = ov
                if key not in self.bbox_cache:
                    self.bbox_cache[key] = self.box_processor.extract_bounding_boxes(queue_id, checksum, overlay, psm=mode)
                bbox_results_batch.append(self.bbox_cache[key])
            batch_crop = img[yb:yb + hb, xb:xb + wb]
            boxes, fragments, lines, _, _ = (list(chain.from_iterable(x)) for x in zip(*bbox_results_batch))
            batch_result, _ = self.icr_processor.recognize(queue_id, checksum, batch_crop, boxes, fragments, lines)
            extended.append(batch_result)
            if "words" in batch_result and len(batch_result["words"]) == len(region_ids):
                for word, rid in zip(batch_result["words"], region_ids):
                    output.append({"id": rid, "text": word["text"], "confidence": word["confidence"]})
            else:
                for rid in region_ids:
                    output.append({"id": rid, "text": "", "confidence": 0.0})
        return {"regions": output, "extended": extended}

    @staticmethod
    def _finish(result, page, lines, line_bboxes, coordinate_format):
        if coordinate_format == CoordinateFormat.XYXY:
            for word in result["words"]:
                x, y, w, h = word["box"]
                word["box"] = [x, y, x + w, y + h]
        result["meta"]["page"] = page
        result["meta"]["lines"] = lines
        result["meta"]["lines_bboxes"] = line_bboxes
        result["meta"]["format"] = coordinate_format.name.lower()
