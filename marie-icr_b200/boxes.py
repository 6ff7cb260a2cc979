"""BoxProcessorCraftB200 — drop-in for BoxProcessorCraft (marie/boxes/craft_box_processor.py:244-562): same
constructor, PSM dispatch, thresholds, return tuple and error behaviour; detection runs on the B200 path
(K1 -> K2-K4 -> K5-K7).  The reference's unconditional debug writes (crop JPEGs, overlay PNG/JPEG, score-map PNGs;
:100,533-550 and craft_utils.py:40-43) are not reproduced.
"""
import copy
import os

import numpy as np
import torch

from . import weights as _weights
from .pipeline import PSM_PRESETS, PagePipeline
from .plugin_api import MODEL_PATH, BoxProcessor, PSMode


class BoxProcessorCraftB200(BoxProcessor):
    def __init__(self, work_dir="/tmp/boxes", models_dir=os.path.join(MODEL_PATH, "craft"), cuda=True, config=None, *,
                 state_dict=None, pipeline=None, device=0, line_refiner_state_dict=None, poly=False):
        """models_dir is the CRAFT directory itself, as in the reference (default `<model_zoo>/craft`,
        craft_box_processor.py:245-249).  state_dict: CRAFT weights (keys of marie/models/craft/craft.py); when omitted
        the reference's checkpoint `<models_dir>/craft_mlt_25k.pth` is loaded (:260-277).
        line_refiner_state_dict: RefineNet weights (marie/models/craft/refinenet.py) — enables the line branch of
        get_prediction (craft_box_processor.py:150-217), which the reference keeps switched off (`:287-312`: the
        refiner is never loaded, so `lines_bboxes` is always [] and every box gets line -1).  With it, `lines_bboxes`
        and the per-box line numbers are produced exactly as that branch would.
        poly: the `poly` argument of get_prediction (craft_box_processor.py:76-135), False in every preset of the
        reference; True adds the polygon refinement of getPoly_core to `prediction_result["polys"]` (host routine over
        the device's label map, polys.py) — boxes, rects and fragments do not depend on it."""
        super().__init__(work_dir, models_dir, cuda, config or {})
        if not cuda:
            raise RuntimeError("BoxProcessorCraftB200 has no CPU path: cuda=True and a B200 are required")
        self.pipeline = pipeline or PagePipeline(device=device)
        if state_dict is None and not self.pipeline.has_craft:
            path = os.path.join(models_dir, "craft_mlt_25k.pth")
            if not os.path.exists(path):
                raise FileNotFoundError(f"CRAFT checkpoint not found: {path}")
            state_dict = torch.load(path, map_location="cpu", weights_only=True)      # a plain tensor dict
        if state_dict is not None:
            self.pipeline.load_craft(_weights.pack_craft(state_dict, self.pipeline.dtype))
        self.device = f"cuda:{self.pipeline.device}"
        self.poly = bool(poly)
        self.line_refiner = line_refiner_state_dict is not None
        if self.line_refiner:
            from . import ops as _ops
            _ops.load_refine(_weights.pack_refine(line_refiner_state_dict, self.pipeline.dtype), self.pipeline.device)

    def unload(self):
        from ._lib import Context
        Context.get(self.pipeline.device).close()

    # ------------------------------------------------------------------ PSM presets (get_prediction, :76-146)
    def _predict(self, image, mode):
        pages = torch.from_numpy(np.ascontiguousarray(image[None])).to(self.device)
        det = self.pipeline.detect(pages, PSM_PRESETS[mode], line_refiner=self.line_refiner, keep_post=self.poly)
        bboxes = det["boxes"].cpu().numpy()
        if self.poly:
            from .polys import adjust_polys, get_poly_core
            nb = det["counts"][0]
            post = det["post"]
            raw = get_poly_core(list(post["det"][0, :nb].cpu().numpy()), post["labels"][0].cpu().numpy(),
                                post["mapper"][0, :nb].cpu().numpy())
            polys = adjust_polys(raw, [b for b in bboxes], 1.0 / det["ratio"], 1.0 / det["ratio"])
        else:
            polys = [b for b in bboxes]                  # poly=False: polys[k] = boxes[k] (:133-135)
        self._last_rects = det["rects"].cpu().numpy()
        return bboxes, polys, None, det.get("lines", [[]])[0]

    def psm_word(self, image):
        return self._predict(image, "word")

    def psm_sparse(self, image):
        return self._predict(image, "sparse")

    def psm_line(self, image):
        return self._predict(image, "line")

    def psm_raw_line(self, image):
        return self._predict(image, "raw_line")

    def psm_multiline(self, image):
        return self._predict(image, "multiline")

    def extract_bounding_boxes_batch(self, _id, key, images, psm=PSMode.SPARSE):
        """extract_bounding_boxes for several images of ONE shape in one pass of the batched kernels (K1 / CRAFT in
        micro-batches, one K5-K7 launch set for all): the region path of the engine (ocr_engine.py:223-414) groups its
        padded region overlays by (mode, shape) and calls this once per group.  Returns one result tuple per image,
        identical to what extract_bounding_boxes returns for it."""
        if not len(images):
            return []
        if psm in (PSMode.RAW_LINE, PSMode.WORD) or self.line_refiner or len({im.shape for im in images}) != 1:
            return [self.extract_bounding_boxes(_id, key, im, psm) for im in images]
        if psm not in (PSMode.SPARSE, PSMode.LINE, PSMode.MULTI_LINE):
            raise Exception(f"PSM mode not supported : {psm}")
        pages = torch.from_numpy(np.ascontiguousarray(np.stack(images))).to(self.device)
        det = self.pipeline.detect(pages, PSM_PRESETS[psm.value])
        rects_all, boxes_all = det["rects"].cpu().numpy(), det["boxes"].cpu().numpy()
        out, k = [], 0
        for image, c in zip(images, det["counts"]):
            rects, bboxes = rects_all[k:k + c], boxes_all[k:k + c]
            k += c
            fragments = [image[y:y + h + 1, x:x + w + 1].copy() for x, y, w, h in rects.tolist()]
            out.append(([list(r) for r in rects.tolist()], fragments, [-1] * c,
                        {"bboxes": bboxes, "polys": [b for b in bboxes], "heatmap": None}, []))
        return out

    # ------------------------------------------------------------------ extract_bounding_boxes (:431-562)
    def extract_bounding_boxes(self, _id, key, img, psm=PSMode.SPARSE):
        if img is None:
            raise Exception("Input image can't be empty")
        image = img
        lines_bboxes = []
        if psm == PSMode.SPARSE:
            bboxes, polys, score_text, lines_bboxes = self.psm_sparse(image)
        elif psm == PSMode.LINE:
            bboxes, polys, score_text, lines_bboxes = self.psm_line(image)
        elif psm == PSMode.MULTI_LINE:
            bboxes, polys, score_text, lines_bboxes = self.psm_multiline(image)
        elif psm == PSMode.RAW_LINE or psm == PSMode.WORD:
            h, w = image.shape[0], image.shape[1]
            return [[0, 0, w, h]], [copy.deepcopy(image)], [0], dict(), lines_bboxes
        else:
            raise Exception(f"PSM mode not supported : {psm}")
        prediction_result = {"bboxes": bboxes, "polys": polys, "heatmap": score_text}
        rects = self._last_rects
        rect_from_poly, fragments, line_numbers = [], [], []
        if len(lines_bboxes):
            from . import lines as _lines
            ids = _lines.find_line_numbers(lines_bboxes, rects)
        for i, (x, y, w, h) in enumerate(rects.tolist()):
            rect_from_poly.append([x, y, w, h])
            fragments.append(image[y:y + h + 1, x:x + w + 1].copy())      # crop_poly_low on the expanded rect (:42-73,524)
            line_numbers.append(int(ids[i]) if len(lines_bboxes) else -1)
        return rect_from_poly, fragments, line_numbers, prediction_result, lines_bboxes
