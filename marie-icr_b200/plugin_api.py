"""Mirror of the reference's plugin interfaces for this path, so the B200 processors are drop-ins:

  PSMode              marie/boxes/box_processor.py:129-162
  BoxProcessor        marie/boxes/box_processor.py:180-256        (constructor, extract_bounding_boxes, psm_*)
  OcrProcessor        marie/document/ocr_processor.py:34-267      (is_available, recognize_from_fragments, recognize)
  CoordinateFormat    marie/ocr/coordinate_format.py:6-60

When the real `marie` package is importable the processors can be registered with it directly; the
classes here carry the same names, argument meaning, return shapes and error behaviour, so the parity tests read like
the reference's own integration scripts (tests/integration/test_icr.py).
"""
import os
from abc import ABC, abstractmethod
from enum import Enum

import numpy as np

# marie/constants.py:91-96: __model_path__ = <MARIE_DEFAULT_MOUNT or the directory above the package>/model_zoo
MODEL_PATH = os.path.join(os.environ.get("MARIE_DEFAULT_MOUNT", os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))),
                          "model_zoo")


class PSMode(Enum):
    WORD = "word"
    SPARSE = "sparse"
    LINE = "line"
    RAW_LINE = "raw_line"
    MULTI_LINE = "multiline"

    @staticmethod
    def from_value(value):
        if value is None:
            return PSMode.SPARSE
        for m in PSMode:
            if m.value == value.lower():
                return m
        return PSMode.SPARSE


class CoordinateFormat(Enum):
    XYWH = "xywh"
    XYXY = "xyxy"

    @staticmethod
    def from_value(value):
        if value is None:
            return CoordinateFormat.XYWH
        for m in CoordinateFormat:
            if m.value == value.lower():
                return m
        return CoordinateFormat.XYWH

    @staticmethod
    def convert(box, from_mode, to_mode):
        arr = np.array(box)
        assert arr.shape == (4,), "CoordinateFormat.convert takes either a 4-tuple/list"
        if from_mode == to_mode:
            return box
        kind = type(box)
        arr = arr.reshape(-1, 4)
        if to_mode == CoordinateFormat.XYXY and from_mode == CoordinateFormat.XYWH:
            arr[:, 2:] += arr[:, :2]
        elif from_mode == CoordinateFormat.XYXY and to_mode == CoordinateFormat.XYWH:
            arr[:, 2:] -= arr[:, :2]
        else:
            raise RuntimeError("Cannot be here!")
        return kind(arr.flatten())


class BoxProcessor(ABC):
    """Box processor: extracts bounding boxes (box_processor.py:180-256)."""

    def __init__(self, work_dir="/tmp/boxes", models_dir="./models", cuda=False, config=None):
        self.cuda = cuda
        self.work_dir = work_dir

    @abstractmethod
    def extract_bounding_boxes(self, _id, key, img, psm=PSMode.SPARSE):
        """-> (boxes [x,y,w,h], fragments, line_numbers, prediction_result{bboxes,polys,heatmap}, lines_bboxes)"""

    @abstractmethod
    def psm_word(self, image): ...

    @abstractmethod
    def psm_sparse(self, image): ...

    @abstractmethod
    def psm_line(self, image): ...

    @abstractmethod
    def psm_raw_line(self, image): ...

    @abstractmethod
    def psm_multiline(self, image): ...


class OcrProcessor(ABC):
    """Base class of the recognisers (ocr_processor.py:34-267)."""

    def __init__(self, work_dir="/tmp/icr", cuda=True, **kwargs):
        self.cuda = cuda
        self.work_dir = work_dir

    @abstractmethod
    def is_available(self) -> bool: ...

    def extract_text(self, _id, key, image):
        results = self.recognize_from_fragments([image])
        if len(results) == 1:
            return results[0]["text"], results[0]["confidence"]
        return None, 0

    def recognize_from_boxes(self, image, boxes, **kwargs):
        raise Exception("Not yet implemented")

    def recognize_from_fragments(self, image_fragments):
        raise Exception("Not Implemented")

    def recognize(self, _id, key, img, boxes, fragments, lines, return_overlay=False):
        """Result assembly of ocr_processor.py:87-267: words re-indexed by x, grouped into lines by line id."""
        if img is None:
            raise Exception("Input image can't be empty")
        if not isinstance(img, np.ndarray):
            try:
                from PIL import Image
                if isinstance(img, Image.Image):
                    img = np.array(img)[:, :, ::-1].copy()
            except ImportError:
                pass
        if not isinstance(img, np.ndarray):
            raise Exception("Expected image in numpy format but got {}".format(type(img)))
        assert len(boxes) == len(fragments), "You must provide the same number of box groups as images."
        assert len(boxes) == len(lines), "You must provide the same number of lines as boxes."
        meta = {"imageSize": {"width": img.shape[1], "height": img.shape[0]}, "page": 0, "lang": "en"}
        if len(boxes) == 0:
            return {"meta": meta, "words": [], "lines": []}, np.ones((img.shape[0], img.shape[1], 3), np.uint8) * 255
        results = self.recognize_from_fragments(fragments)
        assert len(results) == len(fragments), "You must provide the same number of results as fragments."
        return assemble_result(meta, boxes, lines, results), None


def assemble_result(meta, boxes, lines, results):
    """words sorted by x (numpy argsort, as the reference), `id` = rank, confidence rounded to 3; lines in ascending
    line id with word ids, joined text, merged bbox (merge_bboxes_as_block, marie/utils/overlap.py:186-204) and mean
    confidence rounded to 4 (ocr_processor.py:161-253)."""
    boxes = np.array(boxes)
    lines = np.array(lines)
    order = np.argsort(boxes[:, 0])
    words = []
    for i, index in enumerate(order):
        r = results[index]
        words.append({"id": i, "text": r["text"], "confidence": round(r["confidence"], 3), "box": boxes[index],
                      "line": lines[index]})
    unique_ids = sorted(np.unique(lines))
    line_results = np.empty(len(unique_ids), dtype=object)
    aligned, word_index = [], 0
    for i, line_id in enumerate(unique_ids):
        picks = [w for w in words if w["line"] == line_id]
        if not picks:
            raise Exception("Every word needs to be associated with a box")
        for w in picks:
            w["word_index"] = word_index
            word_index += 1
            aligned.append(w)
        b = np.array([w["box"] for w in picks])
        x0, y0 = b[:, 0].min(), b[:, 1].min()
        bbox = [round(k, 6) for k in [x0, y0, (b[:, 0] + b[:, 2]).max() - x0, (b[:, 1] + b[:, 3]).max() - y0]]
        line_results[i] = {"line": i + 1, "wordids": [w["id"] for w in picks], "text": " ".join(w["text"] for w in picks),
                           "bbox": bbox, "confidence": round(np.average([w["confidence"] for w in picks]), 4)}
    if len(words) != len(aligned):
        raise Exception(f"Aligned words should match original words got: {len(aligned)}, {len(words)}")
    return {"meta": meta, "words": aligned, "lines": line_results}
