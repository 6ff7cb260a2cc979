"""Host bindings of the line-grouping entry points (csrc/lines.cu): line_merge / find_line_number of
marie/boxes/line_processor.py:15-171."""
import ctypes

import numpy as np

from ._lib import load_library


def line_merge(bboxes):
    """[[x,y,w,h]...] -> ndarray [m,4] of merged line boxes, y-sorted."""
    b = np.ascontiguousarray(np.asarray(bboxes, dtype=np.int32).reshape(-1, 4))
    if len(b) == 0:
        return []
    out = np.zeros_like(b)
    n = ctypes.c_int(0)
    rc = load_library().mb_line_merge(None, b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(b)),
                                      out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n))
    if rc != 0:
        raise RuntimeError(f"mb_line_merge failed ({rc})")
    return out[:n.value]


def find_line_numbers(lines, boxes):
    l = np.ascontiguousarray(np.asarray(lines, dtype=np.int32).reshape(-1, 4))
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.int32).reshape(-1, 4))
    out = np.zeros(len(b), np.int32)
    rc = load_library().mb_find_line_numbers(None, l.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(l)),
                                             b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(b)),
                                             out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(f"mb_find_line_numbers failed ({rc})")
    return out


def find_line_number(lines, box):
    return int(find_line_numbers(lines, [box])[0])
