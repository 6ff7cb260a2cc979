"""CPU restatement of the line grouping helpers (TEST INFRASTRUCTURE — see oracle/__init__.py).

  vertical_overlaps   find_overlap_vertical        marie/utils/overlap.py:42-103
  find_line_number    find_line_number             marie/boxes/line_processor.py:15-45
  line_merge          line_merge / __line_merge    marie/boxes/line_processor.py:48-171
  merge_block         merge_bboxes_as_block        marie/utils/overlap.py:186-204

Written against the semantics (visit order, strict comparisons, float64 IoU) rather than the text of the reference;
pinned against the reference's own functions in tests/test_oracle_vs_reference.py and tests/golden/lines_*.json.
"""
import numpy as np


def vertical_overlaps(box, data):
    """Indices (ascending) and vertical IoU of every entry of `data` whose y-extent strictly overlaps `box`'s,
    skipping zero-height boxes and exact duplicates of `box`.  IoU = inter / (h + h' - inter), clamped to [0, 1]."""
    data = np.asarray(data, dtype=np.int64).reshape(-1, 4)
    if len(data) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float64)
    x, y, w, h = (int(v) for v in box)
    if h <= 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float64)
    ys, hs = data[:, 1], data[:, 3]
    same = (data[:, 0] == x) & (ys == y) & (data[:, 2] == w) & (hs == h)
    hit = (hs > 0) & ~same & (y < ys + hs) & (ys < y + h)
    idx = np.flatnonzero(hit)
    inter = np.minimum(y + h, ys[idx] + hs[idx]) - np.maximum(y, ys[idx])
    iou = inter / (h + hs[idx] - inter).astype(np.float64)
    return idx, np.clip(iou, 0.0, 1.0)


def find_line_number(lines, box):
    """1-based index of the line with the best vertical IoU; with no overlap the line whose bottom edge is closest
    to the box's vertical centre; -1 when there are no lines."""
    idx, iou = vertical_overlaps(box, lines)
    if len(idx) == 1:
        return int(idx[0]) + 1
    if len(idx) > 1:
        best, line = 0, -1
        for i, s in zip(idx, iou):      # strict '>' keeps the first maximum
            if s > best:
                best, line = s, int(i) + 1
        if line != -1:
            return line
    line, best = -1, 100000
    cy = box[1] + box[3] // 2
    for i, ln in enumerate(lines):
        d = abs(cy - (ln[1] + ln[3]))
        if d < best:
            line, best = i + 1, d
    return line


def _merge_pass(boxes, min_iou, kind=None):
    boxes = np.asarray(boxes, dtype=np.int64).reshape(-1, 4)
    boxes = boxes[np.argsort(boxes[:, 1], kind=kind)]   # kind=None: default (quicksort) argsort, like the reference
    n = len(boxes)
    visited = np.zeros(n, bool)
    out = []
    for i in range(n):
        if visited[i]:
            continue
        visited[i] = True
        idx, iou = vertical_overlaps(boxes[i], boxes)
        group = [i]
        for j, s in zip(idx, iou):
            if visited[j] or s < min_iou:
                continue
            back, _ = vertical_overlaps(boxes[j], boxes)
            if len(back) == len(idx):               # "the ray back is valid": same number of overlaps both ways
                group.append(int(j))
                visited[j] = True
        g = boxes[group]
        x0 = g[:, 0].min()
        out.append([x0, g[:, 1].min(), (g[:, 0] + g[:, 2]).max() - x0, g[:, 3].max()])   # height = max h, not union
    return out


def line_merge(bboxes, kind=None):
    """Seven merge passes at decreasing IoU (early stop after 3 passes without change), containment prune, y-sort.
    kind=None reproduces the reference call for call (numpy's default argsort, whose order among equal y is
    platform-dependent); kind="stable" is the deterministic tie rule the product implements."""
    if len(bboxes) == 0:
        return []
    merged = [list(b) for b in bboxes]
    unchanged = 0
    for thr in (0.8, 0.7, 0.6, 0.5, 0.4, 0.37, 0.35):
        before = len(merged)
        merged = _merge_pass(merged, thr, kind)
        if len(merged) == before:
            unchanged += 1
            if unchanged > 2:
                break
    m = np.asarray(merged, dtype=np.int64).reshape(-1, 4)
    x0, y0, x1, y1 = m[:, 0], m[:, 1], m[:, 0] + m[:, 2], m[:, 1] + m[:, 3]
    inside = (x0[None] > x0[:, None]) & (x1[None] < x1[:, None]) & (y0[None] > y0[:, None]) & (y1[None] < y1[:, None])
    m = m[~inside.any(0)]
    return m[np.argsort(m[:, 1], kind=kind)]


def merge_block(bboxes):
    b = np.asarray(bboxes)
    x0, y0 = b[:, 0].min(), b[:, 1].min()
    return [round(k, 6) for k in [x0, y0, (b[:, 0] + b[:, 2]).max() - x0, (b[:, 1] + b[:, 3]).max() - y0]]
