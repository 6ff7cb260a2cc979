"""Page ingest helpers that sit just before the hot path — host-side, size arithmetic first.

  ensure_max_page_size   marie/utils/image_utils.py:254-321 (known answers: tests/imaging/test_image_resizing.py:7-47)
  hash_frames_fast       marie/utils/image_utils.py:136-149 (md5 over all pixels, the engine's cache key)
  crop_to_content        marie/utils/image_utils.py:190-251 (the engine's optional `crop_to_content=True` pre-step,
                         marie/ocr/ocr_engine.py:169-176; the debug PNG writes are not reproduced)
"""
import hashlib

import numpy as np


def max_page_dims(width, height, max_page_size=(2550, 3300), expand_ratio=0.15):
    """New (width, height) for a frame, or None when it already fits.  Landscape frames swap the limits; limits are
    expanded by int(limit * expand_ratio); aspect ratio is width / height in float64 with int() truncation."""
    mw, mh = max_page_size
    if width > height:
        mw, mh = mh, mw
    mw, mh = mw + int(mw * expand_ratio), mh + int(mh * expand_ratio)
    if not (width > mw or height > mh):
        return None
    aspect = width / height
    if width > height:
        nw = min(width, mw)
        nh = int(nw / aspect)
        if nh > mh:
            nh = mh
            nw = int(nh * aspect)
    else:
        nh = min(height, mh)
        nw = int(nh * aspect)
        if nw > mw:
            nw = mw
            nh = int(nw / aspect)
    return nw, nh


def ensure_max_page_size(frames, max_page_size=(2550, 3300), expand_ratio=0.15):
    """-> (changed, frames): frames exceeding the (expanded) page size are shrunk with cv2.INTER_AREA."""
    import cv2
    out, changed = [], False
    for frame in frames:
        h, w = frame.shape[:2]
        dims = max_page_dims(w, h, max_page_size, expand_ratio)
        if dims is None:
            out.append(frame)
        else:
            changed = True
            out.append(cv2.resize(frame, dims, interpolation=cv2.INTER_AREA))
    return changed, out


def hash_frames_fast(frames, blocksize=2 ** 20):
    md5 = hashlib.md5()
    for frame in frames:
        buf = np.ravel(frame)
        for s in range(0, len(buf), blocksize):
            md5.update(buf[s:s + blocksize])
    return md5.hexdigest()


def crop_to_content(frame, content_aware=True):
    """Crops a page to its content (first / last non-background pixel after Otsu binarisation; `content_aware`: division
    normalisation + 2x3 closing first, horizontal crop only, 16 px of side padding)."""
    import cv2
    gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if frame.ndim == 3 and frame.shape[2] == 3 else frame
    if content_aware:
        blur = cv2.GaussianBlur(gray, (5, 5), sigmaX=0, sigmaY=0)
        divide = cv2.divide(gray, blur, scale=255)
        thresh = cv2.threshold(divide, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[1]
        op_frame = cv2.morphologyEx(thresh, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (2, 3)))
    else:
        op_frame = cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[1]
    ys, xs = np.where(op_frame == 0)
    img_h, img_w = op_frame.shape[:2]
    if len(ys) == 0:
        return frame
    if content_aware:
        x = max(0, xs.min() - 16)
        y, h = 0, img_h
        w = min(img_w, xs.max() - x + 16)
    else:
        x, y = xs.min(), ys.min()
        h, w = ys.max() - y, xs.max() - x
    return frame[y:y + h + 1, x:x + w + 1].copy()
