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


# ------------------------------------------------------------------------------------------ frames from files (rank 4)
def convert_frames(frames, img_format="cv"):
    """marie/utils/docs.py:183-198 — every frame becomes 3-channel: GRAY -> RGB, 3-channel frames go through
    COLOR_BGR2RGB (the reference swaps the channel order of what cv2 decoded here; kept as it is)."""
    import cv2
    out = []
    for frame in frames:
        if isinstance(frame, np.ndarray):
            conv = cv2.cvtColor(frame, cv2.COLOR_GRAY2RGB) if frame.ndim == 2 else cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
            if img_format == "pil":
                from PIL import Image
                conv = Image.fromarray(frame.copy())
            out.append(conv)
        else:
            out.append(frame.copy())
    return out


def document_type(path):
    """file type by magic bytes (marie/utils/docs.py:28-52 uses imghdr + PyPDF4): tiff / png / jpeg / bmp / pdf"""
    with open(path, "rb") as f:
        head = f.read(16)
    if head[:4] in (b"II*\x00", b"MM\x00*"):
        return "tiff"
    if head[:8] == b"\x89PNG\r\n\x1a\n":
        return "png"
    if head[:3] == b"\xff\xd8\xff":
        return "jpeg"
    if head[:2] == b"BM":
        return "bmp"
    if head[:5] == b"%PDF-":
        return "pdf"
    raise Exception("Unsupported file type, expected one of : tiff, png, jpeg, bmp, pdf")


def load_image(img_path, img_format="cv"):
    """marie/utils/docs.py:201-256 — (loaded, frames): a TIFF is read as a multi-page document (one frame per page,
    cv2.imreadmulti + convert_frames), any other raster image is one RGB frame; PDF page rasterisation needs PyPDF4
    (not available here) and raises."""
    import os
    import cv2
    if img_path is None:
        return False, None
    if not os.path.exists(img_path):
        raise Exception(f"File not found : {img_path}")
    kind = document_type(img_path)
    if kind == "pdf":
        raise NotImplementedError("PDF frames need PyPDF4 (marie/utils/docs.py:108-180); convert the document to TIFF first")
    if kind == "tiff":
        loaded, frames = cv2.imreadmulti(img_path, [], cv2.IMREAD_ANYCOLOR)
        if not loaded:
            return False, []
        return True, convert_frames(frames, img_format)
    from PIL import Image
    img = np.array(Image.open(img_path).convert("RGB"), dtype=np.uint8)
    if img_format == "pil":
        return True, [Image.fromarray(cv2.cvtColor(img, cv2.COLOR_BGR2RGB).copy())]
    return True, [img]


def frames_from_file(img_path):
    """marie/utils/docs.py:372-379"""
    import os
    if not os.path.exists(img_path):
        raise FileNotFoundError(f"File not found : {img_path}")
    loaded, frames = load_image(img_path)
    if not loaded:
        raise Exception(f"Unable to load image : {img_path}")
    return frames


def burst_frames(ref_id, frames, root_asset_dir, force=False, bitonal=True):
    """marie/pipe/components.py:529-565 + marie/utils/tiff_ops.py:73-160: one TIFF per page under
    `<root_asset_dir>/burst/<prefix>_<page:05>.<suffix>` (pages numbered from 1), skipped when the directory already
    holds that many files.  Bitonal pages are Otsu-thresholded and written as 1-bit CCITT Group 4 at 300 DPI (the
    reference goes through tifffile + ImageMagick for the same result).  Returns the file paths."""
    import os
    import cv2
    from PIL import Image
    out_dir = os.path.join(root_asset_dir, "burst")
    os.makedirs(out_dir, exist_ok=True)
    filename = ref_id.split("/")[-1]
    prefix, suffix = filename.split(".")[0], filename.split(".")[-1]
    names = [os.path.join(out_dir, f"{prefix}_{i + 1:05}.{suffix}") for i in range(len(frames))]
    if not force and len([f for f in os.listdir(out_dir) if os.path.isfile(os.path.join(out_dir, f))]) == len(frames):
        return names
    for frame, path in zip(frames, names):
        if bitonal:
            if frame.ndim == 3:
                frame = cv2.threshold(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), 0, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)[1]
            Image.fromarray(frame).convert("1").save(path, format="TIFF", compression="group4", dpi=(300, 300))
        else:
            Image.fromarray(frame).save(path, format="TIFF", dpi=(300, 300))
    return names
