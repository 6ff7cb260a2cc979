"""CPU restatements of the two image resamplers on the hot path (TEST INFRASTRUCTURE — see oracle/__init__.py).

  pil_bicubic_resize_u8   PIL.Image.resize((384,384), Image.BICUBIC) on RGB u8, as called by preprocess_image
                          (marie/document/trocr_ocr_processor.py:116-125).  Restates Pillow's two-pass antialiased
                          resampler (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
                          ImagingResampleHorizontal_8bpc / Vertical_8bpc): double-precision bicubic (a = -0.5)
                          coefficients, support scaled by the down-scale factor, 22-bit fixed point, u8 intermediate.
  trocr_normalize         transforms.ToTensor + Normalize(0.5, 0.5) (trocr_ocr_processor.py:95-101) in float32.
  cv_linear_resize_u8     cv2.resize(..., INTER_LINEAR) on u8 as called by resize_aspect_ratio
                          (marie/models/craft/imgproc.py:45-58): 11-bit fixed-point coefficients, two passes.
Both are pinned against the libraries themselves (PIL 12.2 / cv2 4.13 in this image) in tests/test_oracle_cpu.py.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size).  Returns (bounds [out,2], kk [out,ksize] i32)."""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_bicubic_resize_u8(img, out_w=384, out_h=384):
    """img: [H, W, C] u8 -> [out_h, out_w, C] u8, bit-exact with PIL's BICUBIC resize."""
    h, w, c = img.shape
    src = img.astype(np.int64)
    bv, kv = pil_coeffs(h, out_h)
    if out_w != w:
        bh, kh = pil_coeffs(w, out_w)
        y_first = int(bv[0, 0])
        y_last = int(bv[-1, 0] + bv[-1, 1])
        tmp = np.zeros((y_last - y_first, out_w, c), np.uint8)
        for xx in range(out_w):
            xmin, n = bh[xx]
            acc = (src[y_first:y_last, xmin:xmin + n, :] * kh[xx, :n][None, :, None].astype(np.int64)).sum(1)
            tmp[:, xx, :] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        src = tmp.astype(np.int64)
        bv = bv.copy()
        bv[:, 0] -= y_first
    if out_h != h:
        out = np.zeros((out_h, src.shape[1], c), np.uint8)
        for yy in range(out_h):
            ymin, n = bv[yy]
            acc = (src[ymin:ymin + n] * kv[yy, :n][:, None, None].astype(np.int64)).sum(0)
            out[yy] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        return out
    return src.astype(np.uint8)


def trocr_normalize(u8_hwc_rgb):
    """ToTensor (/255 in float32) then Normalize(0.5, 0.5): returns [C, H, W] float32."""
    t = u8_hwc_rgb.astype(np.float32) / np.float32(255)
    t = (t - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(t.transpose(2, 0, 1))


def fragment_to_input(fragment_bgr):
    """MemoryDataset.__getitem__ (BGR->RGB, marie/models/icr/memory_dataset.py:43-53) + preprocess_image
    (trocr_ocr_processor.py:116-125): [h, w, 3] u8 BGR -> [3, 384, 384] float32."""
    rgb = fragment_bgr[:, :, ::-1]
    return trocr_normalize(pil_bicubic_resize_u8(rgb, 384, 384))


def _cv_lin_coeffs(ssize, dsize):
    scale = ssize / dsize
    ofs = np.zeros(dsize, np.int64)
    a0 = np.zeros(dsize, np.int64)
    a1 = np.zeros(dsize, np.int64)
    for d in range(dsize):
        fx = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(fx))
        fx = np.float32(fx - np.float32(s))
        if s < 0:
            fx, s = np.float32(0), 0
        if s >= ssize - 1:
            fx, s = np.float32(0), ssize - 1
        ofs[d] = s
        a0[d] = int(np.rint(np.float32((np.float32(1.0) - fx) * np.float32(2048))))
        a1[d] = int(np.rint(np.float32(fx * np.float32(2048))))
    return ofs, a0, a1


def cv_linear_resize_u8(img, dw, dh):
    """cv2.resize(img, (dw, dh), interpolation=INTER_LINEAR) for u8 [H,W,C] (down-scale / identity, the only
    cases resize_aspect_ratio produces): 11-bit coefficients, int horizontal sums, cv2's vertical combine."""
    h, w, _ = img.shape
    sx, ax0, ax1 = _cv_lin_coeffs(w, dw)
    sy, ay0, ay1 = _cv_lin_coeffs(h, dh)
    src = img.astype(np.int64)
    rows = src[:, sx, :] * ax0[None, :, None] + src[:, np.minimum(sx + 1, w - 1), :] * ax1[None, :, None]
    s0, s1 = rows[sy], rows[np.minimum(sy + 1, h - 1)]
    b0, b1 = ay0[:, None, None], ay1[:, None, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def craft_input(page_bgr, canvas_size=None, mag_ratio=1.0):
    """resize_aspect_ratio + normalizeMeanVariance (imgproc.py:26-32,45-73): returns ([oh, ow, 3] float32, ratio)."""
    ph, pw, _ = page_bgr.shape
    canvas_size = pw if canvas_size is None else canvas_size
    target = mag_ratio * max(ph, pw)
    if target > canvas_size:
        target = canvas_size
    ratio = target / max(ph, pw)
    th, tw = int(ph * ratio), int(pw * ratio)
    proc = cv_linear_resize_u8(page_bgr, tw, th)
    oh = th if th % 32 == 0 else th + (32 - th % 32)
    ow = tw if tw % 32 == 0 else tw + (32 - tw % 32)
    canvas = np.zeros((oh, ow, 3), np.float32)
    canvas[:th, :tw] = proc
    canvas -= np.array([127.5, 127.5, 127.5], np.float32)
    canvas /= np.array([127.5, 127.5, 127.5], np.float32)
    return canvas, ratio
