"""Oracle self-checks that need no GPU and no reference tree: the resampler restatements against the libraries the
reference calls (cv2.resize INTER_LINEAR, PIL BICUBIC), the cv2-free geometry restatement against the cv2 flavour,
and the host build of the device geometry code (csrc/boxgeom.cuh compiled with g++) against both."""
import ctypes
import os

import cv2
import numpy as np
import pytest
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape", [(40, 120), (384, 384), (500, 37), (3, 5), (61, 900), (384, 100), (100, 384)])
def test_pil_bicubic_restatement(shape):
    from oracle import resample
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((384, 384), Image.BICUBIC))
    assert np.array_equal(resample.pil_bicubic_resize_u8(img), ref)


@pytest.mark.parametrize("ph,pw", [(330, 255), (200, 300), (97, 61), (640, 640)])
def test_cv_linear_restatement(ph, pw):
    from oracle import resample
    rng = np.random.default_rng(ph)
    img = rng.integers(0, 256, (ph, pw, 3), dtype=np.uint8)
    ratio = pw / max(ph, pw)
    th, tw = int(ph * ratio), int(pw * ratio)
    ref = cv2.resize(img, (tw, th), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(resample.cv_linear_resize_u8(img, tw, th), ref)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_restated_boxes_match_cv_flavour(seed):
    from oracle import craft_post, synth
    text, link = synth.random_score_maps(seed, 120, 200, n_blobs=30)
    a, la, ma = craft_post.det_boxes_cv(text, link, 0.7, 0.45, 0.3)
    b, lb, mb = craft_post.det_boxes_restated(text, link, 0.7, 0.45, 0.3)
    assert np.array_equal(la, lb) and ma == mb and len(a) > 5
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert np.abs(a - b).max() < 1e-3
    assert (a == b).reshape(len(a), -1).all(1).sum() >= len(a) - 1


def test_edge_maps():
    from oracle import craft_post
    h, w = 48, 64
    z = np.zeros((h, w), np.float32)
    det, labels, mapper = craft_post.det_boxes_cv(z, z, 0.7, 0.45, 0.3)
    assert det == [] and labels.max() == 0
    one = np.ones((h, w), np.float32)
    a, _, _ = craft_post.det_boxes_cv(one, z, 0.7, 0.45, 0.3)
    b, _, _ = craft_post.det_boxes_restated(one, z, 0.7, 0.45, 0.3)
    assert len(a) == 1 and np.array_equal(a[0], b[0])
    rects = craft_post.boxes_to_rects(craft_post.adjust_result_coordinates([a[0].copy()], 1.0, 1.0), 2 * h, 2 * w)
    assert rects == [[0, 0, 2 * w, 2 * h]]


def test_host_build_of_device_geometry():
    """csrc/boxgeom.cuh (the code the CUDA box kernel runs) compiled for the host must reproduce the oracle."""
    so = os.path.join(ROOT, "tests", "native", "libboxgeom_host.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(so)
    from oracle import craft_post, synth
    text, link = synth.random_score_maps(5, 120, 200, n_blobs=30)
    it = craft_post.iter_components(text, link, 0.7, 0.45, 0.3)
    next(it)
    n = n_small = 0
    for k, rows, (x, y, w, h, size), (sx, ex, sy, ey), niter in it:
        ref = craft_post.component_box_restated(rows, sx, ex, sy, ey, niter)
        rmin = np.full(h, 32767, np.int16)
        rmax = np.full(h, -1, np.int16)
        for r, (a, b) in rows.items():
            rmin[r - y], rmax[r - y] = a, b
        box = np.zeros(8, np.float32)
        rc = lib.host_component_box(rmin.ctypes.data_as(ctypes.c_void_p), rmax.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_int(y), ctypes.c_int(h), ctypes.c_int(sx), ctypes.c_int(ex),
                                    ctypes.c_int(sy), ctypes.c_int(ey), ctypes.c_int(niter),
                                    box.ctypes.data_as(ctypes.c_void_p))
        assert rc == 1          # 1 = box written, 0 = empty component, -1 = hull workspace overflow
        assert np.array_equal(box.reshape(4, 2), ref), (k, box.reshape(4, 2), ref)
        # the compact workspace of the one-thread-per-box kernel gives the same box (or reports overflow)
        box2 = np.zeros(8, np.float32)
        rc2 = lib.host_component_box_small(rmin.ctypes.data_as(ctypes.c_void_p), rmax.ctypes.data_as(ctypes.c_void_p),
                                           ctypes.c_int(y), ctypes.c_int(h), ctypes.c_int(sx), ctypes.c_int(ex),
                                           ctypes.c_int(sy), ctypes.c_int(ey), ctypes.c_int(niter),
                                           box2.ctypes.data_as(ctypes.c_void_p))
        assert rc2 in (1, -1)
        if rc2 == 1:
            assert np.array_equal(box2, box), (k, box2, box)
            n_small += 1
        n += 1
    assert n > 5 and n_small > 5
