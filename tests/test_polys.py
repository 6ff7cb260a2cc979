"""Polygon refinement: the product's host routine against a golden produced by the
reference's own getPoly_core (marie/models/craft/craft_utils.py:101-254, executed from its file by
tools/make_golden_polys.py) and, in the container, against that function directly on fresh inputs.  Both sides use the
same OpenCV primitives, so the comparison is exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "polys.npz")


def _same(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert (g is None) == (w is None)
        if g is not None:
            assert g.shape == (14, 2) and g.dtype == np.float64
            assert np.array_equal(g, w)


def test_polys_match_the_reference_golden():
    from marie_icr_b200.polys import get_poly_core
    z = np.load(GOLD)
    n_poly = 0
    for i in range(4):
        boxes, labels, mapper = z[f"boxes{i}"], z[f"labels{i}"], z[f"mapper{i}"]
        has, polys = z[f"has{i}"], z[f"polys{i}"]
        want, k = [], 0
        for flag in has:
            want.append(polys[k] if flag else None)
            k += int(flag)
        _same(get_poly_core(list(boxes), labels, mapper), want)
        n_poly += int(has.sum())
    assert n_poly >= 30                       # the golden exercises the polygon branch, not only the early exits


def test_polys_match_the_reference_function_live():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    from make_golden_polys import curved_maps
    from marie_icr_b200.polys import get_poly_core
    cu = ref_loader.load()["craft_utils"]
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        for seed in (21, 22, 23):
            text, link = curved_maps(seed, h=256, w=512, n_words=8)
            boxes, labels, mapper = cu.getDetBoxes_core(text, link, 0.7, 0.4, 0.4)
            _same(get_poly_core(boxes, labels, mapper, link), cu.getPoly_core(boxes, labels, mapper, link))
    finally:
        os.chdir(cwd)


def test_adjust_polys_scales_and_substitutes():
    from marie_icr_b200.polys import adjust_polys
    boxes = [np.ones((4, 2), np.float32) * 3, np.ones((4, 2), np.float32) * 5]
    poly = np.arange(28, dtype=np.float64).reshape(14, 2)
    out = adjust_polys([None, poly], boxes, 1.25, 1.5)
    assert out[0] is boxes[0]
    assert np.array_equal(out[1], poly * (2.5, 3.0)) and np.array_equal(poly, np.arange(28).reshape(14, 2))
