"""The oracle restatement against the reference itself (oracle/_ref, built from /root/reference by
oracle/Makefile) on a wider, randomised image set than the committed fixtures. Skipped where the
reference build is absent. CPU only."""
import os
import sys

import numpy as np
import pytest

from mrgingham_b200 import synth
from oracle import pyoracle as po

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

pytestmark = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def _images():
    yield "board_720p", synth.board_frame(1280, 720, 10, seed=100)
    yield "board_n14_1080p", synth.board_frame(1920, 1080, 14, seed=101)
    yield "board_odd", synth.board_frame(611, 457, 10, seed=102)
    yield "noise", synth.noise_frame(333, 222, seed=103)
    yield "blurred_noise", synth.blurred_noise_frame(320, 256, seed=104, passes=1)
    yield "checker6", synth.checker_frame(300, 260, period=6, seed=105)
    yield "checker11", synth.checker_frame(480, 270, period=11, seed=106, noise_sigma=3.0)
    yield "blobs", synth.blob_frame(400, 300, seed=107)
    rng = np.random.default_rng(108)
    for i in range(6):
        w, h = int(rng.integers(16, 120)), int(rng.integers(16, 120))
        yield f"rand_{w}x{h}", synth.blurred_noise_frame(w, h, seed=200 + i, passes=1)


@pytest.mark.parametrize("name,img", list(_images()), ids=lambda v: v if isinstance(v, str) else "")
def test_oracle_equals_reference(name, img):
    assert np.array_equal(po.chess_response_5(img, fill=-7), po.ref_chess_response_5(img, fill=-7))
    for level in (0, 1, 2, 3):
        a, b = po.find_corners(img, level), po.ref_find_corners(img, level)
        assert a.shape == b.shape and np.array_equal(a, b), (name, level)
    for start in (1, 2, 3):
        xa, la, ca = cases.refine_chain(po.find_corners, po.refine_corners, img, start)
        xb, lb, cb = cases.refine_chain(po.ref_find_corners, po.ref_refine_corners, img, start)
        assert np.array_equal(ca, cb) and np.array_equal(la, lb), (name, start)
        assert np.array_equal(xa.view(np.uint64), xb.view(np.uint64)), (name, start)


def test_strided_input():
    big = synth.board_frame(700, 500, 10, seed=110)
    view = big[10:490, 20:660]                       # stride 700, width 640
    assert view.strides[0] == 700
    assert np.array_equal(po.chess_response_5(view, fill=0), po.ref_chess_response_5(view, fill=0))
    # level 0 on a non-continuous image is an error in the reference: no points
    assert len(po.ref_find_corners(view, 0)) == 0 and len(po.find_corners(view, 0)) == 0
    for level in (1, 2):
        a, b = po.find_corners(view, level), po.ref_find_corners(view, level)
        assert len(b) > 0 and np.array_equal(a, b)


def test_bad_levels_give_nothing():
    img = synth.board_frame(320, 240, 10, seed=111)
    for level in (-1, 11, 50):
        assert len(po.ref_find_corners(img, level)) == 0
        assert len(po.find_corners(img, level)) == 0


def test_refine_with_shuffled_and_foreign_points():
    # points in arbitrary order, some far from any corner, some outside the image
    img = synth.board_frame(800, 600, 10, seed=112)
    rng = np.random.default_rng(113)
    pts = po.ref_find_corners(img, 2).astype(np.float64) / 1000.0
    rng.shuffle(pts)
    extra = np.stack([rng.uniform(-20, 820, 40), rng.uniform(-20, 620, 40)], axis=1)
    xy = np.concatenate([pts, extra, pts[:10] + 0.7])
    levels = np.full(len(xy), 2, dtype=np.int8)
    levels[::7] = 1
    for level in (1, 0):
        na, xa, la = po.refine_corners(img, level, xy, levels)
        nb, xb, lb = po.ref_refine_corners(img, level, xy, levels)
        assert na == nb and np.array_equal(la, lb)
        assert np.array_equal(xa.view(np.uint64), xb.view(np.uint64))
        xy, levels = xb, lb
