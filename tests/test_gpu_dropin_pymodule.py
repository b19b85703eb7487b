"""Drop-in proof at the reference's own Python boundary: /root/reference/mrgingham_pywrap.c, UNMODIFIED, compiled
where it lies by oracle/Makefile (target `pymodule`) and linked against the product library alone --
libmrgingham_b200.so supplies the three symbols the module binds (mrgingham_ChESS_response_5, ChESS.h:31-34;
find_chessboard_corners_from_image_array_C and find_chessboard_from_image_array_C,
mrgingham_pywrap_cplusplus_bridge.h:10-42). `import mrgingham` below is therefore the reference's module running
on the CUDA path; its results are compared with the reference's C++ code on the CPU (oracle/_ref) and, for blobs,
with the cv2-pinned oracle.

Skipped where the module was not built (it needs /root/reference at build time; the built .so travels to the GPU
box with the snapshot)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

from mrgingham_b200 import api, synth
from oracle import pyoracle as po

_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "pymodule")
_SO = sorted(glob.glob(os.path.join(_DIR, "mrgingham*.so")))

pytestmark = pytest.mark.skipif(not _SO, reason="oracle/_ref/pymodule not built (needs /root/reference)")


def _module():
    spec = importlib.util.spec_from_file_location("mrgingham", _SO[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_module_links_against_the_product_library_alone():
    m = _module()
    for name in ("ChESS_response_5", "find_points", "find_board", "find_chessboard_corners", "find_chessboard"):
        assert callable(getattr(m, name)), name
    # the module's argument checks are the reference's own (mrgingham_pywrap.c:163-178)
    with pytest.raises(Exception):
        m.find_points(np.zeros((8, 8), np.float32))


@pytest.mark.gpu
def test_reference_module_on_the_cuda_path_equals_reference_cpu():
    api._require_gpu()
    assert po.have_ref() and po.have_ref_grid()
    m = _module()
    for (w, h, gridn, seed) in ((640, 480, 10, 0), (1280, 960, 10, 1), (1920, 1080, 14, 2)):
        img = synth.board_frame(w, h, gridn, seed=seed)
        # ChESS_response_5: the border is left unwritten by both (mrgingham_pywrap.c:70-71)
        got = m.ChESS_response_5(img)
        want = po.ref_chess_response_5(img)
        assert got.dtype == np.int16 and got.shape == img.shape
        assert np.array_equal(got[7:-7, 7:-7], want[7:-7, 7:-7])
        # find_points at the levels the reference accepts
        for level in (0, 1, 2):
            pts = m.find_points(img, image_pyramid_level=level)
            ref = po.ref_find_corners(img, level).astype(np.float64) * (1.0 / 1000)
            assert pts.dtype == np.float64 and pts.shape == ref.shape and np.array_equal(pts, ref), (w, level)
        assert np.array_equal(m.find_chessboard_corners(img), m.find_points(img))
        # find_board: corners -> grid -> refinement, auto level and fixed levels
        for level in (-1, 0, 1):
            L, xy, _ = po.ref_find_chessboard(img, gridn, level, True)
            board = m.find_board(img, gridn=gridn, image_pyramid_level=level)
            assert (board is None) == (L < 0), (w, level)
            if L >= 0:
                assert board.shape == (gridn * gridn, 2) and np.array_equal(board, xy), (w, level)
    # batched ChESS_response_5 ([..., H, W] broadcasting, mrgingham_pywrap.c:84-103)
    stack = np.stack([synth.board_frame(320, 240, 6, seed=s) for s in range(3)]).reshape(3, 1, 240, 320)
    got = m.ChESS_response_5(stack)
    assert got.shape == stack.shape
    for i in range(3):
        assert np.array_equal(got[i, 0, 7:-7, 7:-7], po.ref_chess_response_5(stack[i, 0])[7:-7, 7:-7])
    # nothing to find: an empty (0,2) array / None (mrgingham_pywrap.c:190-197)
    flat = np.full((200, 300), 128, np.uint8)
    assert m.find_points(flat).shape == (0, 2)
    assert m.find_board(flat) is None
    # blobs (level 0 only, mrgingham_pywrap.c:154-158)
    dots = synth.circle_grid_frame(800, 608, 10, seed=3)
    pts = m.find_points(dots, blobs=True)
    ref = po.find_blobs(dots).astype(np.float64) * (1.0 / 1000)
    assert pts.shape == ref.shape and np.array_equal(pts, ref)
    with pytest.raises(Exception):
        m.find_points(dots, image_pyramid_level=1, blobs=True)


@pytest.mark.gpu
@pytest.mark.skipif(not po.have_hybrid(), reason="oracle/_ref/libmrgingham_hybrid.so not built (needs /root/reference)")
def test_reference_orchestrator_on_the_cuda_detector_equals_reference_cpu():
    """Source-level drop-in, executed: the reference's mrgingham.cc and find_grid.cc compiled against this repo's
    find_chessboard_corners.hh / find_blobs.hh (cv::Mat overloads) and linked with the product library
    (`make -C oracle hybrid`) against the all-reference build, image by image."""
    api._require_gpu()
    for (w, h, gridn, seed) in ((1280, 960, 10, 1), (800, 608, 10, 3), (1920, 1080, 14, 2)):
        img = synth.board_frame(w, h, gridn, seed=seed)
        for level, refine in ((-1, True), (1, True), (0, False)):
            L, xy, lv = po.ref_find_chessboard(img, gridn, level, refine)
            Lh, xyh, lvh = po.hybrid_find_chessboard(img, gridn, level, refine)
            assert Lh == L, (w, level)
            if L >= 0:
                assert np.array_equal(xyh, xy) and (not refine or np.array_equal(lvh, lv)), (w, level)
    assert po.hybrid_find_chessboard(synth.noise_frame(320, 240, seed=1), 10)[0] == -1
