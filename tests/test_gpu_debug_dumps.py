"""the reference's --debug artefacts of the corner detector (find_chessboard_corners.cc:294-315,
346-348, 391-392, 400-407, 452-459, 513-541). `debug=True` on the drop-in entry points writes the level image, the
two normalised ChESS response images and the self-plotting corner list to /tmp, with the values the reference writes:
checked here against the oracle's response / corners and, where cv2 is present, cv2.normalize."""
import os
import struct
import zlib

import numpy as np
import pytest

from mrgingham_b200 import api, synth
from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def read_png_gray8(path):
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    at, idat, w, h = 8, b"", 0, 0
    while at < len(b):
        n, typ = struct.unpack(">I4s", b[at:at + 8])
        data = b[at + 8:at + 8 + n]
        assert struct.unpack(">I", b[at + 8 + n:at + 12 + n])[0] == (zlib.crc32(typ + data) & 0xFFFFFFFF)
        if typ == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", data[:10])
            assert (depth, ctype) == (8, 0)
        elif typ == b"IDAT":
            idat += data
        at += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, w + 1)
    assert np.all(raw[:, 0] == 0)
    return raw[:, 1:].copy()


def normalized(resp):
    """cv::normalize(resp /* int16 */, out, 0, 255, NORM_MINMAX), then imwrite's saturating conversion to 8 bits"""
    try:
        import cv2
        return np.clip(cv2.normalize(resp, None, 0, 255, cv2.NORM_MINMAX), 0, 255).astype(np.uint8)
    except ImportError:
        mn, mx = float(resp.min()), float(resp.max())
        scale = 255.0 * (1.0 / (mx - mn) if mx - mn > 2.2e-16 else 0.0)
        a, b = np.float32(scale), np.float32(0.0 - mn * scale)
        v = (resp.astype(np.float64) * np.float64(a) + np.float64(b)).astype(np.float32)     # one rounding, as an FMA
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("level", [0, 1])
def test_find_debug_artefacts(level):
    api._require_gpu()
    img = synth.board_frame(640, 480, 10, seed=31)
    files = ["/tmp/mrgingham-scaled-processed-level%d.png" % level, "/tmp/mrgingham-chess-response-level%d.png" % level,
             "/tmp/mrgingham-chess-response-level%d-positive.png" % level, "/tmp/mrgingham-1-corners.vnl"]
    for f in files:
        if os.path.exists(f):
            os.remove(f)
    pts = api.find_points(img, image_pyramid_level=level, debug=True)
    assert len(pts) == len(po.find_corners(img, level))
    lvl = po.pyramid(img, level) if level else img
    assert np.array_equal(read_png_gray8(files[0]), lvl)
    resp = po.chess_response_5(lvl, fill=0)
    assert np.array_equal(read_png_gray8(files[1]), normalized(resp))
    assert np.array_equal(read_png_gray8(files[2]), normalized(np.maximum(resp, 0)))
    lines = open(files[3]).read().splitlines()
    assert lines[0] == "#!/usr/bin/feedgnuplot --dom --square --set 'yr [:] rev'" and lines[1] == "# x y"
    _, want = po.find_corners(img, level, want_double=True)
    assert lines[2:] == ["%f %f" % (x, y) for x, y in want]
    assert os.stat(files[3]).st_mode & 0o111 == 0o111


def test_no_debug_no_files():
    api._require_gpu()
    img = synth.board_frame(640, 480, 10, seed=32)
    f = "/tmp/mrgingham-1-corners.vnl"
    if os.path.exists(f):
        os.remove(f)
    api.find_points(img, image_pyramid_level=0, debug=False)
    assert not os.path.exists(f)
