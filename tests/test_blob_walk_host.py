"""The mark-free, per-border formulation of cv::findContours that the CUDA blob kernels are built from
(mrgingham_b200/csrc/blob_walk.cuh), run on the CPU and compared with the oracle's Suzuki-Abe restatement
(oracle/blob_oracle.c, itself pinned to cv2.findContours): same contours, same order, same point sequences,
same area sums. No GPU needed: the header is plain C++ for g++."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "host", "libblob_walk_host.so")


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(ROOT, "tests", "host", "blob_walk_host.cpp")
    hdr = os.path.join(ROOT, "mrgingham_b200", "csrc", "blob_walk.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", SO])
    lib = ctypes.CDLL(SO)
    lib.blob_walk_host_contours.restype = ctypes.c_int
    lib.blob_walk_host_contours_segments.restype = ctypes.c_int
    return lib


def walk_contours(lib, binary, segments=False):
    binary = np.ascontiguousarray(binary, dtype=np.uint8)
    h, w = binary.shape
    xy = np.empty((4 * w * h + 16, 2), dtype=np.int32)
    lens = np.empty(w * h + 16, dtype=np.int32)
    area2 = np.empty(w * h + 16, dtype=np.int64)
    args = (binary.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), w, h, binary.strides[0],
            xy.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(xy),
            lens.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            area2.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), len(lens))
    if segments:
        nseg = ctypes.c_int(0)
        n = lib.blob_walk_host_contours_segments(*args, ctypes.byref(nseg))
    else:
        n = lib.blob_walk_host_contours(*args)
    assert n >= 0
    ends = np.cumsum(lens[:n])
    return [xy[e - l:e].copy() for e, l in zip(ends, lens[:n])], area2[:n].copy()


def random_binaries(rng, count):
    for t in range(count):
        h = int(rng.integers(1, 48)); w = int(rng.integers(1, 140))
        kind = t % 4
        if kind == 0:
            b = rng.random((h, w)) < rng.choice([0.05, 0.3, 0.5, 0.7, 0.95])
        elif kind == 1:                                   # smooth blobs: long borders, holes, nesting
            f = rng.random((h, w))
            for _ in range(int(rng.integers(1, 4))):
                p = np.pad(f, 1, mode="edge")
                f = sum(p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)) / 9
            b = f > np.quantile(f, rng.uniform(0.2, 0.8))
        elif kind == 2:                                   # thin lines and rings: pixels that lie on several borders
            b = np.zeros((h, w), dtype=bool)
            for _ in range(int(rng.integers(1, 12))):
                y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
                y1, x1 = int(rng.integers(y0, h)), int(rng.integers(x0, w))
                b[y0, x0:x1 + 1] = True; b[y1, x0:x1 + 1] = True; b[y0:y1 + 1, x0] = True; b[y0:y1 + 1, x1] = True
            b ^= rng.random((h, w)) < 0.02
        else:                                             # everything set / frame-touching shapes
            b = np.ones((h, w), dtype=bool)
            b &= ~(rng.random((h, w)) < rng.choice([0.0, 0.02, 0.2]))
        yield b.astype(np.uint8)


def test_walk_equals_suzuki_abe(lib):
    rng = np.random.default_rng(5)
    nborders = 0
    for b in random_binaries(rng, 1600):
        want = po.blob_find_contours(b)
        got, area2 = walk_contours(lib, b)
        assert len(got) == len(want), (b.shape, len(got), len(want))
        for g, wv, a in zip(got, want, area2):
            assert np.array_equal(g, wv)
            x, y = wv[:, 0].astype(np.int64), wv[:, 1].astype(np.int64)
            assert a == int(np.sum(np.roll(x, 1) * y - x * np.roll(y, 1)))
        nborders += len(want)
    assert nborders > 20000


def test_walk_word_boundaries(lib):
    """widths around multiples of 32 and shapes hugging the image edge (the bit-plane words' seams)"""
    rng = np.random.default_rng(6)
    for w in (31, 32, 33, 63, 64, 65, 96, 127, 128, 129):
        for _ in range(12):
            h = int(rng.integers(2, 20))
            b = (rng.random((h, w)) < 0.6).astype(np.uint8)
            b[:, -1] = rng.integers(0, 2); b[:, 0] = rng.integers(0, 2)
            want = po.blob_find_contours(b)
            got, _ = walk_contours(lib, b)
            assert len(got) == len(want)
            assert all(np.array_equal(g, wv) for g, wv in zip(got, want))


def test_segment_chains_equal_suzuki_abe(lib):
    """the segment formulation (candidates + row / column cuts, chained): same contours again, on images large enough
    to hold several cut rows and columns"""
    rng = np.random.default_rng(7)
    nborders = 0
    for t in range(260):
        h = int(rng.integers(1, 330)); w = int(rng.integers(1, 400))
        kind = t % 4
        if kind == 0:
            b = rng.random((h, w)) < rng.choice([0.05, 0.3, 0.5, 0.7, 0.95])
        elif kind == 1:
            f = rng.random((h, w))
            for _ in range(int(rng.integers(2, 7))):
                p = np.pad(f, 1, mode="edge")
                f = sum(p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)) / 9
            b = f > np.quantile(f, rng.uniform(0.2, 0.8))
        elif kind == 2:
            b = np.zeros((h, w), dtype=bool)
            for _ in range(int(rng.integers(1, 14))):
                y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
                y1, x1 = int(rng.integers(y0, h)), int(rng.integers(x0, w))
                b[y0, x0:x1 + 1] = True; b[y1, x0:x1 + 1] = True; b[y0:y1 + 1, x0] = True; b[y0:y1 + 1, x1] = True
            b ^= rng.random((h, w)) < 0.01
        else:
            b = np.ones((h, w), dtype=bool)
            b &= ~(rng.random((h, w)) < rng.choice([0.0, 0.02, 0.2]))
        b = b.astype(np.uint8)
        want = po.blob_find_contours(b)
        got, area2 = walk_contours(lib, b, segments=True)
        assert len(got) == len(want), (b.shape, len(got), len(want))
        for g, wv in zip(got, want):
            assert np.array_equal(g, wv)
        nborders += len(want)
    assert nborders > 20000
