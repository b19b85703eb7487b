"""TEST INFRASTRUCTURE ONLY: ctypes handles onto the CPU oracle (oracle/libmrg_oracle.so, the
restatement in mrg_oracle.c) and, where it was built, onto the unmodified reference hot path
(oracle/_ref/libmrgingham_ref.so, see oracle/Makefile) and the unmodified reference grid finder / board
pipeline (oracle/_ref/libmrgingham_ref_grid.so: find_grid.cc + mrgingham.cc over a Boost.Polygon voronoi stand-in).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module.
Nothing under mrgingham_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libmrg_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libmrgingham_ref.so")
_REF_GRID_SO = os.path.join(_HERE, "_ref", "libmrgingham_ref_grid.so")

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i16p = ctypes.POINTER(ctypes.c_int16)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)
_i8p = ctypes.POINTER(ctypes.c_int8)


def build(ref=True):
    """compile the oracle (always) and the reference build (only where /root/reference exists)"""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref", "refgrid"], check=True)
        # the reference's own Python module linked against the product library (drop-in demonstration for the tests)
        if os.path.exists(os.path.join(os.path.dirname(_HERE), "mrgingham_b200", "libmrgingham_b200.so")):
            subprocess.run(["make", "-s", "-C", _HERE, "pymodule", "hybrid"], check=True)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _check_image(image):
    image = np.asarray(image)
    assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1
    return image


class _Lib:
    def __init__(self, path, prefix):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix

    def fn(self, name):
        return getattr(self.lib, name)


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        _oracle = ctypes.CDLL(_ORACLE_SO)
        _oracle.oracle_find_corners.restype = ctypes.c_int
        _oracle.oracle_refine_corners.restype = ctypes.c_int
        _oracle.oracle_pyramid.restype = ctypes.c_int
        _oracle.blob_oracle_find_contours.restype = ctypes.c_int
        _oracle.blob_oracle_centers.restype = ctypes.c_int
        _oracle.blob_oracle_find_blobs.restype = ctypes.c_int
    return _oracle


def have_ref():
    return os.path.exists(_REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(_REF_SO)
        _ref.ref_find_chessboard_corners.restype = ctypes.c_int
        _ref.ref_refine_chessboard_corners.restype = ctypes.c_int
    return _ref


# ---------------------------------------------------------------------------------------------
# oracle (restatement)
# ---------------------------------------------------------------------------------------------
def chess_response_5(image, fill=0):
    """dense int16 response; pixels the reference never writes hold `fill`"""
    image = _check_image(image)
    h, w = image.shape
    out = np.full((h, w), fill, dtype=np.int16)
    oracle_lib().oracle_chess_response_5(_ptr(out, _i16p), _ptr(image, _u8p), w, h, image.strides[0])
    return out


def pyramid(image, level):
    image = _check_image(image)
    h, w = image.shape
    oh, ow = ctypes.c_int(), ctypes.c_int()
    rc = oracle_lib().oracle_pyramid(_ptr(image, _u8p), h, w, image.strides[0], level, None,
                                     ctypes.byref(oh), ctypes.byref(ow))
    if rc != 0:
        return None
    out = np.empty((oh.value, ow.value), dtype=np.uint8)
    oracle_lib().oracle_pyramid(_ptr(image, _u8p), h, w, image.strides[0], level, _ptr(out, _u8p),
                                ctypes.byref(oh), ctypes.byref(ow))
    return out


def find_corners(image, level=0, cap=1 << 20, want_double=False):
    """(N,2) int32 PointInt list (x,y scaled by 1000), in the reference's output order"""
    image = _check_image(image)
    h, w = image.shape
    xy = np.empty((cap, 2), dtype=np.int32)
    xyd = np.empty((cap, 2), dtype=np.float64) if want_double else None
    n = oracle_lib().oracle_find_corners(_ptr(image, _u8p), h, w, image.strides[0], level,
                                         _ptr(xy, _i32p), _ptr(xyd, _f64p) if want_double else None, cap)
    assert n <= cap
    return (xy[:n].copy(), xyd[:n].copy()) if want_double else xy[:n].copy()


def refine_corners(image, level, xy, levels):
    """returns (nrefined, xy', levels')"""
    image = _check_image(image)
    h, w = image.shape
    xy = np.ascontiguousarray(xy, dtype=np.float64).copy()
    levels = np.ascontiguousarray(levels, dtype=np.int8).copy()
    n = oracle_lib().oracle_refine_corners(_ptr(image, _u8p), h, w, image.strides[0], level,
                                           _ptr(xy, _f64p), _ptr(levels, _i8p), len(levels))
    return n, xy, levels


# ---------------------------------------------------------------------------------------------
# blob path (oracle/blob_oracle.c)
# ---------------------------------------------------------------------------------------------
def blob_find_contours(binary):
    """cv2.findContours(binary, RETR_LIST, CHAIN_APPROX_NONE): list of (n,2) int32 arrays (x,y)"""
    binary = _check_image(binary)
    h, w = binary.shape
    xy = np.empty((4 * w * h + 16, 2), dtype=np.int32)
    lens = np.empty(w * h + 16, dtype=np.int32)
    n = oracle_lib().blob_oracle_find_contours(_ptr(binary, _u8p), w, h, binary.strides[0], _ptr(xy, _i32p), len(xy),
                                               _ptr(lens, _i32p), len(lens))
    assert n >= 0
    ends = np.cumsum(lens[:n])
    return [xy[e - l:e].copy() for e, l in zip(ends, lens[:n])]


def blob_centers(image, thresh, cap=1 << 16):
    """SimpleBlobDetector::findBlobs for one threshold: (n,4) float64 rows x, y, radius, confidence"""
    image = _check_image(image)
    h, w = image.shape
    out = np.empty((cap, 4), dtype=np.float64)
    n = oracle_lib().blob_oracle_centers(_ptr(image, _u8p), w, h, image.strides[0], int(thresh), _ptr(out, _f64p), cap)
    assert n >= 0
    return out[:n].copy()


def find_blobs(image, cap=1 << 16):
    """find_blobs_from_image_array: (N,2) int32 PointInt list (x,y scaled by 1000), reference order"""
    image = _check_image(image)
    h, w = image.shape
    xy = np.empty((cap, 2), dtype=np.int32)
    n = oracle_lib().blob_oracle_find_blobs(_ptr(image, _u8p), w, h, image.strides[0], _ptr(xy, _i32p), cap)
    assert n >= 0
    return xy[:n].copy()


def box_blur(image, radius=1):
    """cv::blur(image, Size(1+2R,1+2R)) as the reference CLI applies it (mrgingham-from-image.cc:106-111)"""
    image = _check_image(image)
    h, w = image.shape
    out = np.empty((h, w), dtype=np.uint8)
    oracle_lib().blob_oracle_box_blur(_ptr(image, _u8p), w, h, image.strides[0], int(radius), _ptr(out, _u8p))
    return out


def normalize_minmax(image):
    """cv::normalize(image, image, 0, 255, NORM_MINMAX) for 8-bit images (mrgingham-from-image.cc:77)"""
    image = _check_image(image)
    h, w = image.shape
    lut = np.empty(256, dtype=np.uint8)
    oracle_lib().preproc_oracle_normalize_lut(_ptr(image, _u8p), w, h, image.strides[0], _ptr(lut, _u8p))
    return lut[image]


def clahe(image, clip_limit=8.0):
    """cv::createCLAHE(clip_limit).apply(image), 8x8 tiles (mrgingham-from-image.cc:43-44, :78)"""
    image = _check_image(image)
    h, w = image.shape
    out = np.empty((h, w), dtype=np.uint8)
    oracle_lib().preproc_oracle_clahe(_ptr(image, _u8p), w, h, image.strides[0], ctypes.c_double(clip_limit), _ptr(out, _u8p))
    return out


def normalize_clahe(image):
    """the CLI's --clahe chain: normalize, then CLAHE with clip limit 8"""
    image = _check_image(image)
    h, w = image.shape
    out = np.empty((h, w), dtype=np.uint8)
    oracle_lib().preproc_oracle_normalize_clahe(_ptr(image, _u8p), w, h, image.strides[0], _ptr(out, _u8p))
    return out


_u16p = ctypes.POINTER(ctypes.c_uint16)


def _check_image16(image):
    image = np.asarray(image)
    assert image.ndim == 2 and image.dtype == np.uint16 and (image.shape[1] == 1 or image.strides[1] == 2) and image.strides[0] % 2 == 0
    return image


def convert16to8(image):
    """cv::Mat::convertTo(CV_8U, 255./65535.) of a 16-bit image (mrgingham-from-image.cc:93)"""
    image = _check_image16(image)
    h, w = image.shape
    out = np.empty((h, w), dtype=np.uint8)
    oracle_lib().preproc_oracle_convert16to8(_ptr(image, _u16p), w, h, image.strides[0] // 2, _ptr(out, _u8p))
    return out


def normalize16(image):
    """cv::normalize(image, image, 0, 65535, NORM_MINMAX) for 16-bit images (mrgingham-from-image.cc:90)"""
    image = _check_image16(image)
    h, w = image.shape
    out = np.empty((h, w), dtype=np.uint16)
    oracle_lib().preproc_oracle_normalize16(_ptr(image, _u16p), w, h, image.strides[0] // 2, _ptr(out, _u16p))
    return out


def clahe16(image, clip_limit=8.0):
    """cv::createCLAHE(clip_limit).apply(image) for 16-bit images, 8x8 tiles (mrgingham-from-image.cc:91)"""
    image = _check_image16(image)
    h, w = image.shape
    out = np.empty((h, w), dtype=np.uint16)
    oracle_lib().preproc_oracle_clahe16(_ptr(image, _u16p), w, h, image.strides[0] // 2, ctypes.c_double(clip_limit), _ptr(out, _u16p))
    return out


def chain16(image, clahe=False):
    """the CLI's handling of a 16-bit image (mrgingham-from-image.cc:83-93): [normalize + CLAHE(8)], then 8 bits"""
    image = _check_image16(image)
    h, w = image.shape
    out = np.empty((h, w), dtype=np.uint8)
    oracle_lib().preproc_oracle_chain16(_ptr(image, _u16p), w, h, image.strides[0] // 2, int(bool(clahe)), _ptr(out, _u8p))
    return out


# ---------------------------------------------------------------------------------------------
# the reference itself (oracle/_ref)
# ---------------------------------------------------------------------------------------------
def ref_chess_response_5(image, fill=0):
    image = _check_image(image)
    h, w = image.shape
    out = np.full((h, w), fill, dtype=np.int16)
    ref_lib().ref_ChESS_response_5(_ptr(out, _i16p), _ptr(image, _u8p), w, h, image.strides[0])
    return out


def ref_find_corners(image, level=0, cap=1 << 20):
    image = _check_image(image)
    h, w = image.shape
    xy = np.empty((cap, 2), dtype=np.int32)
    n = ref_lib().ref_find_chessboard_corners(_ptr(image, _u8p), h, w, image.strides[0], level,
                                              _ptr(xy, _i32p), cap)
    assert n <= cap
    return xy[:n].copy()


def ref_refine_corners(image, level, xy, levels):
    image = _check_image(image)
    h, w = image.shape
    xy = np.ascontiguousarray(xy, dtype=np.float64).copy()
    levels = np.ascontiguousarray(levels, dtype=np.int8).copy()
    n = ref_lib().ref_refine_chessboard_corners(_ptr(image, _u8p), h, w, image.strides[0], level,
                                                _ptr(xy, _f64p), _ptr(levels, _i8p), len(levels))
    return n, xy, levels


def ref_shim_resize(image, level):
    image = _check_image(image)
    h, w = image.shape
    oh, ow = ctypes.c_int(), ctypes.c_int()
    ref_lib().ref_shim_resize(_ptr(image, _u8p), h, w, image.strides[0], level, None,
                              ctypes.byref(oh), ctypes.byref(ow))
    out = np.empty((oh.value, ow.value), dtype=np.uint8)
    ref_lib().ref_shim_resize(_ptr(image, _u8p), h, w, image.strides[0], level, _ptr(out, _u8p),
                              ctypes.byref(oh), ctypes.byref(ow))
    return out


# ---------------------------------------------------------------------------------------------
# the reference's grid finder and board pipeline (oracle/_ref/libmrgingham_ref_grid.so): find_grid.cc and
# mrgingham.cc, unmodified, over the Voronoi stand-in oracle/shim/boost/polygon/voronoi.hpp
# ---------------------------------------------------------------------------------------------
_ref_grid = None


def have_ref_grid():
    return os.path.exists(_REF_GRID_SO)


def ref_grid_lib():
    global _ref_grid
    if _ref_grid is None:
        _ref_grid = ctypes.CDLL(_REF_GRID_SO)
        _ref_grid.ref_find_grid_from_points.restype = ctypes.c_int
        _ref_grid.ref_find_chessboard_from_image_array.restype = ctypes.c_int
        _ref_grid.ref_shim_voronoi_rings.restype = ctypes.c_int
    return _ref_grid


def ref_find_grid_from_points_debug(points, gridn, debug=True, debug_sequence=None):
    """the compiled reference's find_grid_from_points() with its diagnostics on (files under /tmp, text on stderr)"""
    pts = np.ascontiguousarray(points, dtype=np.int32).reshape(-1, 2)
    out = np.empty((gridn * gridn, 2), dtype=np.float64)
    L = ref_grid_lib()
    L.ref_find_grid_from_points_debug.restype = ctypes.c_int
    sx, sy = (-1, -1) if debug_sequence is None else (int(debug_sequence[0]), int(debug_sequence[1]))
    r = L.ref_find_grid_from_points_debug(_ptr(pts, _i32p), len(pts), gridn, _ptr(out, _f64p), int(bool(debug)), sx, sy)
    return out if r == 1 else None


def ref_find_grid_from_points(points, gridn):
    """mrgingham::find_grid_from_points (find_grid.cc:1216). points: int32 [n,2] scaled by 1000.
    Returns float64 [gridn*gridn, 2] or None."""
    pts = np.ascontiguousarray(np.asarray(points).reshape(-1, 2), dtype=np.int32)
    out = np.empty((gridn * gridn, 2), dtype=np.float64)
    r = ref_grid_lib().ref_find_grid_from_points(_ptr(pts, _i32p), len(pts), gridn, _ptr(out, _f64p))
    assert r in (0, 1)
    return out if r == 1 else None


def ref_find_chessboard(image, gridn, level=-1, refine=True):
    """mrgingham::find_chessboard_from_image_array (mrgingham.cc:106). Returns (level found at or -1, points or
    None, per-point refinement levels or None)."""
    image = _check_image(image)
    h, w = image.shape
    out = np.empty((gridn * gridn, 2), dtype=np.float64)
    lv = np.empty(gridn * gridn, dtype=np.int8)
    r = ref_grid_lib().ref_find_chessboard_from_image_array(_ptr(image, _u8p), h, w, image.strides[0], gridn, level,
                                                            1 if refine else 0, _ptr(out, _f64p), _ptr(lv, _i8p))
    assert r >= -1
    return (r, out, lv) if r >= 0 else (-1, None, None)


def ref_shim_voronoi_rings(points):
    """the stand-in's diagram: list of (source index, [neighbour source indices in next() order]) per cell, in
    cell creation order"""
    pts = np.ascontiguousarray(np.asarray(points).reshape(-1, 2), dtype=np.int32)
    n = ref_grid_lib().ref_shim_voronoi_rings(_ptr(pts, _i32p), len(pts), None, 0)
    buf = np.empty(n, dtype=np.int32)
    ref_grid_lib().ref_shim_voronoi_rings(_ptr(pts, _i32p), len(pts), _ptr(buf, _i32p), n)
    cells, at = [], 1
    for _ in range(int(buf[0])):
        src, k = int(buf[at]), int(buf[at + 1])
        cells.append((src, [int(v) for v in buf[at + 2:at + 2 + k]]))
        at += 2 + k
    return cells


# ---------------------------------------------------------------------------------------------
# the hybrid build (oracle/_ref/libmrgingham_hybrid.so): the reference's mrgingham.cc + find_grid.cc compiled
# against THIS repo's adapter headers and linked with the product library -- the reference's orchestration
# running on the CUDA detector. Same extern "C" handles as the reference grid build.
# ---------------------------------------------------------------------------------------------
_HYBRID_SO = os.path.join(_HERE, "_ref", "libmrgingham_hybrid.so")
_hybrid = None


def have_hybrid():
    return os.path.exists(_HYBRID_SO)


def hybrid_find_chessboard(image, gridn, level=-1, refine=True):
    """mrgingham::find_chessboard_from_image_array of the reference's mrgingham.cc, with the corner detector and the
    refinement coming from libmrgingham_b200.so through the cv::Mat adapter overloads. Needs a GPU."""
    global _hybrid
    if _hybrid is None:
        _hybrid = ctypes.CDLL(_HYBRID_SO)
        _hybrid.ref_find_chessboard_from_image_array.restype = ctypes.c_int
    image = _check_image(image)
    h, w = image.shape
    out = np.empty((gridn * gridn, 2), dtype=np.float64)
    lv = np.empty(gridn * gridn, dtype=np.int8)
    r = _hybrid.ref_find_chessboard_from_image_array(_ptr(image, _u8p), h, w, image.strides[0], gridn, level,
                                                     1 if refine else 0, _ptr(out, _f64p), _ptr(lv, _i8p))
    assert r >= -1
    return (r, out, lv) if r >= 0 else (-1, None, None)
