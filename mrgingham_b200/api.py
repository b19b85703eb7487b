"""ctypes binding of libmrgingham_b200.so and the Python mirror of the reference's module."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmrgingham_b200.so")
_lib = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i16p = ctypes.POINTER(ctypes.c_int16)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)
_i8p = ctypes.POINTER(ctypes.c_int8)
_ADD_POINTS_D = ctypes.CFUNCTYPE(ctypes.c_bool, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_void_p)
_ADD_POINTS = ctypes.CFUNCTYPE(ctypes.c_bool, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_double, ctypes.c_void_p)

EXPORTS = (
    "mrgingham_ChESS_response_5", "find_chessboard_corners_from_image_array_C",
    "mrg_b200_find_chessboard_corners", "mrg_b200_refine_chessboard_corners",
    "mrg_b200_detector_create", "mrg_b200_detector_destroy",
    "mrg_b200_find_corners_batch", "mrg_b200_find_corners_batch_enqueue", "mrg_b200_find_corners_batch_collect",
    "mrg_b200_find_corners_mixed_batch", "mrg_b200_debug_dump_corners",
    "mrg_b200_refine_corners_batch", "mrg_b200_find_blobs", "mrg_b200_find_blobs_batch", "mrg_b200_box_blur_batch",
    "mrg_b200_preprocess_batch", "mrg_b200_preprocess16_batch",
    "find_chessboard_from_image_array_C", "mrg_b200_find_grid_from_points", "mrg_b200_find_grid_from_points_debug", "mrg_b200_voronoi_neighbours",
    "mrg_b200_find_chessboard_from_image_array", "mrg_b200_find_circle_grid_from_image_array", "mrg_b200_find_boards_batch",
    "mrg_b200_chess_response_batch", "mrg_b200_chess_candidates_batch", "mrg_b200_pyramid_level",
    "mrg_b200_last_kernel_ms", "mrg_b200_set_profiling", "mrg_b200_last_candidate_counts", "mrg_b200_version",
    "mrg_b200_device_count",
)


class ImageDesc(ctypes.Structure):
    """mrg_b200_image_desc: one image of a mixed-size batch"""
    _fields_ = [("data", ctypes.c_void_p), ("rows", ctypes.c_int), ("cols", ctypes.c_int), ("row_pitch", ctypes.c_size_t)]


class DetectorConfig(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("max_frames", ctypes.c_int),
                ("max_rows", ctypes.c_int), ("max_cols", ctypes.c_int),
                ("candidate_capacity", ctypes.c_int), ("max_points", ctypes.c_int),
                ("kernel_variant", ctypes.c_int), ("blur_radius", ctypes.c_int), ("clahe", ctypes.c_int)]


def library_path():
    return _LIB_PATH


def lib():
    """the loaded C-ABI library; raises if it is missing (there is no other implementation)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: build it with `python -m mrgingham_b200.build` "
                          "(needs nvcc). mrgingham_b200 has no CPU fallback.")
    L = ctypes.CDLL(_LIB_PATH)
    L.mrgingham_ChESS_response_5.restype = None
    L.mrgingham_ChESS_response_5.argtypes = [_i16p, _u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.find_chessboard_corners_from_image_array_C.restype = ctypes.c_bool
    L.find_chessboard_corners_from_image_array_C.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_bool, ctypes.c_bool,
        _ADD_POINTS, ctypes.c_void_p]
    L.find_chessboard_from_image_array_C.restype = ctypes.c_bool
    L.find_chessboard_from_image_array_C.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_bool, ctypes.c_bool,
        ctypes.c_int, ctypes.c_int, _ADD_POINTS_D, ctypes.c_void_p]
    L.mrg_b200_find_grid_from_points.restype = ctypes.c_int
    L.mrg_b200_find_grid_from_points.argtypes = [_i32p, ctypes.c_int, ctypes.c_int, _f64p]
    L.mrg_b200_find_grid_from_points_debug.restype = ctypes.c_int
    L.mrg_b200_find_grid_from_points_debug.argtypes = [_i32p, ctypes.c_int, ctypes.c_int, _f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.mrg_b200_voronoi_neighbours.restype = ctypes.c_int
    L.mrg_b200_voronoi_neighbours.argtypes = [_i32p, ctypes.c_int, _i32p, _i32p, ctypes.c_int]
    L.mrg_b200_find_chessboard_from_image_array.restype = ctypes.c_int
    L.mrg_b200_find_chessboard_from_image_array.argtypes = [_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                            ctypes.c_int, _f64p, _i8p]
    L.mrg_b200_find_circle_grid_from_image_array.restype = ctypes.c_int
    L.mrg_b200_find_circle_grid_from_image_array.argtypes = [_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p]
    L.mrg_b200_find_boards_batch.restype = ctypes.c_int
    L.mrg_b200_find_boards_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             _f64p, _i8p, _i32p, ctypes.c_void_p]
    L.mrg_b200_find_chessboard_corners.restype = ctypes.c_int
    L.mrg_b200_find_chessboard_corners.argtypes = [_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, ctypes.c_int]
    L.mrg_b200_refine_chessboard_corners.restype = ctypes.c_int
    L.mrg_b200_refine_chessboard_corners.argtypes = [_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p, _i8p, ctypes.c_int]
    L.mrg_b200_detector_create.restype = ctypes.c_int
    L.mrg_b200_detector_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(DetectorConfig)]
    L.mrg_b200_detector_destroy.restype = None
    L.mrg_b200_detector_destroy.argtypes = [ctypes.c_void_p]
    batch_args = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                  ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
    L.mrg_b200_find_corners_batch.restype = ctypes.c_int
    L.mrg_b200_find_corners_batch.argtypes = batch_args + [_i32p, _i32p, ctypes.c_void_p]
    L.mrg_b200_find_corners_batch_enqueue.restype = ctypes.c_int
    L.mrg_b200_find_corners_batch_enqueue.argtypes = batch_args + [ctypes.c_void_p]
    L.mrg_b200_find_corners_mixed_batch.restype = ctypes.c_int
    L.mrg_b200_find_corners_mixed_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(ImageDesc), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    _i32p, _i32p, ctypes.c_void_p]
    L.mrg_b200_find_corners_batch_collect.restype = ctypes.c_int
    L.mrg_b200_find_corners_batch_collect.argtypes = [ctypes.c_void_p, _i32p, _i32p]
    L.mrg_b200_refine_corners_batch.restype = ctypes.c_int
    L.mrg_b200_refine_corners_batch.argtypes = batch_args + [_f64p, _i8p, ctypes.c_int, _i32p, ctypes.c_void_p]
    L.mrg_b200_find_blobs.restype = ctypes.c_int
    L.mrg_b200_find_blobs.argtypes = [_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, ctypes.c_int]
    L.mrg_b200_find_blobs_batch.restype = ctypes.c_int
    L.mrg_b200_find_blobs_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_size_t, ctypes.c_size_t, _i32p, _i32p, ctypes.c_void_p]
    L.mrg_b200_preprocess16_batch.restype = ctypes.c_int
    L.mrg_b200_preprocess16_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.mrg_b200_box_blur_batch.restype = ctypes.c_int
    L.mrg_b200_box_blur_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.mrg_b200_preprocess_batch.restype = ctypes.c_int
    L.mrg_b200_preprocess_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_void_p]
    L.mrg_b200_chess_response_batch.restype = ctypes.c_int
    L.mrg_b200_chess_response_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.mrg_b200_chess_candidates_batch.restype = ctypes.c_int
    L.mrg_b200_chess_candidates_batch.argtypes = batch_args + [ctypes.c_void_p, ctypes.c_int, _i32p, ctypes.c_void_p]
    L.mrg_b200_pyramid_level.restype = ctypes.c_int
    L.mrg_b200_pyramid_level.argtypes = [ctypes.c_void_p, _u8p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int,
                                         _u8p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.mrg_b200_last_kernel_ms.restype = ctypes.c_int
    L.mrg_b200_last_kernel_ms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]
    L.mrg_b200_set_profiling.restype = None
    L.mrg_b200_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.mrg_b200_last_candidate_counts.restype = ctypes.c_int
    L.mrg_b200_last_candidate_counts.argtypes = [ctypes.c_void_p, _i32p, ctypes.c_int]
    L.mrg_b200_version.restype = ctypes.c_char_p
    L.mrg_b200_device_count.restype = ctypes.c_int
    _lib = L
    return L


def version():
    return lib().mrg_b200_version().decode()


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _require_gpu():
    if lib().mrg_b200_device_count() <= 0:
        raise RuntimeError("mrgingham_b200: no usable CUDA device, and there is no CPU fallback")


def _check_image(image, ndim_exact=None):
    image = np.asarray(image)
    if ndim_exact is not None and image.ndim != ndim_exact:
        raise RuntimeError(f"The input image array must have exactly {ndim_exact} dims (broadcasting not supported here); got {image.ndim}")
    if image.dtype != np.uint8:
        raise RuntimeError("The input image array must contain 8-bit unsigned data")
    if image.strides[-1] != 1:
        raise RuntimeError("Image rows must live in contiguous memory")
    return image


# ---------------------------------------------------------------------------------------------
# Mirror of the reference's Python module (mrgingham_pywrap.c)
# ---------------------------------------------------------------------------------------------
def ChESS_response_5(image):
    """mrgingham.ChESS_response_5 (mrgingham_pywrap.c:40-112): int16 response of uint8 image(s)
    [..., H, W]; leading dimensions are broadcast. Elements within 7 pixels of the border, which
    the reference leaves uninitialised, are 0 here."""
    image = np.asarray(image)
    if image.ndim < 2:
        raise RuntimeError(f"The input image array must have at least 2 dims (extra ones will be broadcasted); got {image.ndim}")
    image = _check_image(image)
    _require_gpu()
    out = np.zeros(image.shape, dtype=np.int16)
    h, w = image.shape[-2:]
    L = lib()
    for idx in np.ndindex(*image.shape[:-2]):
        sl = image[idx]
        L.mrgingham_ChESS_response_5(_ptr(out[idx], _i16p), _ptr(sl, _u8p), w, h, sl.strides[0])
    return out


def find_points(image, image_pyramid_level=0, blobs=False, debug=False):
    """mrgingham.find_points (mrgingham_pywrap.c:128-225): (N,2) float64 corner coordinates
    (x,y) in full-resolution pixels; (0,2) if nothing was found."""
    if blobs and image_pyramid_level != 0:
        raise RuntimeError("blob detector requires that image_pyramid_level == 0")
    image = _check_image(image, ndim_exact=2)
    _require_gpu()
    result = {}

    def add_points(xy, n, scale, cookie):
        a = np.ctypeslib.as_array(xy, shape=(2 * n,)).astype(np.float64)
        result["xy"] = (scale * a).reshape(n, 2)
        return True

    cb = _ADD_POINTS(add_points)
    ok = lib().find_chessboard_corners_from_image_array_C(image.shape[0], image.shape[1], image.strides[0],
                                                          image.ctypes.data, int(image_pyramid_level),
                                                          bool(blobs), bool(debug), cb, None)
    if not ok:
        if "xy" in result:
            # false WITH a result is the bridge's error return (mrgingham_pywrap.c:212-219): the GPU path failed
            raise RuntimeError("find_chessboard_corners_from_image_array_C() failed")
        return np.zeros((0, 2), dtype=np.float64)      # nothing found: an empty array, no error (:203-211)
    return result["xy"]


find_chessboard_corners = find_points


def find_board(image, image_pyramid_level=-1, gridn=10, blobs=False, debug=False, debug_sequence=None):
    """mrgingham.find_board (mrgingham_pywrap.c:227-337): corners (or blobs) -> gridn x gridn grid -> refinement.
    Returns the ordered (gridn*gridn, 2) float64 pixel coordinates, or None if no board was found.
    image_pyramid_level < 0 tries levels 3,2,1,0 in turn (mrgingham.cc:127-138)."""
    if blobs and image_pyramid_level != 0:
        raise RuntimeError("blob detector requires that image_pyramid_level == 0")
    dsx = dsy = -1
    if debug_sequence is not None:
        try:
            dsx, dsy = (int(v) for v in str(debug_sequence).split(","))
        except ValueError:
            raise RuntimeError("Couldn't parse debug_sequence as an 'INTEGER,INTEGER' string")
    image = _check_image(image, ndim_exact=2)
    _require_gpu()
    result = {}

    def add_points(xy, n, cookie):
        result["xy"] = np.ctypeslib.as_array(xy, shape=(2 * n,)).astype(np.float64).reshape(n, 2)
        return True

    cb = _ADD_POINTS_D(add_points)
    ok = lib().find_chessboard_from_image_array_C(image.shape[0], image.shape[1], image.strides[0], image.ctypes.data,
                                                  int(gridn), int(image_pyramid_level), bool(blobs), bool(debug), dsx, dsy, cb, None)
    if not ok or "xy" not in result:
        return None
    return result["xy"]


find_chessboard = find_board


# ---------------------------------------------------------------------------------------------
# C++ API mirrors and batch interface
# ---------------------------------------------------------------------------------------------
def find_chessboard_corners_int(image, image_pyramid_level=0, cap=1 << 16):
    """mrgingham::find_chessboard_corners_from_image_array: (N,2) int32 PointInt list (x1000)"""
    image = _check_image(image, ndim_exact=2)
    _require_gpu()
    xy = np.empty((cap, 2), dtype=np.int32)
    n = lib().mrg_b200_find_chessboard_corners(_ptr(image, _u8p), image.shape[0], image.shape[1], image.strides[0],
                                               int(image_pyramid_level), _ptr(xy, _i32p), cap)
    if n < 0:
        raise RuntimeError("mrg_b200_find_chessboard_corners() failed")
    if n > cap:
        return find_chessboard_corners_int(image, image_pyramid_level, cap=n)
    return xy[:n].copy()


def find_blobs_int(image, cap=1 << 14):
    """mrgingham::find_blobs_from_image_array (find_blobs.cc:14-46): (N,2) int32 PointInt list (x1000)"""
    image = _check_image(image, ndim_exact=2)
    _require_gpu()
    xy = np.empty((cap, 2), dtype=np.int32)
    n = lib().mrg_b200_find_blobs(_ptr(image, _u8p), image.shape[0], image.shape[1], image.strides[0], _ptr(xy, _i32p), cap)
    if n < 0:
        raise RuntimeError("mrg_b200_find_blobs() failed")
    if n > cap:
        return find_blobs_int(image, cap=n)
    return xy[:n].copy()


def find_grid_from_points(points, gridn=10, debug=False, debug_sequence=None):
    """mrgingham::find_grid_from_points (find_grid.cc:1216-1445): points = (N,2) int PointInt list (x1000);
    returns the ordered (gridn*gridn, 2) float64 pixel coordinates, or None. Host code: works without a GPU.
    debug: the reference's /tmp/mrgingham-{2..6}-* dumps and stderr messages; debug_sequence = (x, y) in pixels: its
    stderr trace of the walks from the point nearest to that pixel."""
    pts = np.ascontiguousarray(points, dtype=np.int32).reshape(-1, 2)
    out = np.empty((gridn * gridn, 2), dtype=np.float64)
    if len(pts) == 0:
        return None
    if debug or debug_sequence is not None:
        sx, sy = (-1, -1) if debug_sequence is None else (int(debug_sequence[0]), int(debug_sequence[1]))
        rc = lib().mrg_b200_find_grid_from_points_debug(_ptr(pts, _i32p), len(pts), int(gridn), _ptr(out, _f64p), int(bool(debug)), sx, sy)
    else:
        rc = lib().mrg_b200_find_grid_from_points(_ptr(pts, _i32p), len(pts), int(gridn), _ptr(out, _f64p))
    return out if rc == 1 else None


def voronoi_neighbours(points):
    """per point, the indices of the points whose Voronoi cells share an edge with its cell, counter-clockwise
    in (x,y) from +x: the graph find_grid_from_points() walks (find_grid.cc:36-140)"""
    pts = np.ascontiguousarray(points, dtype=np.int32).reshape(-1, 2)
    n = len(pts)
    off = np.zeros(n + 1, dtype=np.int32)
    ring = np.zeros(8 * n + 16, dtype=np.int32)
    total = lib().mrg_b200_voronoi_neighbours(_ptr(pts, _i32p), n, _ptr(off, _i32p), _ptr(ring, _i32p), len(ring))
    if total > len(ring):
        ring = np.zeros(total, dtype=np.int32)
        total = lib().mrg_b200_voronoi_neighbours(_ptr(pts, _i32p), n, _ptr(off, _i32p), _ptr(ring, _i32p), len(ring))
    if total < 0:
        raise RuntimeError("mrg_b200_voronoi_neighbours() failed")
    return [ring[off[i]:off[i + 1]].tolist() for i in range(n)]


def find_chessboard_from_image_array(image, gridn=10, image_pyramid_level=-1, refine=True):
    """mrgingham::find_chessboard_from_image_array (mrgingham.cc:106-140): returns (level_found, xy, levels);
    level_found < 0 if there is no board (xy, levels are then None)"""
    image = _check_image(image, ndim_exact=2)
    _require_gpu()
    xy = np.zeros((gridn * gridn, 2), dtype=np.float64)
    lv = np.zeros(gridn * gridn, dtype=np.int8)
    r = lib().mrg_b200_find_chessboard_from_image_array(_ptr(image, _u8p), image.shape[0], image.shape[1], image.strides[0],
                                                        int(gridn), int(image_pyramid_level), int(bool(refine)), _ptr(xy, _f64p), _ptr(lv, _i8p))
    if r < -1:
        raise RuntimeError("mrg_b200_find_chessboard_from_image_array() failed on the GPU")
    if r < 0:
        return r, None, None
    return r, xy, (lv if refine else None)


VNLOG_LEGEND = "# filename x y level"


def format_vnlog(filename, xy, levels):
    """the reference CLI's output records for one image (mrgingham-from-image.cc:174-187, legend :349):
    'filename x y level' per point ("%s %f %f %d"), or 'filename - - -' when no board was found (xy is None).
    levels: one level per point, or a single int (the level the board was found at, when not refined)."""
    if xy is None:
        return "%s - - -\n" % filename
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    lv = np.broadcast_to(np.asarray(levels, dtype=np.int64), (len(xy),))
    return "".join("%s %f %f %d\n" % (filename, x, y, l) for (x, y), l in zip(xy, lv))


def refine_chessboard_corners(image, image_pyramid_level, xy, levels):
    """mrgingham::refine_chessboard_corners_from_image_array: returns (nrefined, xy', levels')"""
    image = _check_image(image, ndim_exact=2)
    xy = np.ascontiguousarray(xy, dtype=np.float64).copy()
    levels = np.ascontiguousarray(levels, dtype=np.int8).copy()
    n = lib().mrg_b200_refine_chessboard_corners(_ptr(image, _u8p), image.shape[0], image.shape[1], image.strides[0],
                                                 int(image_pyramid_level), _ptr(xy, _f64p), _ptr(levels, _i8p), len(levels))
    if n < 0:
        raise RuntimeError("mrg_b200_refine_chessboard_corners() failed")
    return n, xy, levels


class Detector:
    """Batched detector over equally-sized frames (include/mrgingham_b200.h section C)."""

    def __init__(self, max_frames=64, max_rows=0, max_cols=0, candidate_capacity=0, max_points=0, device=-1,
                 kernel_variant=0, blur_radius=0, clahe=False):
        self._h = ctypes.c_void_p()
        cfg = DetectorConfig(device, max_frames, max_rows, max_cols, candidate_capacity, max_points, kernel_variant, blur_radius,
                             int(bool(clahe)))
        if lib().mrg_b200_detector_create(ctypes.byref(self._h), ctypes.byref(cfg)) != 0:
            raise RuntimeError("mrg_b200_detector_create() failed (no CUDA device?)")
        self.max_points = max_points if max_points > 0 else 1024

    def close(self):
        if self._h:
            lib().mrg_b200_detector_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _describe(images):
        """(pointer, on_device, n, rows, cols, pitch, frame_stride, keepalive) of a numpy array or a
        torch tensor shaped [n, rows, cols] uint8 with contiguous rows"""
        if isinstance(images, np.ndarray):
            a = images if images.ndim == 3 else images[None]
            if a.shape[0] == 0:
                return 0, 0, 0, a.shape[1], a.shape[2], a.shape[2], a.shape[1] * a.shape[2], a
            assert a.dtype == np.uint8 and a.strides[2] == 1
            return a.ctypes.data, 0, a.shape[0], a.shape[1], a.shape[2], a.strides[1], a.strides[0] if a.shape[0] > 1 else a.strides[1] * a.shape[1], a
        import torch
        assert isinstance(images, torch.Tensor) and images.dtype == torch.uint8
        t = images if images.dim() == 3 else images[None]
        assert t.stride(2) == 1
        fstride = t.stride(0) if t.shape[0] > 1 else t.stride(1) * t.shape[1]
        return t.data_ptr(), 1 if t.is_cuda else 0, t.shape[0], t.shape[1], t.shape[2], t.stride(1), fstride, t

    def enqueue(self, images, level=0, stream=None):
        ptr, on_dev, n, rows, cols, pitch, fstride, keep = self._describe(images)
        self._pending = (n, keep)
        rc = lib().mrg_b200_find_corners_batch_enqueue(self._h, ptr, on_dev, n, rows, cols, pitch, fstride, int(level),
                                                      ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError("mrg_b200_find_corners_batch_enqueue() failed")

    def collect(self):
        n, _ = self._pending
        xy = np.empty((n, self.max_points, 2), dtype=np.int32)
        counts = np.empty(n, dtype=np.int32)
        if lib().mrg_b200_find_corners_batch_collect(self._h, _ptr(xy, _i32p), _ptr(counts, _i32p)) != 0:
            raise RuntimeError("mrg_b200_find_corners_batch_collect() failed")
        self._pending = None
        return xy, counts

    def find_corners(self, images, level=0, stream=None):
        """returns (xy int32 [n, max_points, 2] scaled by 1000, counts int32 [n])"""
        self.enqueue(images, level, stream)
        return self.collect()

    def find_corners_mixed(self, images, level=0, stream=None):
        """One call over a list of images of ANY sizes (each a 2-D uint8 numpy array, or each a 2-D CUDA torch tensor),
        as the reference CLI takes a glob of images (mrgingham-from-image.cc:50-54):
        returns (xy int32 [n, max_points, 2] scaled by 1000, counts int32 [n]) in the order of `images`."""
        n = len(images)
        descs = (ImageDesc * max(n, 1))()
        on_dev = None
        for i, im in enumerate(images):
            if isinstance(im, np.ndarray):
                assert im.ndim == 2 and im.dtype == np.uint8 and (im.shape[1] == 1 or im.strides[1] == 1)
                dev, ptr, pitch = 0, im.ctypes.data, im.strides[0]
            else:
                assert im.dim() == 2 and im.element_size() == 1 and im.stride(1) == 1
                dev, ptr, pitch = (1 if im.is_cuda else 0), im.data_ptr(), im.stride(0)
            assert on_dev in (None, dev), "host and device images cannot be mixed in one call"
            on_dev = dev
            descs[i] = ImageDesc(ptr, im.shape[0], im.shape[1], pitch)
        xy = np.empty((n, self.max_points, 2), dtype=np.int32)
        counts = np.zeros(n, dtype=np.int32)
        rc = lib().mrg_b200_find_corners_mixed_batch(self._h, descs, n, on_dev or 0, int(level), _ptr(xy, _i32p), _ptr(counts, _i32p),
                                                     ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError("mrg_b200_find_corners_mixed_batch() failed")
        return xy, counts

    def find_blobs(self, images, stream=None):
        """batched blob detector: returns (xy int32 [n, max_points, 2] scaled by 1000, counts int32 [n])"""
        ptr, on_dev, n, rows, cols, pitch, fstride, keep = self._describe(images)
        xy = np.empty((n, self.max_points, 2), dtype=np.int32)
        counts = np.zeros(n, dtype=np.int32)
        rc = lib().mrg_b200_find_blobs_batch(self._h, ptr, on_dev, n, rows, cols, pitch, fstride,
                                             _ptr(xy, _i32p), _ptr(counts, _i32p), ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError("mrg_b200_find_blobs_batch() failed")
        return xy, counts

    def box_blur(self, images, radius=1, out=None, stream=None):
        """cv::blur(Size(1+2R,1+2R)) as the reference CLI applies it by default (mrgingham-from-image.cc:106-111).
        Host images -> numpy result; a CUDA torch tensor -> a new CUDA tensor (or `out`), no host round trip."""
        if not 1 <= int(radius) <= 4:
            raise ValueError("radius must be in [1,4]")
        return self.preprocess(images, clahe=False, blur_radius=radius, out=out, stream=stream)

    def preprocess(self, images, clahe=False, blur_radius=1, out=None, stream=None):
        """The reference CLI's preprocessing chain (mrgingham-from-image.cc:71-111): with clahe,
        cv::normalize(0,255,NORM_MINMAX) + CLAHE(clipLimit 8); then cv::blur of the given radius (0 = none).
        Host images -> numpy result; a CUDA torch tensor -> a new CUDA tensor (or `out`), no host round trip."""
        ptr, on_dev, n, rows, cols, pitch, fstride, keep = self._describe(images)
        if on_dev:
            import torch
            if out is None:
                out = torch.empty((n, rows, cols), dtype=torch.uint8, device=keep.device)
            assert out.is_contiguous() and out.shape == (n, rows, cols)
            optr = out.data_ptr()
        else:
            out = np.empty((n, rows, cols), dtype=np.uint8)
            optr = out.ctypes.data
        rc = lib().mrg_b200_preprocess_batch(self._h, ptr, on_dev, n, rows, cols, pitch, fstride, int(bool(clahe)), int(blur_radius),
                                             optr, on_dev, ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError("mrg_b200_preprocess_batch() failed")
        return out

    def preprocess16(self, images, clahe=False, blur_radius=1, out=None, stream=None):
        """The reference CLI's chain for 16-bit images (mrgingham-from-image.cc:83-111): with clahe, normalize to
        [0,65535] + CLAHE(clipLimit 8) on the 16-bit data; convertTo 8 bits (x 255/65535); blur of the given radius.
        images: uint16 [n, rows, cols] numpy array or CUDA torch tensor (int16 storage is read as uint16).
        Host images -> numpy uint8 result; CUDA tensor -> a new CUDA uint8 tensor (or `out`)."""
        if isinstance(images, np.ndarray):
            a = images if images.ndim == 3 else images[None]
            assert a.dtype == np.uint16 and (a.shape[2] == 1 or a.strides[2] == 2)
            n, rows, cols = a.shape
            ptr, on_dev, pitch = a.ctypes.data, 0, a.strides[1]
            fstride = a.strides[0] if n > 1 else pitch * rows
            out = np.empty((n, rows, cols), dtype=np.uint8)
            optr = out.ctypes.data
        else:
            import torch
            t = images if images.dim() == 3 else images[None]
            assert t.is_cuda and t.element_size() == 2 and t.stride(2) == 1
            n, rows, cols = t.shape
            ptr, on_dev, pitch = t.data_ptr(), 1, t.stride(1) * 2
            fstride = t.stride(0) * 2 if n > 1 else pitch * rows
            if out is None:
                out = torch.empty((n, rows, cols), dtype=torch.uint8, device=t.device)
            assert out.is_contiguous() and tuple(out.shape) == (n, rows, cols)
            optr = out.data_ptr()
        rc = lib().mrg_b200_preprocess16_batch(self._h, ptr, on_dev, n, rows, cols, pitch, fstride, int(bool(clahe)), int(blur_radius),
                                               optr, on_dev, ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError("mrg_b200_preprocess16_batch() failed")
        return out

    def find_boards(self, images, gridn=10, level=-1, blobs=False, refine=True, stream=None):
        """whole boards over a batch (mrg_b200_find_boards_batch): returns (found_level int32 [n] (-1 = no board),
        xy float64 [n, gridn*gridn, 2], levels int8 [n, gridn*gridn])"""
        ptr, on_dev, n, rows, cols, pitch, fstride, keep = self._describe(images)
        xy = np.zeros((n, gridn * gridn, 2), dtype=np.float64)
        lv = np.zeros((n, gridn * gridn), dtype=np.int8)
        found = np.full(n, -1, dtype=np.int32)
        rc = lib().mrg_b200_find_boards_batch(self._h, ptr, on_dev, n, rows, cols, pitch, fstride, int(gridn), int(level),
                                              int(bool(blobs)), int(bool(refine)), _ptr(xy, _f64p), _ptr(lv, _i8p), _ptr(found, _i32p),
                                              ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError("mrg_b200_find_boards_batch() failed")
        return found, xy, lv

    def refine_corners(self, images, level, xy, levels, stream=None):
        """batched refinement: xy float64 [n, npoints, 2], levels int8 [n, npoints];
        returns (nrefined int32 [n], xy', levels')"""
        ptr, on_dev, n, rows, cols, pitch, fstride, keep = self._describe(images)
        xy = np.ascontiguousarray(xy, dtype=np.float64).copy().reshape(n, -1, 2)
        levels = np.ascontiguousarray(levels, dtype=np.int8).copy().reshape(n, -1)
        npoints = levels.shape[1]
        nref = np.zeros(n, dtype=np.int32)
        rc = lib().mrg_b200_refine_corners_batch(self._h, ptr, on_dev, n, rows, cols, pitch, fstride, int(level),
                                                 _ptr(xy, _f64p), _ptr(levels, _i8p), npoints, _ptr(nref, _i32p),
                                                 ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError("mrg_b200_refine_corners_batch() failed")
        return nref, xy, levels

    def chess_candidates(self, images, level=0, cand_cap=0, stream=None):
        """K1 alone: (counts int32 [n], candidates uint64 [n, cand_cap] or None) -- the pixels with response > 15 as
        y << 32 | x << 16 | r words, unordered (mrg_b200_chess_candidates_batch)"""
        ptr, on_dev, n, rows, cols, pitch, fstride, keep = self._describe(images)
        counts = np.zeros(n, dtype=np.int32)
        cand = np.zeros((n, cand_cap), dtype=np.uint64) if cand_cap > 0 else None
        rc = lib().mrg_b200_chess_candidates_batch(self._h, ptr, on_dev, n, rows, cols, pitch, fstride, int(level),
                                                   cand.ctypes.data if cand is not None else None, int(cand_cap), _ptr(counts, _i32p),
                                                   ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise RuntimeError("mrg_b200_chess_candidates_batch() failed")
        return counts, cand

    def chess_response(self, images):
        """dense int16 response of host images [n, rows, cols] (border elements are 0)"""
        ptr, on_dev, n, rows, cols, pitch, fstride, keep = self._describe(images)
        assert not on_dev
        out = np.zeros((n, rows, cols), dtype=np.int16)
        if lib().mrg_b200_chess_response_batch(self._h, ptr, 0, n, rows, cols, pitch, fstride, out.ctypes.data, 0, None) != 0:
            raise RuntimeError("mrg_b200_chess_response_batch() failed")
        return out

    def set_profiling(self, enabled=True):
        lib().mrg_b200_set_profiling(self._h, 1 if enabled else 0)

    def last_kernel_ms(self, which):
        ms, n = ctypes.c_float(), ctypes.c_int()
        lib().mrg_b200_last_kernel_ms(self._h, which, ctypes.byref(ms), ctypes.byref(n))
        return ms.value, n.value

    def last_candidate_counts(self, n):
        c = np.empty(n, dtype=np.int32)
        lib().mrg_b200_last_candidate_counts(self._h, _ptr(c, _i32p), n)
        return c

    def pyramid_level(self, image, level):
        image = _check_image(image, ndim_exact=2)
        oh, ow = ctypes.c_int(), ctypes.c_int()
        if lib().mrg_b200_pyramid_level(self._h, _ptr(image, _u8p), image.shape[0], image.shape[1], image.strides[0], level,
                                        None, ctypes.byref(oh), ctypes.byref(ow)) != 0:
            return None
        out = np.empty((oh.value, ow.value), dtype=np.uint8)
        lib().mrg_b200_pyramid_level(self._h, _ptr(image, _u8p), image.shape[0], image.shape[1], image.strides[0], level,
                                     _ptr(out, _u8p), ctypes.byref(oh), ctypes.byref(ow))
        return out


_default_detector = None


def pyramid_level(image, level):
    global _default_detector
    if _default_detector is None:
        _default_detector = Detector(max_frames=1)
    return _default_detector.pyramid_level(image, level)
