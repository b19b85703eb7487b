"""mrgingham_b200: B200-native (sm_100a) replacement for the per-image chessboard-corner hot path
of dkogan/mrgingham. The Python surface mirrors the reference's `mrgingham` module for that path
(mrgingham_pywrap.c:357-368): ChESS_response_5, find_points (alias find_chessboard_corners).

All compute happens in hand-written CUDA behind the C ABI of include/mrgingham_b200.h, loaded
from mrgingham_b200/libmrgingham_b200.so. There is no CPU fallback: without the library (or
without a CUDA device at call time) every function raises.
"""
from .api import (ChESS_response_5, Detector, find_board, find_chessboard, find_chessboard_corners,  # noqa: F401
                  find_points, lib, library_path, pyramid_level, refine_chessboard_corners, version)
