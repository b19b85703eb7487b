#!/usr/bin/env python
"""Regenerates tests/golden/golden_v1.npz FROM THE REFERENCE ITSELF (oracle/_ref, i.e. the
unmodified /root/reference/ChESS.c + find_chessboard_corners.cc compiled by oracle/Makefile).
Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import pyoracle as po  # noqa: E402
import cases  # noqa: E402


def main():
    po.build(ref=True)
    assert po.have_ref(), "oracle/_ref was not built (is /root/reference mounted?)"
    out = {}
    for name, img in cases.golden_images().items():
        out[f"{name}/image"] = img
        out[f"{name}/response"] = po.ref_chess_response_5(img, fill=0)
        for level in cases.LEVELS:
            out[f"{name}/corners_L{level}"] = po.ref_find_corners(img, level)
        for start in (1, 2, 3):
            xy, levels, counts = cases.refine_chain(po.ref_find_corners, po.ref_refine_corners, img, start)
            out[f"{name}/refine_from_L{start}/xy"] = xy
            out[f"{name}/refine_from_L{start}/levels"] = levels
            out[f"{name}/refine_from_L{start}/counts"] = counts
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
    for name in cases.golden_images():
        print(f"  {name:22s}", " ".join(f"L{l}:{len(out[f'{name}/corners_L{l}'])}" for l in cases.LEVELS),
              " refine3:", out[f"{name}/refine_from_L3/counts"].tolist(),
              " levels:", np.bincount(out[f"{name}/refine_from_L3/levels"].astype(np.int64) , minlength=4).tolist())


if __name__ == "__main__":
    main()
