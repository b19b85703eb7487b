#!/usr/bin/env python
"""Single-image latency of the drop-in entry points (host image in, host points out), next to the
reference's own code on one host core: the per-image use the reference API is written for."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mrgingham_b200 import api, synth
from oracle import pyoracle as po

ref = po.ref_find_corners if po.have_ref() else po.find_corners
kind = "reference" if po.have_ref() else "port"
for (w, h) in ((640, 480), (1920, 1080), (3840, 2160)):
    img = synth.board_frame(w, h, 10, seed=1)
    dots = synth.circle_grid_frame(w, h, 10, seed=2)
    for level in (0, 2):
        got = api.find_chessboard_corners_int(img, level)          # warm-up (allocations, module load)
        assert np.array_equal(got, po.find_corners(img, level))
        t0 = time.perf_counter()
        for _ in range(20):
            api.find_chessboard_corners_int(img, level)
        gpu = (time.perf_counter() - t0) / 20
        t0 = time.perf_counter()
        for _ in range(3):
            ref(img, level)
        cpu = (time.perf_counter() - t0) / 3
        print(json.dumps({"call": "find_chessboard_corners_from_image_array", "image": f"{w}x{h}", "level": level,
                          "gpu_ms": gpu * 1e3, "cpu_ms": cpu * 1e3, "cpu_kind": kind + ", 1 thread", "corners": int(len(got))}), flush=True)
    got = api.find_blobs_int(dots)
    assert np.array_equal(got, po.find_blobs(dots))
    t0 = time.perf_counter()
    for _ in range(5):
        api.find_blobs_int(dots)
    gpu = (time.perf_counter() - t0) / 5
    cpu = None
    try:
        import cv2
        cv2.setNumThreads(1)
        p = cv2.SimpleBlobDetector_Params(); p.minArea = 20; p.maxArea = 80000; p.minDistBetweenBlobs = 5; p.blobColor = 0
        d = cv2.SimpleBlobDetector_create(p)
        t0 = time.perf_counter(); d.detect(dots); cpu = time.perf_counter() - t0
    except ImportError:
        pass
    print(json.dumps({"call": "find_blobs_from_image_array", "image": f"{w}x{h}", "gpu_ms": gpu * 1e3,
                      "cpu_ms": cpu * 1e3 if cpu else None, "cpu_kind": "cv2.SimpleBlobDetector, 1 thread", "blobs": int(len(got))}), flush=True)
