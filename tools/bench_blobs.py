#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[3], part ii): the blob detector on 4K 14x14 frames.
Not the headline bench.py line; prints one JSON line with the GPU rate through the C ABI
(device-resident frames), the kernel time, and the CPU baselines timed on this box's host:
cv2.SimpleBlobDetector (the reference's own dependency, kind "reference") and the C oracle ("port").
Every frame's result is compared with the oracle in the same run."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--gridn", type=int, default=14)
    ap.add_argument("--kind", default="board", choices=["board", "circles"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--base-frames", type=int, default=4)
    ap.add_argument("--chunk", type=int, default=256)
    a = ap.parse_args()
    import torch
    from mrgingham_b200 import api, synth
    from oracle import pyoracle as po
    W, H = a.width, a.height
    gen = synth.board_frame if a.kind == "board" else synth.circle_grid_frame
    base = [gen(W, H, a.gridn, seed=s) for s in range(a.base_frames)]
    frames = torch.from_numpy(np.stack([base[i % len(base)] for i in range(a.frames)])).cuda()
    det = api.Detector(max_frames=a.chunk, max_rows=H, max_cols=W, max_points=4096)
    det.set_profiling(True)
    xy, counts = det.find_blobs(frames)
    for _ in range(a.warmup - 1):
        det.find_blobs(frames)
    want = [po.find_blobs(b) for b in base]
    ok = all(counts[i] == len(want[i % len(base)]) and np.array_equal(xy[i, :counts[i]], want[i % len(base)]) for i in range(a.frames))
    torch.cuda.synchronize()
    t0 = time.perf_counter(); kms = 0.0
    for _ in range(a.steps):
        det.find_blobs(frames)
        kms += det.last_kernel_ms(3)[0]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    # CPU baselines, single thread each, on the distinct frames
    t0 = time.perf_counter()
    for b in base:
        po.find_blobs(b)
    t_port = (time.perf_counter() - t0) / len(base)
    t_cv2 = None
    try:
        import cv2
        cv2.setNumThreads(1)
        p = cv2.SimpleBlobDetector_Params(); p.minArea = 20; p.maxArea = 80000; p.minDistBetweenBlobs = 5; p.blobColor = 0
        d = cv2.SimpleBlobDetector_create(p)
        t0 = time.perf_counter()
        for b in base:
            d.detect(b)
        t_cv2 = (time.perf_counter() - t0) / len(base)
    except ImportError:
        pass
    px = W * H / 1e6
    print(json.dumps({
        "metric": "Mpix/s blob detector (find_blobs.cc = cv::SimpleBlobDetector) over 4K frames", "unit": "Mpix/s",
        "value": a.frames * px / dt, "frames_per_s": a.frames / dt, "ms_per_step": dt * 1e3, "steps": a.steps,
        "config": {"workload": f"{a.frames} x {W}x{H} {a.kind} n={a.gridn}, 17 thresholds, device-resident frames", "blobs_per_frame": int(len(want[0]))},
        "kernel_ms_per_step": kms / a.steps,
        "cpu_baseline": {"reference_cv2_1thread_mpix_s": (px / t_cv2) if t_cv2 else None, "port_oracle_1thread_mpix_s": px / t_port,
                         "cores": 1, "sample": f"{len(base)} frames each"},
        "parity": {"frames_checked": a.frames, "identical_to_oracle": bool(ok)}}))


if __name__ == "__main__":
    main()
