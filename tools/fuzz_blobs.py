#!/usr/bin/env python
"""Randomised sweep of the GPU blob detector against the CPU oracle (oracle/blob_oracle.c, itself swept against
cv2.SimpleBlobDetector 4.13.0 on 20 800 frames): random frames of many kinds and sizes through
mrg_b200_find_blobs_batch, every frame's PointInt list compared in value and order. Prints one summary line.
usage: python tools/fuzz_blobs.py [--seconds 120] [--seed 1]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def frame(kind, w, h, seed, rng):
    from mrgingham_b200 import synth
    if kind == 0:
        return synth.blob_frame(w, h, seed=seed)
    if kind == 1:
        return synth.circle_grid_frame(w, h, int(rng.integers(4, 12)), seed=seed, noise_sigma=float(rng.uniform(0, 6)))
    if kind == 2:
        return synth.board_frame(w, h, int(rng.integers(4, 14)), seed=seed, noise_sigma=float(rng.uniform(0, 8)), blur=bool(rng.integers(0, 2)))
    if kind == 3:
        return synth.blurred_noise_frame(w, h, seed=seed, passes=int(rng.integers(1, 5)))
    if kind == 4:
        return synth.checker_frame(w, h, period=int(rng.integers(5, 40)), seed=seed)
    if kind == 5:
        return synth.noise_frame(w, h, seed=seed)
    # thresholded smooth noise: large regions with holes, nested, touching the frame
    f = synth.blurred_noise_frame(w, h, seed=seed, passes=4).astype(np.float32)
    lo, hi = np.percentile(f, [rng.uniform(20, 45), rng.uniform(55, 80)])
    return np.where(f < lo, 20, np.where(f > hi, 235, 128)).astype(np.uint8) + rng.integers(0, 12, (h, w)).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    from mrgingham_b200 import api
    from oracle import pyoracle as po
    rng = np.random.default_rng(a.seed)
    sizes = [(320, 240), (333, 217), (640, 480), (800, 608), (517, 389), (1280, 720), (96, 64), (257, 1030), (1030, 129)]
    dets = {}
    t0 = time.perf_counter()
    frames = blobs = bad = 0
    kinds = [0] * 7
    while time.perf_counter() - t0 < a.seconds:
        w, h = sizes[int(rng.integers(0, len(sizes)))]
        n = int(rng.integers(1, 13))
        ks = [int(rng.choice(7, p=[0.2, 0.2, 0.2, 0.15, 0.03, 0.02, 0.2])) for _ in range(n)]
        batch = np.stack([frame(k, w, h, int(rng.integers(0, 1 << 30)), rng) for k in ks])
        det = dets.setdefault((w, h), api.Detector(max_frames=8, max_rows=h, max_cols=w, max_points=1 << 15))
        xy, counts = det.find_blobs(batch)
        for i in range(n):
            want = po.find_blobs(batch[i])
            frames += 1; kinds[ks[i]] += 1; blobs += len(want)
            if counts[i] != len(want) or not np.array_equal(xy[i, :min(counts[i], 1 << 15)], want[:1 << 15]):
                bad += 1
                np.save(os.path.join(ROOT, "gpurun_out", f"blob_fuzz_fail_{frames}.npy"), batch[i])
    print(f"blob fuzz: {frames} frames ({kinds} of kinds blobs/circles/board/blurred noise/checker/noise/thresholded) "
          f"of {len(sizes)} sizes, {blobs} blobs, {bad} frames differ from the oracle, seed {a.seed}, {time.perf_counter() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
