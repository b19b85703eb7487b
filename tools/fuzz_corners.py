#!/usr/bin/env python
"""Randomised sweep of the GPU corner detector against the CPU oracle (oracle/mrg_oracle.c, itself tied to the compiled
reference by tests/test_oracle_vs_ref.py): random frames of many kinds and sizes at pyramid levels 0-3 through
mrg_b200_find_corners_batch (uniform batches) and mrg_b200_find_corners_mixed_batch (all sizes in one call), every
frame's PointInt list compared in value and order. Prints one summary line.
usage: python tools/fuzz_corners.py [--seconds 120] [--seed 1]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.fuzz_blobs import frame          # noqa: E402  (the same frame generators)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    from mrgingham_b200 import api
    from oracle import pyoracle as po
    rng = np.random.default_rng(a.seed)
    sizes = [(320, 240), (333, 217), (640, 480), (800, 608), (517, 389), (1280, 720), (96, 64), (257, 1030), (1030, 129), (1920, 1080)]
    MP = 1 << 14
    det = api.Detector(max_frames=8, max_points=MP, candidate_capacity=1 << 19)
    t0 = time.perf_counter()
    frames = corners = bad = 0
    levels = [0] * 4
    rounds = 0
    while time.perf_counter() - t0 < a.seconds:
        rounds += 1
        level = int(rng.choice(4, p=[0.55, 0.2, 0.15, 0.1]))
        if rounds % 3:
            w, h = sizes[int(rng.integers(0, len(sizes)))]
            n = int(rng.integers(1, 13))
            batch = [frame(int(rng.choice(7, p=[0.1, 0.1, 0.4, 0.15, 0.05, 0.05, 0.15])), w, h, int(rng.integers(0, 1 << 30)), rng) for _ in range(n)]
            xy, counts = det.find_corners(np.stack(batch), level)
        else:
            batch = []
            for _ in range(int(rng.integers(2, 10))):
                w, h = sizes[int(rng.integers(0, len(sizes)))]
                batch.append(frame(int(rng.choice(7, p=[0.1, 0.1, 0.4, 0.15, 0.05, 0.05, 0.15])), w, h, int(rng.integers(0, 1 << 30)), rng))
            xy, counts = det.find_corners_mixed(batch, level)
        for i, img in enumerate(batch):
            want = po.find_corners(img, level)
            frames += 1; levels[level] += 1; corners += len(want)
            if counts[i] != len(want) or not np.array_equal(xy[i, :min(counts[i], MP)], want[:MP]):
                bad += 1
                np.save(os.path.join(ROOT, "gpurun_out", f"corner_fuzz_fail_{frames}_L{level}.npy"), img)
    print(f"corner fuzz: {frames} frames of {len(sizes)} sizes (levels 0..3: {levels}; every third batch mixed-size), "
          f"{corners} corners, {bad} frames differ from the oracle, seed {a.seed}, {time.perf_counter() - t0:.0f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
