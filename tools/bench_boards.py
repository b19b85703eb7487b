#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[3] part i): whole boards over 4K frames --
corners at the auto-selected pyramid level (3,2,1,0 until the grid finder succeeds), the grid finder on host
threads, and refinement down to level 0 -- through mrg_b200_find_boards_batch, against the CPU path (the
reference's corner/refinement code where it was compiled, else its port, plus the grid oracle's graph replaced
by the library's own host grid finder, which is what costs time there) on one host thread.
Prints one JSON line; every distinct frame's result is compared with the oracle pipeline in the same run."""
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
    ap.add_argument("--level", type=int, default=-1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--base-frames", type=int, default=4)
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--host", action="store_true", help="frames start in host memory (copied inside the timed region)")
    a = ap.parse_args()
    import torch
    from mrgingham_b200 import api, synth
    from oracle import pyoracle as po
    W, H = a.width, a.height
    base = [synth.board_frame(W, H, a.gridn, seed=s) for s in range(a.base_frames)]
    stack = np.stack([base[i % len(base)] for i in range(a.frames)])
    frames = stack if a.host else torch.from_numpy(stack).cuda()
    det = api.Detector(max_frames=a.chunk, max_rows=H, max_cols=W, max_points=2048)
    for _ in range(a.warmup):
        found, xy, lv = det.find_boards(frames, gridn=a.gridn, level=a.level)

    # the same pipeline from the CPU checkers, one thread (grid step: the library's host grid finder, checked
    # against the grid oracle by the tests)
    def cpu_board(img):
        for L in ([3, 2, 1, 0] if a.level < 0 else [a.level]):
            grid = api.find_grid_from_points(po.find_corners(img, L), a.gridn)
            if grid is None:
                continue
            l = np.full(len(grid), L, np.int8)
            for r in range(L - 1, -1, -1):
                n, grid, l = po.refine_corners(img, r, grid, l)
                if n <= 0:
                    break
            return L, grid, l
        return -1, None, None
    t0 = time.perf_counter()
    want = [cpu_board(b) for b in base]
    t_cpu = (time.perf_counter() - t0) / len(base)
    ok = all(found[i] == want[i % len(base)][0] and (found[i] < 0 or (np.array_equal(xy[i], want[i % len(base)][1]) and
             np.array_equal(lv[i], want[i % len(base)][2]))) for i in range(a.frames))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        det.find_boards(frames, gridn=a.gridn, level=a.level)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    px = W * H / 1e6
    print(json.dumps({
        "metric": "boards/s: corners + grid + refinement over 4K frames (find_chessboard_from_image_array)", "unit": "boards/s",
        "value": a.frames / dt, "mpix_s": a.frames * px / dt, "ms_per_step": dt * 1e3, "steps": a.steps,
        "config": {"workload": f"{a.frames} x {W}x{H} board n={a.gridn}, level {a.level}, "
                               f"{'host' if a.host else 'device-resident'} frames", "found_levels": sorted(set(int(f) for f in found))},
        "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "boards/s", "cores": 1, "kind": "port",
                         "sample": f"{len(base)} frames"},
        "parity": {"frames_checked": a.frames, "identical_to_oracle_pipeline": bool(ok)}}))


if __name__ == "__main__":
    main()
