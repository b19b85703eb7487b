#!/usr/bin/env python
"""BASELINE.json configs[4]: K1 (ChESS + candidate emission) over a mixed-resolution set, VGA to 8K.
One JSON line per resolution: achieved algorithmic GB/s of the K1 launch (CUDA events on its stream,
via mrg_b200_last_kernel_ms) against the measured HBM peak, with the corner lists of every distinct
frame checked against the oracle in the same run. Secondary to bench.py's headline line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIZES = ((640, 480), (1280, 720), (1920, 1080), (2560, 1440), (3840, 2160), (7680, 4320))


def main():
    import torch
    from mrgingham_b200 import api, synth
    from oracle import pyoracle as po
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    target_bytes = 2 << 30                              # ~2 GB of frames per launch at every size
    for (w, h) in SIZES:
        n = max(8, min(4096, target_bytes // (w * h)))
        base = [synth.board_frame(w, h, 10, seed=s) for s in range(4)]
        frames = torch.from_numpy(np.stack([base[i % 4] for i in range(n)])).cuda()
        det = api.Detector(max_frames=n, max_rows=h, max_cols=w, max_points=1024)
        det.set_profiling(True)
        xy, counts = det.find_corners(frames, 0)
        want = [po.find_corners(b, 0) for b in base]
        ok = all(counts[i] == len(want[i % 4]) and np.array_equal(xy[i, :counts[i]], want[i % 4]) for i in range(n))
        ms = []
        for _ in range(5):
            det.find_corners(frames, 0)
            ms.append(det.last_kernel_ms(0)[0])
        k1 = float(np.median(ms))
        gbs = n * w * h / (k1 * 1e-3) / 1e9
        print(json.dumps({"workload": f"{n} x {w}x{h} board 10x10, level 0", "k1_ms": k1, "achieved_gbs": gbs, "peak_gbs": peak,
                          "frac": gbs / peak, "corners": int(len(want[0])), "identical_to_oracle": bool(ok)}), flush=True)
        det.close()
        del frames
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
