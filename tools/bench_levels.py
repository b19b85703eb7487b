#!/usr/bin/env python
"""Per-level timing of the corner detector on 4K 14x14 frames (BASELINE.json configs[3], part i):
kernel times of the pyramid (K0), ChESS (K1) and clustering (K2) kernels from CUDA events, per level."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mrgingham_b200 import api, synth
from oracle import pyoracle as po

n, W, H = 256, 3840, 2160
base = [synth.board_frame(W, H, 14, seed=s) for s in range(4)]
frames = torch.from_numpy(np.stack([base[i % 4] for i in range(n)])).cuda()
det = api.Detector(max_frames=n, max_rows=H, max_cols=W, max_points=1024)
det.set_profiling(True)
for level in (0, 1, 2, 3):
    xy, counts = det.find_corners(frames, level)
    want = [po.find_corners(b, level) for b in base]
    ok = all(counts[i] == len(want[i % 4]) and np.array_equal(xy[i, :counts[i]], want[i % 4]) for i in range(n))
    ms = [[], [], []]
    for _ in range(3):
        det.find_corners(frames, level)
        for k in range(3):
            ms[k].append(det.last_kernel_ms(k)[0])
    k1, k2, k0 = (float(np.median(m)) for m in ms)
    print(json.dumps({"workload": f"{n} x {W}x{H} board 14x14, level {level}", "k0_pyramid_ms": k0, "k1_chess_ms": k1, "k2_cluster_ms": k2,
                      "full_res_gpix_s": n * W * H / ((k0 + k1) * 1e-3) / 1e9, "corners": int(len(want[0])), "identical_to_oracle": bool(ok)}), flush=True)
