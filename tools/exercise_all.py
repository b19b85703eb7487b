#!/usr/bin/env python
"""Launches every production kernel of the library once or twice on 4K frames (device-resident), so that one
`ncu --set full` run can capture them all (tools/measure_round.sh; summaries by tools/ncu_all_summary.py)."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    a = ap.parse_args()
    import torch
    from mrgingham_b200 import api, synth
    W, H, n = a.width, a.height, a.frames
    base = [synth.board_frame(W, H, 10, seed=s) for s in range(4)]
    frames = torch.from_numpy(np.stack([base[i % 4] for i in range(n)])).cuda()
    det = api.Detector(max_frames=n, max_rows=H, max_cols=W, max_points=256)
    for level in (0, 1, 2, 3):                                  # K1 cascade + K2; K0 level 1 / 2 / generic
        xy, counts = det.find_corners(frames, level)
    # K2r: refine the level-1 corners at level 0
    xy1, c1 = det.find_corners(frames, 1)
    pts = np.zeros((n, 100, 2)); lv = np.full((n, 100), 1, dtype=np.int8)
    for i in range(n):
        k = min(100, int(c1[i])); pts[i, :k] = xy1[i, :k] / 1000.0
    det.refine_corners(frames, 0, pts, lv)
    det.chess_response(np.stack(base * 4))                      # dense response (tiled, TMA): host frames in, int16 out
    det.box_blur(frames, 1); det.box_blur(frames[:16], 2)       # 3x3 blur (HBM-bound kernel), generic radius
    det.preprocess(frames, clahe=True, blur_radius=0)           # min/max, normalisation table, CLAHE tables, CLAHE apply
    raw16 = torch.from_numpy((np.stack([base[i % 4] for i in range(16)]).astype(np.int32) * 200 + 500).astype(np.uint16).view(np.int16)).cuda()
    det.preprocess16(raw16, clahe=True, blur_radius=0)           # 16-bit input: min/max, 65536-bin histograms, tables, apply + convert
    det.preprocess16(raw16, clahe=False, blur_radius=0)
    det.find_corners_mixed([frames[0], frames[1][:1080, :1920], frames[2]], 0)      # gather kernel
    circles = torch.from_numpy(np.stack([synth.circle_grid_frame(W, H, 10, seed=s % 4) for s in range(n)])).cuda()
    det.find_blobs(circles)                                     # B1, scan, walk, points, contour kernels
    det.find_blobs(frames)
    torch.cuda.synchronize()
    det.close()
    print("exercised")


if __name__ == "__main__":
    main()
