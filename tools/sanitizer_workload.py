import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mrgingham_b200 import api, synth
from oracle import pyoracle as po
frames = np.stack([synth.board_frame(1280, 720, 10, seed=1), synth.blurred_noise_frame(1280, 720, seed=2)])
det = api.Detector(max_frames=2, candidate_capacity=1<<18, max_points=1<<14)
xy, c = det.find_corners(frames, 0)
for i in range(2):
    w = po.find_corners(frames[i], 0)
    assert c[i] == len(w) and np.array_equal(xy[i,:c[i]], w)
small = np.stack([synth.circle_grid_frame(320, 240, 6, seed=3), synth.blob_frame(320, 240, seed=4)])
bx, bc = det.find_blobs(small)
for i in range(2):
    w = po.find_blobs(small[i]); assert bc[i] == len(w) and np.array_equal(bx[i,:bc[i]], w)
b = det.box_blur(small, 1); assert np.array_equal(b[0], po.box_blur(small[0], 1))
# pyramid fast paths (levels 1, 2 on 16-byte aligned 1280-wide frames), the generic level kernel, the dense
# response through the tiled kernel, the aligned 3x3 blur and the detector with on-device blur
for level in (1, 2, 3):
    xy, c2 = det.find_corners(frames, level)
    for i in range(2):
        w = po.find_corners(frames[i], level)
        assert c2[i] == len(w) and np.array_equal(xy[i, :c2[i]], w)
resp = det.chess_response(frames[:1])
assert np.array_equal(resp[0], po.chess_response_5(frames[0], fill=0))
assert np.array_equal(det.box_blur(frames[:1], 1)[0], po.box_blur(frames[0], 1))
det2 = api.Detector(max_frames=2, max_points=1 << 14, candidate_capacity=1 << 18, blur_radius=1)
xy, c3 = det2.find_corners(frames, 0)
for i in range(2):
    w = po.find_corners(po.box_blur(frames[i], 1), 0)
    assert c3[i] == len(w) and np.array_equal(xy[i, :c3[i]], w)
print("sanitizer workload ok", c, bc, c3)
# the --clahe chain (both apply kernels: tall and short tiles, aligned and odd widths) and whole boards
for img in (frames[0], synth.board_frame(403, 351, 10, seed=5), synth.noise_frame(131, 77, seed=6)):
    assert np.array_equal(det.preprocess(img[None], clahe=True, blur_radius=0)[0], po.normalize_clahe(img))
found, bxy, blv = det.find_boards(frames, gridn=10, level=-1)
assert found[0] >= 0 and found[1] < 0
print("preproc + boards ok", found)
