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
# round 2: 16-bit input, mixed-size batch (gather kernel), parallel refinement, the segment-based blob path on a frame
# with cut rows / columns (> 256 pixels each way), noise (queue regrowth)
rng = np.random.default_rng(7)
img16 = np.clip(synth.board_frame(403, 351, 10, seed=8).astype(np.float64) * 200 + rng.normal(0, 100, size=(351, 403)) + 900, 0, 65535).astype(np.uint16)
for cl in (False, True):
    assert np.array_equal(det.preprocess16(img16[None], clahe=cl, blur_radius=0)[0], po.chain16(img16, cl))
mixed = [frames[0], synth.board_frame(403, 351, 10, seed=9), frames[1], synth.noise_frame(131, 77, seed=10)]
mxy, mc = det.find_corners_mixed(mixed, 0)
for i, im in enumerate(mixed):
    w = po.find_corners(im, 0); assert mc[i] == len(w) and np.array_equal(mxy[i, :mc[i]], w[:1 << 14])
_, xy2 = po.find_corners(frames[0], 1, want_double=True); lv = np.full(len(xy2), 1, dtype=np.int8)
n_o, xy_o, lv_o = po.refine_corners(frames[0], 0, xy2, lv)
n_g, xy_g, lv_g = api.refine_chessboard_corners(frames[0], 0, xy2, lv)
assert n_g == n_o and np.array_equal(xy_g, xy_o) and np.array_equal(lv_g, lv_o)
for im in (synth.circle_grid_frame(1280, 720, 10, seed=11), synth.blob_frame(640, 600, seed=12), synth.noise_frame(300, 280, seed=13)):
    assert np.array_equal(api.find_blobs_int(im), po.find_blobs(im))
print("round-2 paths ok", mc, n_g)
# the pipelined board finder (chunks of >= 16 frames: corner passes over all frames ahead of the grid searches, refinement on
# the helper detector from a second host thread) against the pass-by-pass form
small_boards = np.stack([synth.board_frame(640, 480, 6, seed=20 + s, noise_sigma=(8.0 if s % 3 == 0 else 2.0)) for s in range(15)] +
                        [synth.noise_frame(640, 480, seed=40)])
det3 = api.Detector(max_frames=16, max_points=512)
f1, xy1, lv1 = det3.find_boards(small_boards, gridn=6, level=-1)
os.environ["MRG_B200_BOARDS_SERIAL"] = "1"
f0, xy0, lv0 = det3.find_boards(small_boards, gridn=6, level=-1)
del os.environ["MRG_B200_BOARDS_SERIAL"]
ok = f0 >= 0
assert np.array_equal(f0, f1) and np.array_equal(xy0[ok], xy1[ok]) and np.array_equal(lv0[ok], lv1[ok]) and f0[15] < 0
print("pipelined boards ok", f1)
