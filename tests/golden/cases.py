"""The image set behind tests/golden/golden_v1.npz (shared by make_golden.py and the tests)."""
import numpy as np

from mrgingham_b200 import synth


def golden_images():
    """name -> uint8 image. Small on purpose: the images themselves are committed in the .npz."""
    out = {}
    out["board_vga_n10"] = synth.board_frame(640, 480, 10, seed=0)
    out["board_small_n10"] = synth.board_frame(320, 240, 10, seed=3)
    out["board_odd_n10"] = synth.board_frame(403, 351, 10, seed=5)      # 403 = 3 mod 4, 351 = 3 mod 4
    out["board_n14"] = synth.board_frame(512, 512, 14, seed=7)
    out["board_noblur_noisy"] = synth.board_frame(400, 300, 10, seed=9, noise_sigma=6.0, blur=False)
    out["noise"] = synth.noise_frame(200, 160, seed=11)
    out["blurred_noise"] = synth.blurred_noise_frame(240, 200, seed=13)
    out["checker8"] = synth.checker_frame(256, 192, period=8, seed=15)
    out["checker5_noisy"] = synth.checker_frame(160, 128, period=5, seed=17, noise_sigma=8.0)
    out["blobs"] = synth.blob_frame(300, 200, seed=19)
    out["tiny_15x40"] = synth.noise_frame(15, 40, seed=21)              # narrower than the ring
    out["tiny_17x17"] = synth.noise_frame(17, 17, seed=23)
    out["flat"] = np.full((64, 64), 128, dtype=np.uint8)
    return out


LEVELS = (0, 1, 2, 3)


def refine_chain(find_fn, refine_fn, image, start_level):
    """find at start_level, then refine down to level 0 the way mrgingham.cc:81-99 does.
    Returns (xy_double, levels, [nrefined per level])."""
    pts = find_fn(image, start_level)
    xy = pts.astype(np.float64) / 1000.0
    levels = np.full(len(xy), start_level, dtype=np.int8)
    counts = []
    level = start_level
    while level > 0:
        level -= 1
        n, xy, levels = refine_fn(image, level, xy, levels)
        counts.append(n)
        if n <= 0:
            break
    return xy, levels, np.asarray(counts, dtype=np.int32)


def blob_golden_images():
    """name -> uint8 image for tests/golden/blobs_v1.npz (the blob path)."""
    out = {}
    out["circles_vga_n10"] = synth.circle_grid_frame(640, 480, 10, seed=1)
    out["circles_small_n7"] = synth.circle_grid_frame(331, 257, 7, seed=2)
    out["circles_noblur"] = synth.circle_grid_frame(400, 300, 8, seed=3, noise_sigma=5.0, blur=False)
    out["board_vga_n10"] = synth.board_frame(640, 480, 10, seed=0)       # black squares qualify as blobs
    out["board_n14"] = synth.board_frame(512, 512, 14, seed=7)
    out["blobs"] = synth.blob_frame(300, 200, seed=19)
    out["blobs_dense"] = synth.blob_frame(257, 193, seed=29, nblobs=120)
    out["blurred_noise"] = synth.blurred_noise_frame(240, 200, seed=13)
    out["noise"] = synth.noise_frame(120, 90, seed=11)
    out["checker8"] = synth.checker_frame(256, 192, period=8, seed=15)
    out["tiny_9x7"] = synth.noise_frame(9, 7, seed=31)
    out["flat"] = np.full((64, 64), 128, dtype=np.uint8)
    out["all_dark"] = np.full((40, 50), 10, dtype=np.uint8)
    return out
