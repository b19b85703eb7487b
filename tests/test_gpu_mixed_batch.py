"""BASELINE.json configs[4]: a mixed-resolution batch, VGA to 4K and odd sizes, in ONE call
(mrg_b200_find_corners_mixed_batch; reference: the CLI's glob of images of any sizes,
mrgingham-from-image.cc:50-54). Every image's corner list must equal the oracle's, in the caller's order."""
import numpy as np
import pytest

from mrgingham_b200 import api, synth
from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

SIZES = [(640, 480), (1920, 1080), (1280, 720), (640, 480), (3840, 2160), (333, 217), (1920, 1080), (801, 603),
         (640, 480), (2560, 1440), (64, 48), (1920, 1080), (20, 20), (14, 30)]


def _images():
    out = []
    for i, (w, h) in enumerate(SIZES):
        if min(w, h) >= 200:
            out.append(synth.board_frame(w, h, 10 if i % 3 else 6, seed=100 + i))
        else:
            out.append(synth.blurred_noise_frame(w, h, seed=i))
    return out


@pytest.mark.parametrize("level", [0, 1])
def test_mixed_sizes_host(level):
    api._require_gpu()
    imgs = _images()
    det = api.Detector(max_frames=4, max_rows=2160, max_cols=3840, max_points=256)
    xy, counts = det.find_corners_mixed(imgs, level)
    for i, im in enumerate(imgs):
        want = po.find_corners(im, level)
        assert counts[i] == len(want), (i, SIZES[i])
        assert np.array_equal(xy[i, :min(counts[i], 256)], want[:256]), (i, SIZES[i])
    # pitched rows (views into wider arrays) and an empty list
    wide = [np.zeros((im.shape[0], im.shape[1] + 13), dtype=np.uint8) for im in imgs[:4]]
    views = []
    for wbuf, im in zip(wide, imgs[:4]):
        wbuf[:, 5:5 + im.shape[1]] = im
        views.append(wbuf[:, 5:5 + im.shape[1]])
    xy2, counts2 = det.find_corners_mixed(views, level)
    assert np.array_equal(counts2, counts[:4]) and all(np.array_equal(xy2[i, :counts[i]], xy[i, :counts[i]]) for i in range(4))
    xy3, counts3 = det.find_corners_mixed([], level)
    assert xy3.shape == (0, 256, 2) and counts3.shape == (0,)
    det.close()


def test_mixed_sizes_device():
    import torch
    api._require_gpu()
    imgs = _images()
    det = api.Detector(max_frames=8, max_rows=2160, max_cols=3840, max_points=256)
    want_xy, want_counts = det.find_corners_mixed(imgs, 0)
    # separate device tensors (gathered inside) ...
    dev = [torch.from_numpy(im).cuda() for im in imgs]
    xy, counts = det.find_corners_mixed(dev, 0)
    assert np.array_equal(counts, want_counts) and all(np.array_equal(xy[i, :counts[i]], want_xy[i, :counts[i]]) for i in range(len(imgs)))
    # ... and same-size frames that are slices of one tensor (read in place), between others
    stack = torch.from_numpy(np.stack([imgs[1], imgs[6], imgs[11]])).cuda()
    mixed = [dev[0], stack[0], stack[1], dev[4], stack[2]]
    xy2, counts2 = det.find_corners_mixed(mixed, 0)
    idx = [0, 1, 6, 4, 11]
    assert np.array_equal(counts2, want_counts[idx]) and all(np.array_equal(xy2[j, :counts2[j]], want_xy[i, :counts2[j]]) for j, i in enumerate(idx))
    det.close()
