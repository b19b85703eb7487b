"""GPU parity tests of the blob path through the C ABI: the CUDA
detector against the committed cv2.SimpleBlobDetector golden vectors and against the CPU oracle
(oracle/blob_oracle.c) on seeded frames, single-image, batched, device-resident and strided."""
import os

import numpy as np
import pytest

from mrgingham_b200 import synth

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "blobs_v1.npz"))
NAMES = sorted(k[4:] for k in GOLD.files if k.startswith("img/"))


@pytest.fixture(scope="module")
def api():
    from mrgingham_b200 import api as a
    a._require_gpu()
    return a


@pytest.fixture(scope="module")
def oracle():
    from oracle import pyoracle as po
    return po


@pytest.mark.parametrize("name", NAMES)
def test_blobs_match_golden(api, name):
    got = api.find_blobs_int(np.ascontiguousarray(GOLD["img/" + name]))
    want = GOLD["pts/" + name]
    assert got.shape == want.shape, (name, got.shape, want.shape)
    assert np.array_equal(got, want), name


def test_find_points_blobs_mirror(api):
    img = np.ascontiguousarray(GOLD["img/circles_vga_n10"])
    xy = api.find_points(img, blobs=True)
    assert np.array_equal(xy, GOLD["pts/circles_vga_n10"].astype(np.float64) * (1.0 / 1000))
    with pytest.raises(RuntimeError):
        api.find_points(img, image_pyramid_level=1, blobs=True)


def test_blobs_batch_matches_oracle(api, oracle):
    frames = np.stack([synth.circle_grid_frame(800, 608, 10, seed=40), synth.board_frame(800, 608, 10, seed=41),
                       synth.blob_frame(800, 608, seed=42), synth.blurred_noise_frame(800, 608, seed=43)])
    det = api.Detector(max_frames=3, max_points=4096)      # 4 frames in chunks of 3
    xy, counts = det.find_blobs(frames)
    for i in range(len(frames)):
        want = oracle.find_blobs(frames[i])
        assert counts[i] == len(want), (i, counts[i], len(want))
        assert np.array_equal(xy[i, :counts[i]], want), i
    assert counts[0] == 100
    det.close()


def test_blobs_device_resident_and_strided(api, oracle):
    import torch
    base = np.stack([synth.circle_grid_frame(640, 480, 9, seed=50 + s) for s in range(2)])
    wide = np.zeros((2, 480, 704), np.uint8)
    wide[:, :, 3:643] = base
    t = torch.from_numpy(wide).cuda()[:, :, 3:643]          # unaligned device view with a pitch
    det = api.Detector(max_frames=2, max_points=2048)
    xy, counts = det.find_blobs(t)
    xy2, counts2 = det.find_blobs(wide[:, :, 3:643])         # same thing from the host
    for i in range(2):
        want = oracle.find_blobs(np.ascontiguousarray(base[i]))
        assert counts[i] == len(want) and np.array_equal(xy[i, :counts[i]], want)
        assert counts2[i] == len(want) and np.array_equal(xy2[i, :counts2[i]], want)
    det.close()


def test_blobs_noise_overflows_default_scratch(api, oracle):
    # uniform noise: a huge number of tiny borders per threshold
    frames = synth.noise_frame(640, 480, seed=60)[None]
    det = api.Detector(max_frames=1, max_points=4096)
    xy, counts = det.find_blobs(frames)
    want = oracle.find_blobs(frames[0])
    assert counts[0] == len(want) and np.array_equal(xy[0, :counts[0]], want)
    det.close()


def test_blobs_busy_frame_inside_a_chunk(api, oracle):
    # one noise frame among ordinary ones: the chunk overflows its scratch and is redone frame by frame
    frames = np.stack([synth.circle_grid_frame(640, 480, 8, seed=61), synth.noise_frame(640, 480, seed=62),
                       synth.board_frame(640, 480, 10, seed=63)])
    det = api.Detector(max_frames=3, max_points=4096)
    xy, counts = det.find_blobs(frames)
    for i in range(3):
        want = oracle.find_blobs(frames[i])
        assert counts[i] == len(want) and np.array_equal(xy[i, :counts[i]], want), i
    xy, counts = det.find_blobs(frames[::2])          # and the workspace is back to normal afterwards
    for k, i in enumerate((0, 2)):
        want = oracle.find_blobs(frames[i])
        assert counts[k] == len(want) and np.array_equal(xy[k, :counts[k]], want)
    det.close()


def test_blobs_4k_board(api, oracle):
    frame = synth.board_frame(3840, 2160, 14, seed=70)
    got = api.find_blobs_int(frame)
    want = oracle.find_blobs(frame)
    assert len(want) > 50
    assert np.array_equal(got, want)


def test_blobs_tiny_and_empty(api, oracle):
    det = api.Detector(max_frames=4)
    for (w, h) in ((1, 1), (2, 3), (5, 5), (33, 7), (64, 1)):
        frames = np.stack([synth.noise_frame(w, h, seed=s) for s in range(3)])
        xy, counts = det.find_blobs(frames)
        for i in range(3):
            want = oracle.find_blobs(frames[i])
            assert counts[i] == len(want) and np.array_equal(xy[i, :counts[i]], want)
    xy, counts = det.find_blobs(np.zeros((0, 16, 16), np.uint8))
    assert len(counts) == 0
    det.close()


def test_blobs_random_frames_match_oracle(api, oracle):
    # the same kinds of seeded random frames the oracle is pinned on against cv2 (tests/test_blob_oracle.py)
    rng = np.random.default_rng(78)
    for t in range(24):
        w, h = int(rng.integers(40, 700)), int(rng.integers(40, 500))
        kind = t % 4
        if kind == 0:
            img = synth.blob_frame(w, h, seed=2000 + t, nblobs=int(rng.integers(3, 80)))
        elif kind == 1:
            img = synth.circle_grid_frame(max(w, 160), max(h, 160), int(rng.integers(3, 10)), seed=2000 + t,
                                          noise_sigma=float(rng.uniform(0, 6)), blur=bool(t & 4))
        elif kind == 2:
            img = synth.blurred_noise_frame(w, h, seed=2000 + t, passes=int(rng.integers(1, 4)))
        else:
            img = synth.board_frame(max(w, 200), max(h, 160), int(rng.integers(4, 9)), seed=2000 + t)
        img = np.ascontiguousarray(img)
        assert np.array_equal(api.find_blobs_int(img), oracle.find_blobs(img)), (t, kind, img.shape)


def test_blobs_wide_and_tall_frames(api, oracle):
    """coordinates beyond 4096 in x and in y (state keys, queue entries, cut rows / columns far from the origin)"""
    po = oracle
    rng = np.random.default_rng(21)
    wide = synth.blob_frame(7700, 300, seed=41, nblobs=400)
    tall = np.ascontiguousarray(synth.blob_frame(5100, 260, seed=42, nblobs=300).T)
    ring = np.full((700, 9000), 200, dtype=np.uint8); ring[40:660, 30:8970] = 60; ring[80:620, 70:8930] = 200      # a border 18 000 states long
    ring = np.clip(ring.astype(np.int16) + rng.integers(-3, 4, size=ring.shape), 0, 255).astype(np.uint8)
    for img in (wide, tall, ring):
        assert np.array_equal(api.find_blobs_int(img), po.find_blobs(img)), img.shape
