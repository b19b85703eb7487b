"""16-bit input of the reference CLI (mrgingham-from-image.cc:83-93): normalize to [0,65535] +
CLAHE on 16 bits (with --clahe), then convertTo(CV_8U, 255./65535.). Oracle pinned against cv2 (CPU); CUDA kernels
against the oracle through the C ABI (GPU)."""
import numpy as np
import pytest

from mrgingham_b200 import api, synth
from oracle import pyoracle as po


def _images16(rng, count):
    for t in range(count):
        w = int(rng.choice([8, 16, 31, 64, 97, 160, 321, 640])); h = int(rng.choice([8, 9, 24, 50, 64, 121, 240, 480]))
        kind = t % 4
        if kind == 0:
            yield rng.integers(0, 65536, size=(h, w)).astype(np.uint16)
        elif kind == 1:
            b = synth.board_frame(max(w, 64), max(h, 64), 6, seed=int(rng.integers(1000)))[:h, :w].astype(np.float64)
            yield np.clip(b * rng.uniform(20, 257) + rng.normal(0, 200, size=b.shape) + rng.uniform(0, 3000), 0, 65535).astype(np.uint16)
        elif kind == 2:
            yield np.full((h, w), int(rng.integers(0, 65536)), dtype=np.uint16)
        else:
            lo = int(rng.integers(0, 60000))
            yield rng.integers(lo, lo + int(rng.integers(2, 3000)), size=(h, w)).astype(np.uint16)


def test_oracle16_equals_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    cl = cv2.createCLAHE(clipLimit=8.0)
    for img in _images16(rng, 80):
        n = cv2.normalize(img, None, 0, 65535, cv2.NORM_MINMAX)
        assert np.array_equal(po.normalize16(img), n), img.shape
        assert np.array_equal(po.clahe16(img), cl.apply(img)), img.shape
        # convertTo(CV_8U, 255./65535.) is what normalize(0, 255, dtype=CV_8U) runs on an image spanning [0, 65535]
        c = cl.apply(n); c[0, 0] = 0; c[0, -1] = 65535
        assert np.array_equal(po.convert16to8(c), cv2.normalize(c, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8U)), img.shape
    ramp = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    assert np.array_equal(po.convert16to8(ramp), cv2.normalize(ramp, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8U))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_gpu_preprocess16_equals_oracle():
    api._require_gpu()
    rng = np.random.default_rng(12)
    det = api.Detector(max_frames=4)
    for img in _images16(rng, 40):
        for clahe in (False, True):
            want = po.chain16(img, clahe)
            got = det.preprocess16(img[None], clahe=clahe, blur_radius=0)[0]
            assert np.array_equal(got, want), (img.shape, clahe)
        got = det.preprocess16(img[None], clahe=True, blur_radius=1)[0]
        assert np.array_equal(got, po.box_blur(po.chain16(img, True), 1)), img.shape
    # a batch, pitched rows, device-resident tensors, a 4K frame
    import torch
    raw = np.stack([np.clip(synth.board_frame(800, 608, 10, seed=s).astype(np.float64) * 200 + 1000, 0, 65535).astype(np.uint16) for s in range(5)])
    want = np.stack([po.chain16(f, True) for f in raw])
    assert np.array_equal(det.preprocess16(raw, clahe=True, blur_radius=0), want)
    wide = np.zeros((5, 608, 811), dtype=np.uint16); wide[:, :, 3:803] = raw
    assert np.array_equal(det.preprocess16(wide[:, :, 3:803], clahe=True, blur_radius=0), want)
    t = torch.from_numpy(raw.view(np.int16)).cuda()
    assert np.array_equal(det.preprocess16(t, clahe=True, blur_radius=0).cpu().numpy(), want)
    assert np.array_equal(det.preprocess16(t, clahe=False, blur_radius=0).cpu().numpy(), np.stack([po.chain16(f, False) for f in raw]))
    big = np.clip(synth.board_frame(3840, 2160, 10, seed=3).astype(np.float64) * 257 + rng.normal(0, 300, size=(2160, 3840)), 0, 65535).astype(np.uint16)
    assert np.array_equal(det.preprocess16(big[None], clahe=True, blur_radius=1)[0], po.box_blur(po.chain16(big, True), 1))
    # the chain feeds the detector: corners of the 8-bit image the CLI would have computed
    got8 = det.preprocess16(big[None], clahe=False, blur_radius=1)[0]
    assert np.array_equal(api.find_chessboard_corners_int(got8, 0), po.find_corners(po.box_blur(po.chain16(big, False), 1), 0))
    det.close()
