"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU
oracle and the committed golden vectors (generated from the reference itself). Bit-exact bar:
integer corner lists identical in value AND order; dense responses identical; refined doubles
bit-identical (north_star asks for 1e-3 px; we hold the stronger bar and state it here)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

from mrgingham_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def m():
    import mrgingham_b200
    assert mrgingham_b200.lib().mrg_b200_device_count() > 0, "no CUDA device: these tests need the B200"
    return mrgingham_b200


@pytest.fixture(scope="module")
def api(m):
    from mrgingham_b200 import api
    return api


def _names(golden):
    return sorted({k.split("/")[0] for k in golden.files})


def _extra_images():
    yield "board_720p", synth.board_frame(1280, 720, 10, seed=100)
    yield "board_n14_1080p", synth.board_frame(1920, 1080, 14, seed=101)
    yield "board_odd", synth.board_frame(611, 457, 10, seed=102)
    yield "noise", synth.noise_frame(333, 222, seed=103)
    yield "blurred_noise", synth.blurred_noise_frame(320, 256, seed=104, passes=1)
    yield "checker6", synth.checker_frame(300, 260, period=6, seed=105)
    yield "checker11", synth.checker_frame(480, 270, period=11, seed=106, noise_sigma=3.0)
    yield "blobs", synth.blob_frame(400, 300, seed=107)
    rng = np.random.default_rng(108)
    for i in range(10):
        w, h = int(rng.integers(1, 140)), int(rng.integers(1, 140))
        yield f"rand_{w}x{h}", synth.blurred_noise_frame(w, h, seed=200 + i, passes=1)


# ---------------------------------------------------------------------------------------------
# A1: dense response (mrgingham_ChESS_response_5)
# ---------------------------------------------------------------------------------------------
def test_dense_response_matches_golden(m, golden):
    for name in _names(golden):
        img = golden[f"{name}/image"]
        got = m.ChESS_response_5(img)
        assert np.array_equal(got, golden[f"{name}/response"]), name


def test_dense_response_leaves_border_untouched(m, api, oracle):
    img = synth.board_frame(200, 150, 10, seed=1)
    out = np.full(img.shape, -999, dtype=np.int16)
    api.lib().mrgingham_ChESS_response_5(out.ctypes.data_as(api._i16p), img.ctypes.data_as(api._u8p), 200, 150, 200)
    assert np.array_equal(out, oracle.chess_response_5(img, fill=-999))


def test_dense_response_strided_and_batched(m, oracle):
    big = synth.board_frame(700, 500, 10, seed=110)
    view = big[10:490, 20:660]
    assert np.array_equal(m.ChESS_response_5(view), oracle.chess_response_5(view, fill=0))
    stack = np.stack([synth.noise_frame(96, 80, seed=s) for s in range(6)]).reshape(2, 3, 80, 96)
    got = m.ChESS_response_5(stack)
    for i in range(2):
        for j in range(3):
            assert np.array_equal(got[i, j], oracle.chess_response_5(stack[i, j], fill=0))


# ---------------------------------------------------------------------------------------------
# A2: pyramid level
# ---------------------------------------------------------------------------------------------
def test_pyramid_matches_oracle(m, oracle):
    rng = np.random.default_rng(5)
    for (w, h) in ((43, 35), (47, 39), (64, 48), (99, 77), (403, 351), (640, 480), (1920, 1080)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        for level in (1, 2, 3, 4):
            got, want = m.pyramid_level(img, level), oracle.pyramid(img, level)
            assert got.shape == want.shape and np.array_equal(got, want), (w, h, level)


# ---------------------------------------------------------------------------------------------
# A3-A7: find_chessboard_corners_from_image_array
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("level", cases.LEVELS)
def test_corners_match_golden(api, golden, level):
    for name in _names(golden):
        img = golden[f"{name}/image"]
        got = api.find_chessboard_corners_int(img, level)
        want = golden[f"{name}/corners_L{level}"]
        assert got.shape == want.shape and np.array_equal(got, want), (name, level)


def test_corners_match_oracle_wider_set(api, oracle):
    for name, img in _extra_images():
        for level in (0, 1, 2, 3):
            got, want = api.find_chessboard_corners_int(img, level), oracle.find_corners(img, level)
            assert got.shape == want.shape and np.array_equal(got, want), (name, level)


def test_find_points_python_surface(m, oracle):
    img = synth.board_frame(640, 480, 10, seed=0)
    pts = m.find_points(img)
    want = oracle.find_corners(img, 0)
    assert pts.shape == (100, 2) and pts.dtype == np.float64
    assert np.array_equal(pts, (1.0 / 1000.0) * want.astype(np.float64))
    assert m.find_points(np.full((64, 64), 128, np.uint8)).shape == (0, 2)
    # reference error paths: bad level, non-continuous at level 0 -> nothing
    assert m.find_points(img, image_pyramid_level=-1).shape == (0, 2)
    assert m.find_points(img, image_pyramid_level=11).shape == (0, 2)
    big = synth.board_frame(700, 500, 10, seed=110)
    assert m.find_points(big[10:490, 20:660], image_pyramid_level=0).shape == (0, 2)
    v = big[10:490, 20:660]
    assert np.array_equal(m.find_points(v, image_pyramid_level=1), (1.0 / 1000.0) * oracle.find_corners(v, 1).astype(np.float64))
    with pytest.raises(RuntimeError):
        m.find_points(img, image_pyramid_level=1, blobs=True)
    with pytest.raises(RuntimeError):
        m.find_points(img.astype(np.float32))


def test_config_sizes_1080p_4k(api, oracle):
    for (w, h, n, seed) in ((1920, 1080, 10, 31), (3840, 2160, 10, 32), (3840, 2160, 14, 33)):
        img = synth.board_frame(w, h, n, seed=seed)
        for level in (0, 1, 2, 3):
            got, want = api.find_chessboard_corners_int(img, level), oracle.find_corners(img, level)
            assert np.array_equal(got, want), (w, h, n, level)
        assert len(oracle.find_corners(img, 0)) == n * n


# ---------------------------------------------------------------------------------------------
# A8: refinement
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("start", (1, 2, 3))
def test_refine_chain_matches_golden(api, golden, start):
    for name in _names(golden):
        img = golden[f"{name}/image"]
        xy, levels, counts = cases.refine_chain(api.find_chessboard_corners_int, api.refine_chessboard_corners, img, start)
        assert np.array_equal(counts, golden[f"{name}/refine_from_L{start}/counts"]), name
        assert np.array_equal(levels, golden[f"{name}/refine_from_L{start}/levels"]), name
        assert np.array_equal(xy.view(np.uint64), golden[f"{name}/refine_from_L{start}/xy"].view(np.uint64)), name


def test_refine_shuffled_foreign_points(api, oracle):
    img = synth.board_frame(800, 600, 10, seed=112)
    rng = np.random.default_rng(113)
    pts = oracle.find_corners(img, 2).astype(np.float64) / 1000.0
    rng.shuffle(pts)
    extra = np.stack([rng.uniform(-20, 820, 40), rng.uniform(-20, 620, 40)], axis=1)
    xy = np.concatenate([pts, extra, pts[:10] + 0.7])
    levels = np.full(len(xy), 2, dtype=np.int8)
    levels[::7] = 1
    for level in (1, 0):
        na, xa, la = api.refine_chessboard_corners(img, level, xy, levels)
        nb, xb, lb = oracle.refine_corners(img, level, xy, levels)
        assert na == nb and np.array_equal(la, lb)
        assert np.array_equal(xa.view(np.uint64), xb.view(np.uint64))
        xy, levels = xb, lb


def test_refine_4k_n14(api, oracle):
    img = synth.board_frame(3840, 2160, 14, seed=33)
    xa, la, ca = cases.refine_chain(api.find_chessboard_corners_int, api.refine_chessboard_corners, img, 3)
    xb, lb, cb = cases.refine_chain(oracle.find_corners, oracle.refine_corners, img, 3)
    assert np.array_equal(ca, cb) and np.array_equal(la, lb)
    assert np.array_equal(xa.view(np.uint64), xb.view(np.uint64))


# ---------------------------------------------------------------------------------------------
# batch API, device-resident frames, overflow path, kernel variants
# ---------------------------------------------------------------------------------------------
def _check_batch(xy, counts, frames, oracle, level=0):
    for i, img in enumerate(frames):
        want = oracle.find_corners(np.ascontiguousarray(img), level)
        assert counts[i] == len(want), (i, counts[i], len(want))
        assert np.array_equal(xy[i, :counts[i]], want), i


def test_batch_host_frames(api, oracle):
    frames = np.stack([synth.board_frame(480, 360, 10, seed=s) for s in range(9)] +
                      [synth.noise_frame(480, 360, seed=50), synth.checker_frame(480, 360, 8, seed=51)])
    det = api.Detector(max_frames=4, max_points=4096)      # forces 3 chunks
    for level in (0, 1):
        xy, counts = det.find_corners(frames, level)
        _check_batch(xy, counts, frames, oracle, level)
    det.close()


def test_batch_device_frames_torch(api, oracle):
    import torch
    frames = np.stack([synth.board_frame(640, 480, 10, seed=60 + s) for s in range(5)])
    t = torch.from_numpy(frames).cuda()
    det = api.Detector(max_frames=8)
    xy, counts = det.find_corners(t, 0, stream=torch.cuda.current_stream().cuda_stream)
    _check_batch(xy, counts, frames, oracle, 0)
    # pitched device frames (row pitch 768 > 640)
    tp = torch.zeros((5, 480, 768), dtype=torch.uint8, device="cuda")
    tp[:, :, :640] = t
    xy, counts = det.find_corners(tp[:, :, :640], 0)
    _check_batch(xy, counts, frames, oracle, 0)
    xy, counts = det.find_corners(tp[:, :, :640], 2)
    _check_batch(xy, counts, frames, oracle, 2)
    det.close()


def test_candidate_overflow_reruns_on_gpu(api, oracle):
    frames = np.stack([synth.noise_frame(400, 300, seed=70), synth.board_frame(400, 300, 10, seed=71),
                       synth.checker_frame(400, 300, 6, seed=72)])
    det = api.Detector(max_frames=3, candidate_capacity=256, max_points=4096)
    xy, counts = det.find_corners(frames, 0)
    cand = det.last_candidate_counts(3)
    assert (cand > 256).all(), cand          # every frame overflowed the tiny capacity ...
    _check_batch(xy, counts, frames, oracle, 0)   # ... and is still exact
    det.close()


def test_large_candidate_lists_use_global_scratch(api, oracle):
    # > 4096 candidates per frame: the clustering kernel leaves shared memory
    frames = np.stack([synth.noise_frame(640, 480, seed=80), synth.checker_frame(640, 480, 8, seed=81)])
    det = api.Detector(max_frames=2, candidate_capacity=1 << 17, max_points=8192)
    xy, counts = det.find_corners(frames, 0)
    cand = det.last_candidate_counts(2)
    assert (cand > 4096).all() and (cand < (1 << 17)).all(), cand
    _check_batch(xy, counts, frames, oracle, 0)
    det.close()


def test_kernel_variants_agree(api, oracle):
    # 0 = cascade (production), 1 = one thread per pixel, 2 = tiled: identical candidate sets and corners
    frames = np.stack([synth.board_frame(800, 608, 10, seed=90), synth.blurred_noise_frame(800, 608, seed=91)])
    dets = [api.Detector(max_frames=2, max_points=4096, kernel_variant=v) for v in (0, 1, 2)]
    res = [d.find_corners(frames, 0) for d in dets]
    cands = [d.last_candidate_counts(2) for d in dets]
    for v in (1, 2):
        assert np.array_equal(res[0][1], res[v][1]) and np.array_equal(cands[0], cands[v]), (v, cands)
        for i in range(2):
            assert np.array_equal(res[0][0][i, :res[0][1][i]], res[v][0][i, :res[v][1][i]])
    _check_batch(res[0][0], res[0][1], frames, oracle, 0)
    for d in dets:
        d.close()


def test_empty_batch_and_tiny_frames(api, oracle):
    det = api.Detector(max_frames=4)
    xy, counts = det.find_corners(np.zeros((0, 32, 32), np.uint8), 0)
    assert len(counts) == 0
    for (w, h) in ((1, 1), (14, 14), (15, 15), (16, 40), (40, 16)):
        frames = np.stack([synth.noise_frame(w, h, seed=s) for s in range(3)])
        xy, counts = det.find_corners(frames, 0)
        _check_batch(xy, counts, frames, oracle, 0)
    det.close()


# ---------------------------------------------------------------------------------------------
# loader paths of the tiled kernel, randomised sizes
# ---------------------------------------------------------------------------------------------
def test_device_frames_that_defeat_tma(api, oracle):
    """unaligned base address, odd width, pitch not a multiple of 16: the tiled kernel must take its
    cooperative loader and still match"""
    import torch
    det = api.Detector(max_frames=4, max_points=4096)
    base = np.stack([synth.board_frame(648, 488, 10, seed=120 + s) for s in range(3)])
    big = torch.zeros((3, 488, 700), dtype=torch.uint8, device="cuda")
    big[:, :, 1:649] = torch.from_numpy(base).cuda()
    for (x0, wv) in ((1, 640), (1, 641), (3, 645), (16, 632)):
        view = big[:, 4:484, x0:x0 + wv]                 # pitch 700, base offset 4*700 + x0
        frames = base[:, 4:484, x0 - 1:x0 - 1 + wv]
        xy, counts = det.find_corners(view, 0)
        _check_batch(xy, counts, frames, oracle, 0)
    det.close()


def test_random_sizes_and_levels(api, oracle):
    rng = np.random.default_rng(2024)
    det = api.Detector(max_frames=2, max_points=8192)
    for it in range(24):
        w, h = int(rng.integers(20, 700)), int(rng.integers(20, 500))
        kind = it % 4
        if kind == 0:
            img = synth.board_frame(w, h, 10, seed=300 + it) if min(w, h) > 120 else synth.noise_frame(w, h, seed=300 + it)
        elif kind == 1:
            img = synth.blurred_noise_frame(w, h, seed=300 + it, passes=1)
        elif kind == 2:
            img = synth.checker_frame(w, h, period=int(rng.integers(4, 14)), seed=300 + it, noise_sigma=float(rng.uniform(0, 6)))
        else:
            img = synth.blob_frame(w, h, seed=300 + it)
        level = int(rng.integers(0, 3))
        xy, counts = det.find_corners(img[None], level)
        _check_batch(xy, counts, img[None], oracle, level)
    det.close()


def test_segmented_single_frame_matches(api, oracle):
    """a single large frame is cut into row segments so the chip is filled: seams must not show"""
    img = synth.board_frame(2560, 1440, 14, seed=77)
    got, want = api.find_chessboard_corners_int(img, 0), oracle.find_corners(img, 0)
    assert np.array_equal(got, want) and len(want) == 196
    tex = synth.blurred_noise_frame(1600, 1200, seed=78, passes=1)
    assert np.array_equal(api.find_chessboard_corners_int(tex, 0), oracle.find_corners(tex, 0))


def test_refine_batch_matches_oracle(api, oracle):
    frames = np.stack([synth.board_frame(800, 600, 10, seed=130 + s) for s in range(5)] +
                      [synth.checker_frame(800, 600, 9, seed=140)])
    det = api.Detector(max_frames=4, max_points=8192)
    # points found at level 2, padded to a common count with dummies that must stay untouched
    found = [oracle.find_corners(f, 2).astype(np.float64) / 1000.0 for f in frames]
    npts = max(len(p) for p in found) + 3
    xy = np.full((len(frames), npts, 2), -100.0)
    lv = np.full((len(frames), npts), 7, dtype=np.int8)
    for i, p in enumerate(found):
        xy[i, :len(p)] = p
        lv[i, :len(p)] = 2
    for level in (1, 0):
        nref, xy2, lv2 = det.refine_corners(frames, level, xy, lv)
        for i, f in enumerate(frames):
            n0, x0, l0 = oracle.refine_corners(f, level, xy[i], lv[i])
            assert nref[i] == n0 and np.array_equal(lv2[i], l0), (level, i)
            assert np.array_equal(xy2[i].view(np.uint64), x0.view(np.uint64)), (level, i)
        xy, lv = xy2, lv2
    # tiny candidate capacity: every frame overflows and is re-run one by one
    det2 = api.Detector(max_frames=8, candidate_capacity=128)
    xy = np.stack([np.pad(p, ((0, npts - len(p)), (0, 0))) for p in found]); lv = np.full((len(frames), npts), 2, dtype=np.int8)
    nref, xy2, lv2 = det2.refine_corners(frames, 1, xy, lv)
    for i, f in enumerate(frames):
        n0, x0, l0 = oracle.refine_corners(f, 1, xy[i], lv[i])
        assert nref[i] == n0 and np.array_equal(lv2[i], l0) and np.array_equal(xy2[i].view(np.uint64), x0.view(np.uint64)), i
    det.close(); det2.close()


def test_8k_frame(api, oracle):
    # BASELINE.json configs[4] tops out at 7680x4320: ten strips of 768 pixels per frame
    frame = synth.board_frame(7680, 4320, 10, seed=123)
    got = api.find_chessboard_corners_int(frame, 0)
    want = oracle.find_corners(frame, 0)
    assert len(want) == 100 and np.array_equal(got, want)
    got2 = api.find_chessboard_corners_int(frame, 2)
    assert np.array_equal(got2, oracle.find_corners(frame, 2))


def test_cascade_adversarial_frames(api, oracle):
    # frames built to stress the cascade kernel's packed-byte tests: extreme contrasts (byte-lane wrap
    # guards), one-pixel stripes and checkers (everything reaches L2), and widths that select every
    # strip geometry (1, 2 and 3 warps per CTA, partial last strips, widths not multiples of 4)
    rng = np.random.default_rng(7)
    for (w, h) in ((250, 64), (301, 97), (530, 120), (799, 150), (1100, 90), (1537, 70)):
        imgs = [
            (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8),                                  # salt and pepper
            np.where((np.arange(w)[None, :] + np.arange(h)[:, None]) % 2 == 0, 255, 0).astype(np.uint8) + np.zeros((h, w), np.uint8),
            np.tile(np.where(np.arange(w) % 2 == 0, 255, 0).astype(np.uint8), (h, 1)),            # vertical stripes
            np.tile(np.where(np.arange(h) % 3 == 0, 255, 0).astype(np.uint8)[:, None], (1, w)),   # horizontal stripes
            np.clip(rng.normal(128, 60, (h, w)), 0, 255).astype(np.uint8),                        # wide noise
            synth.checker_frame(w, h, period=6, seed=int(w)),
        ]
        frames = np.stack([np.ascontiguousarray(i) for i in imgs])
        det = api.Detector(max_frames=len(frames), candidate_capacity=1 << 18, max_points=1 << 15)
        xy, counts = det.find_corners(frames, 0)
        _check_batch(xy, counts, frames, oracle, 0)
        det.close()


def test_dense_response_kernels_agree(api, oracle):
    # the dense API runs the tiled kernel's dense mode by default and the one-thread-per-pixel kernel with
    # kernel_variant=1; both must equal the oracle on every interior pixel, for aligned and odd shapes
    rng = np.random.default_rng(11)
    for (w, h) in ((640, 480), (611, 457), (250, 64), (1030, 40), (15, 15), (16, 33), (257, 15)):
        imgs = np.stack([synth.blurred_noise_frame(w, h, seed=int(w + h), passes=1),
                         (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)])
        for variant in (0, 1):
            det = api.Detector(max_frames=2, kernel_variant=variant)
            got = det.chess_response(imgs)
            for i in range(2):
                assert np.array_equal(got[i], oracle.chess_response_5(imgs[i], fill=0)), (w, h, variant, i)
            det.close()


def test_full_size_symmetry_properties(api):
    # Size-independent properties at BASELINE's full frame size, no oracle involved. The ChESS ring is
    # symmetric under a horizontal mirror and a 180-degree turn, and adding a constant to every pixel
    # changes neither sum, diff nor |mean - local_mean| (16*(s+3c)/3 truncates like 16*s/3 + 16c):
    #   dense response of the transformed frame == transformed dense response,
    #   number of pixels with response > 15 (the candidate list K1 emits) is unchanged.
    frame = synth.board_frame(3840, 2160, 10, seed=321)
    frame = np.minimum(frame, 235).astype(np.uint8)              # head-room for the +20 offset
    variants = {"mirror": np.ascontiguousarray(frame[:, ::-1]), "rot180": np.ascontiguousarray(frame[::-1, ::-1]),
                "offset": (frame + 20).astype(np.uint8)}
    det = api.Detector(max_frames=4, max_points=1024)
    base = det.chess_response(frame[None])[0]
    assert (base > 15).sum() > 500
    for name, img in variants.items():
        r = det.chess_response(img[None])[0]
        want = {"mirror": base[:, ::-1], "rot180": base[::-1, ::-1], "offset": base}[name]
        assert np.array_equal(r, want), name
    batch = np.stack([frame] + list(variants.values()))
    xy, counts = det.find_corners(batch, 0)
    cand = det.last_candidate_counts(4)
    assert cand[0] == (base > 15).sum() and (cand == cand[0]).all(), cand
    assert (counts == 100).all()
    # the offset frame has the same corners to the last digit; the mirrored frame has the mirrored corners, up
    # to the traversal-order dependence of the reference's clustering (membership is tested against the
    # running peak), a few hundredths of a pixel
    assert np.array_equal(xy[3, :100], xy[0, :100])
    a = np.sort(xy[0, :100, 0]); b = np.sort((3839 * 1000) - xy[1, :100, 0])
    assert np.abs(a - b).max() <= 100
    det.close()


def test_candidate_lists_equal_the_oracles_dense_response(api, oracle):
    """mrg_b200_chess_candidates_batch (the production ChESS kernel alone): the set {(x, y, r) : r > 15} of every
    frame equals the one read off the oracle's dense response, at levels 0 and 1, on boards, noise and checkers."""
    for (w, h) in ((640, 480), (1537, 70), (799, 150)):
        frames = np.stack([synth.board_frame(w, h, 6, seed=7), synth.noise_frame(w, h, seed=8),
                           synth.checker_frame(w, h, period=6, seed=9), synth.blurred_noise_frame(w, h, seed=10, passes=1)])
        det = api.Detector(max_frames=3, candidate_capacity=1 << 18)          # two chunks
        for level in (0, 1):
            counts, cand = det.chess_candidates(frames, level, cand_cap=1 << 18)
            for i, f in enumerate(frames):
                img = f if level == 0 else oracle.pyramid(f, level)
                r = oracle.chess_response_5(img, fill=0).astype(np.int64)
                hh, ww = img.shape
                ys, xs = np.nonzero(r[7:hh - 7, 7:ww - 7] > 15) if hh > 14 and ww > 14 else (np.zeros(0, int), np.zeros(0, int))
                ys, xs = ys + 7, xs + 7
                want = np.sort((ys.astype(np.uint64) << np.uint64(32)) | (xs.astype(np.uint64) << np.uint64(16)) | r[ys, xs].astype(np.uint64))
                assert counts[i] == len(want), (w, h, level, i)
                assert np.array_equal(np.sort(cand[i, :counts[i]]), want), (w, h, level, i)
        det.close()


def test_camera_like_4k_frames(api, oracle):
    """Not only clean synthetic boards: vignetting, a textured background and sensor noise of sigma 3-8 that reaches
    the detector unblurred, at 4K. Corner lists at levels 0-2, a refinement chain and the blob list against the oracle."""
    for seed, sigma in ((1, 3.0), (2, 5.0), (3, 8.0)):
        img = synth.camera_frame(3840, 2160, 10, seed=seed, noise_sigma=sigma)
        for level in (0, 1, 2):
            want = oracle.find_corners(img, level)
            got = api.find_chessboard_corners_int(img, level)
            assert np.array_equal(got, want), (seed, sigma, level, len(got), len(want))
        _, xy2 = oracle.find_corners(img, 2, want_double=True)
        if len(xy2):
            lv = np.full(len(xy2), 2, dtype=np.int8)
            n_o, xy_o, lv_o = oracle.refine_corners(img, 1, xy2, lv)
            n_g, xy_g, lv_g = api.refine_chessboard_corners(img, 1, xy2, lv)
            assert n_g == n_o and np.array_equal(xy_g, xy_o) and np.array_equal(lv_g, lv_o), (seed, sigma)
    dots = synth.circle_grid_frame(3840, 2160, 10, seed=4).astype(np.float32)
    dots += np.random.default_rng(5).normal(0.0, 5.0, size=dots.shape).astype(np.float32)
    dots = np.clip(np.rint(dots), 0, 255).astype(np.uint8)
    assert np.array_equal(api.find_blobs_int(dots), oracle.find_blobs(dots))
