"""Contracts of the C ABI that are not about numbers: output capacities are never changed behind the caller's back
(find_boards used to grow the detector's max_points, after which find_corners overflowed the caller's
xy_out), the calling thread's current CUDA device is left alone, failures are not reported as "nothing found"."""
import ctypes

import numpy as np
import pytest

from mrgingham_b200 import api, synth
from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_find_boards_does_not_grow_the_callers_output_capacity():
    api._require_gpu()
    # a dense checker has thousands of corners: far more than the detector's max_points of 64
    clutter = synth.checker_frame(640, 480, 8, seed=1)
    board = synth.board_frame(640, 480, 10, seed=2)
    nclutter = len(po.find_corners(clutter, 0))
    assert nclutter > 1024
    det = api.Detector(max_frames=4, max_points=64)
    frames = np.stack([clutter, board])
    found, xy, lv = det.find_boards(frames, gridn=10, level=0, refine=False)
    assert found[0] == -1 and found[1] == 0
    # the same detector, the configured capacity: xy_out is [n][64][2] and must not be overrun
    guard = np.full((2, 64 + 64, 2), -7, dtype=np.int32)          # room behind each frame's 64 points that must stay untouched
    xy2, counts = det.find_corners(frames, 0)
    assert xy2.shape == (2, 64, 2)
    assert counts[0] == nclutter and counts[1] == 100
    assert np.array_equal(xy2[0], po.find_corners(clutter, 0)[:64])
    assert np.array_equal(xy2[1], po.find_corners(board, 0)[:64])
    # through the raw ABI with a guarded buffer
    flat = np.full(2 * 64 * 2 + 1024, -7, dtype=np.int32)
    cnt = np.zeros(2, dtype=np.int32)
    rc = api.lib().mrg_b200_find_corners_batch(det._h, frames.ctypes.data, 0, 2, 480, 640, 640, 640 * 480, 0,
                                               flat.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                               cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None)
    assert rc == 0 and np.all(flat[2 * 64 * 2:] == -7) and np.all(guard == -7)
    bxy, bcounts = det.find_blobs(np.stack([synth.circle_grid_frame(640, 480, 10, seed=3)]))
    assert bxy.shape == (1, 64, 2) and bcounts[0] == 100
    det.close()


def test_current_device_is_left_alone():
    import torch
    api._require_gpu()
    before = torch.cuda.current_device()
    img = synth.board_frame(640, 480, 10, seed=5)
    api.find_chessboard_corners_int(img, 0)
    api.ChESS_response_5(img)
    api.find_board(img)
    det = api.Detector(max_frames=2, device=0)
    det.find_corners(np.stack([img, img]), 0)
    det.close()
    assert torch.cuda.current_device() == before
    if torch.cuda.device_count() > 1:
        # a thread working on cuda:1 stays there, and its one-image calls run there
        torch.cuda.set_device(1)
        got = api.find_chessboard_corners_int(img, 0)
        assert np.array_equal(got, po.find_corners(img, 0))
        assert torch.cuda.current_device() == 1
        det = api.Detector(max_frames=2, device=0)
        det.find_corners(np.stack([img, img]), 0)
        det.close()
        assert torch.cuda.current_device() == 1
        torch.cuda.set_device(before)


def test_overlapped_passes_on_two_detectors():
    """bench.py's loop: pass p+1 is enqueued (other detector, same stream, same frames) before pass p is collected"""
    import torch
    api._require_gpu()
    base = [synth.board_frame(800, 608, 10, seed=80 + s) for s in range(6)]
    want = [po.find_corners(b, 0) for b in base]
    frames = torch.from_numpy(np.stack(base)).cuda()
    dets = [api.Detector(max_frames=2, max_points=128) for _ in range(2)]     # three chunks per pass
    stream = torch.cuda.current_stream().cuda_stream
    inflight, results = [], []
    for p in range(6):
        d = dets[p % 2]
        d.enqueue(frames, 0, stream=stream)
        inflight.append(d)
        if len(inflight) == 2:
            results.append(inflight.pop(0).collect())
    results += [d.collect() for d in inflight]
    assert len(results) == 6
    for xy, counts in results:
        for k, w in enumerate(want):
            assert counts[k] == len(w) and np.array_equal(xy[k, :counts[k]], w)
    for d in dets:
        d.close()
