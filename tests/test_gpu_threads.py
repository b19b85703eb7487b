"""The drop-in entry points called concurrently, the way the reference CLI calls the library: `-j N` worker
threads, image i to worker i mod N, each calling the one-image functions (mrgingham-from-image.cc:50,374-379);
the reference library is re-entrant (SURVEY.md 8b "Threading"), so the replacement has to be thread-safe.
ctypes releases the GIL around every call, so these threads really do enter the C ABI at the same time. Every
result must equal the one the same call gives alone (and the oracle's)."""
import threading

import numpy as np
import pytest

from mrgingham_b200 import api, synth
from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]      # a deadlock must fail, not hang the run


def _run_threads(n_threads, work):
    errors, results = [], [None] * n_threads

    def runner(t):
        try:
            results[t] = work(t)
        except Exception as e:                                    # noqa: BLE001 - reported below
            errors.append((t, repr(e)))

    ths = [threading.Thread(target=runner, args=(t,)) for t in range(n_threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errors, errors
    return results


def test_one_image_entry_points_from_many_threads():
    api._require_gpu()
    sizes = [(640, 480), (800, 608), (1280, 720), (611, 457)]
    images = [synth.board_frame(w, h, 10, seed=20 + i) for i, (w, h) in enumerate(sizes * 2)]
    dots = [synth.circle_grid_frame(640, 480, 10, seed=40 + i) for i in range(4)]
    want_corners = [[po.find_corners(im, L) for L in (0, 1)] for im in images]
    want_resp = [po.chess_response_5(im, fill=0) for im in images]
    want_blobs = [po.find_blobs(d) for d in dots]
    want_board = [api.find_board(im) for im in images]                       # alone, before the threads start
    n_threads = 8

    def work(t):
        out = []
        for rep in range(3):
            for i in range(t, len(images), 2):                               # overlapping image sets between threads
                im = images[i]
                kind = (t + rep + i) % 4
                if kind == 0:
                    out.append(("corners0", i, api.find_chessboard_corners_int(im, 0)))
                elif kind == 1:
                    out.append(("corners1", i, api.find_chessboard_corners_int(im, 1)))
                elif kind == 2:
                    out.append(("resp", i, api.ChESS_response_5(im)))
                else:
                    out.append(("board", i, api.find_board(im)))
            out.append(("blobs", t % 4, api.find_blobs_int(dots[t % 4])))
        return out

    for res in _run_threads(n_threads, work):
        for kind, i, got in res:
            if kind == "corners0":
                assert np.array_equal(got, want_corners[i][0]), (kind, i)
            elif kind == "corners1":
                assert np.array_equal(got, want_corners[i][1]), (kind, i)
            elif kind == "resp":
                assert np.array_equal(got[7:-7, 7:-7], want_resp[i][7:-7, 7:-7]), (kind, i)
            elif kind == "board":
                w = want_board[i]
                assert (got is None) == (w is None) and (w is None or np.array_equal(got, w)), (kind, i)
            else:
                assert np.array_equal(got, want_blobs[i]), (kind, i)


def test_detectors_of_their_own_per_thread():
    """one batch detector per worker thread (own streams and scratch), batches of different shapes at once"""
    api._require_gpu()
    shapes = [(480, 640), (608, 800), (720, 1280), (600, 960)]
    batches = [np.stack([synth.board_frame(w, h, 10, seed=60 + 10 * t + k) for k in range(3)]) for t, (h, w) in enumerate(shapes)]
    want = [[po.find_corners(f, 0) for f in b] for b in batches]

    def work(t):
        det = api.Detector(max_frames=2, max_points=256)                      # two chunks per call
        out = []
        for _ in range(3):
            xy, counts = det.find_corners(batches[t], 0)
            out.append((xy.copy(), counts.copy()))
        det.close()
        return out

    for t, res in enumerate(_run_threads(len(shapes), work)):
        for xy, counts in res:
            for k, w in enumerate(want[t]):
                assert counts[k] == len(w) and np.array_equal(xy[k, :counts[k]], w), (t, k)


def test_one_image_calls_overlap_across_threads():
    """The reference CLI's -j N (mrgingham-from-image.cc:374-379): N threads, one image each at a time. The one-image
    entry points borrow a detector each from a per-device pool (own stream, own scratch, pinned staging), so eight
    callers must get through clearly more images per second than one."""
    import time
    api._require_gpu()
    images = [synth.board_frame(1920, 1080, 10, seed=70 + i) for i in range(8)]
    want = [po.find_corners(im, 0) for im in images]
    for im in images:                                   # warm the pool: one detector per future thread is created lazily
        api.find_chessboard_corners_int(im, 0)

    def rate(n_threads, seconds=1.5):
        done = [0] * n_threads
        stop = time.perf_counter() + seconds

        def work(t):
            while time.perf_counter() < stop:
                got = api.find_chessboard_corners_int(images[t], 0)
                assert np.array_equal(got, want[t])
                done[t] += 1
            return done[t]

        t0 = time.perf_counter()
        _run_threads(n_threads, work)
        return sum(done) / (time.perf_counter() - t0)

    rate(8, 0.5)                                        # all eight detectors exist after this
    r1 = rate(1)
    r8 = rate(8)
    print(f"one-image 1080p calls: {r1:.0f}/s with 1 thread, {r8:.0f}/s with 8 threads ({r8 / r1:.2f}x)")
    assert r8 >= 3.0 * r1, (r1, r8)
