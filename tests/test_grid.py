"""Grid finder and whole-board entry points (find_grid.cc:1216-1445, mrgingham.cc:10-140).

The comparison with the reference's own compiled find_grid.cc / mrgingham.cc (over a Boost.Polygon voronoi stand-in)
lives in tests/test_grid_vs_ref.py. This file checks the pieces that do not need the reference build:
  * the library's neighbour graph (its own exact Delaunay triangulation) against the oracle's, which decides
    Voronoi adjacency from the definition, including degenerate inputs (lattices, cocircular, collinear,
    repeated points);
  * the library's grid against the oracle's restatement of find_grid.cc on the same points, and against the
    ground-truth ordering of synthetic boards (rows from the top edge, left to right);
  * on the GPU: find_board / find_boards against the oracle pipeline (corner oracle -> grid oracle ->
    refinement oracle), which for the corner and refinement steps is pinned to the compiled reference.
"""
import numpy as np
import pytest

from mrgingham_b200 import api, synth
from oracle import grid_oracle as go
from oracle import pyoracle as po


def board_points(gridn, w, h, seed, rot=0.3, persp=0.12, noise=0.05, extras=0):
    """corner positions of a synthetic board seen under a mild homography: (shuffled PointInt list, ground-truth
    float positions row by row from the top-left)"""
    rng = np.random.default_rng(seed)
    side = 0.8 * min(w, h)
    th = rng.uniform(-rot, rot)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    q = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], float) * side / 2
    q = (q * (1 + rng.uniform(-persp, persp, (4, 1)))) @ R.T + np.array([w / 2, h / 2])
    src = np.array([[0, 0], [gridn - 1, 0], [gridn - 1, gridn - 1], [0, gridn - 1]], float)
    H = synth._homography(src, q)
    ij = np.array([[j, i] for i in range(gridn) for j in range(gridn)], float)
    ph = np.c_[ij, np.ones(len(ij))] @ H.T
    xy = ph[:, :2] / ph[:, 2:] + rng.normal(0, noise, (len(ij), 2))
    pts = (xy * 1000 + 0.5).astype(np.int64)
    allp = pts
    if extras:
        ex = np.c_[rng.uniform(0, w, extras * 4), rng.uniform(0, h, extras * 4)]
        lo, hi = xy.min(0) - side / gridn, xy.max(0) + side / gridn
        ex = ex[((ex < lo) | (ex > hi)).any(1)][:extras]
        allp = np.r_[pts, (ex * 1000 + 0.5).astype(np.int64)]
    return allp[rng.permutation(len(allp))].astype(np.int32), pts / 1000.0


def same(a, b):
    return (a is None and b is None) or (a is not None and b is not None and np.array_equal(a, b))


def test_neighbour_graph_matches_the_definition():
    rng = np.random.default_rng(0)
    cases = []
    for t in range(48):
        n = int(rng.integers(2, 60))
        kind = t % 6
        if kind == 0:
            pts = rng.integers(0, 1000000, (n, 2))
        elif kind == 1:
            pts = rng.integers(0, 6, (n, 2))                                   # repeated, collinear and cocircular points
        elif kind == 2:
            pts = np.array([[i * 1000, j * 1000] for i in range(5) for j in range(6)])           # a perfect lattice
        elif kind == 3:
            pts = np.array([[i * 300 + j * 400, -i * 400 + j * 300] for i in range(5) for j in range(5)])   # rotated lattice
        elif kind == 4:
            pts = np.c_[np.arange(n) * 7, np.arange(n) * 3]                      # every point on one line
        else:
            pts = rng.integers(0, 40, (n, 2)) * np.array([1, 1000])
        cases.append(pts[rng.permutation(len(pts))])
    ring = [(3, 4), (4, 3), (5, 0), (0, 5), (-3, 4), (-4, 3), (-5, 0), (0, -5), (3, -4), (4, -3), (-3, -4), (-4, -3)]
    cases.append(np.array(ring + [(0, 0)]) * 1000)                              # twelve cocircular points and their centre
    cases.append(np.array(ring) * 1000)
    cases.append(np.array([[5, 5]]))
    for pts in cases:
        g = go._Graph(pts)
        got = api.voronoi_neighbours(pts)
        for i in range(len(pts)):
            assert got[i] == g.ring.get(i, []), (pts.tolist(), i)


def test_grid_matches_oracle_and_ground_truth():
    for seed in range(24):
        gridn = (10, 14, 6, 4)[seed % 4]
        pts, truth = board_points(gridn, 1920, 1080, seed, extras=(seed % 3) * 5)
        got = api.find_grid_from_points(pts, gridn)
        assert got is not None and np.array_equal(got, truth), seed
        if gridn != 14 or seed < 8:                                             # the definition-based oracle is slow
            assert same(got, go.find_grid_from_points(pts, gridn)), seed


def test_grid_strong_perspective_and_rotation():
    # beyond the gentle cases: the walk's in-between neighbours (find_grid.cc:44-84) come into play
    n_found = 0
    for seed in range(12):
        pts, truth = board_points(8, 1280, 960, 100 + seed, rot=0.6, persp=0.35, noise=0.3)
        got = api.find_grid_from_points(pts, 8)
        want = go.find_grid_from_points(pts, 8)
        assert same(got, want), seed
        if got is not None:
            n_found += 1
            assert np.array_equal(np.sort(got.view("f8,f8"), axis=0), np.sort(truth.view("f8,f8"), axis=0))   # the same points
    assert n_found >= 8


def test_grid_failures_agree_with_oracle():
    rng = np.random.default_rng(3)
    pts, _ = board_points(10, 1920, 1080, 1)
    cases = [pts[:50],                                   # half a board
             np.delete(pts, 17, axis=0),                 # one corner missing
             np.r_[pts, pts[:5]],                        # repeated points are one site
             rng.integers(0, 1000000, (80, 2)).astype(np.int32),
             pts[:1], pts[:2], pts[:3],
             np.array([[i * 50000 + 100000, j * 50000 + 100000] for i in range(10) for j in range(10)], np.int32)]   # exact lattice
    two = np.r_[board_points(6, 900, 900, 5)[0], board_points(6, 900, 900, 6)[0] + np.array([1000000, 0], np.int32)]
    cases.append(two)                                    # two boards: more than one pair of cycles
    for k, c in enumerate(cases):
        gridn = 6 if k == len(cases) - 1 else 10
        assert same(api.find_grid_from_points(c, gridn), go.find_grid_from_points(c, gridn)), k
    assert api.find_grid_from_points(np.r_[pts, pts[:5]], 10) is not None
    assert api.find_grid_from_points(np.zeros((0, 2), np.int32), 10) is None
    # wrong gridn for the board
    assert api.find_grid_from_points(pts, 9) is None and go.find_grid_from_points(pts, 9) is None


def test_vnlog_records():
    # mrgingham-from-image.cc:174-187
    assert api.format_vnlog("a.png", None, 0) == "a.png - - -\n"
    got = api.format_vnlog("a.png", [[1.5, 2.25], [100.0, 7.0]], [0, 2])
    assert got == "a.png 1.500000 2.250000 0\na.png 100.000000 7.000000 2\n"
    assert api.format_vnlog("b", [[1, 2]], 3) == "b 1.000000 2.000000 3\n"
    assert api.VNLOG_LEGEND == "# filename x y level"


_grid_cache = {}


def oracle_board(img, gridn, level, refine=True):
    """mrgingham.cc:36-140 from the oracles: returns (level found, xy, levels) or (-1, None, None)"""
    for L in ([3, 2, 1, 0] if level < 0 else [level]):
        key = (img.tobytes(), gridn, L)
        if key not in _grid_cache:
            pts = po.find_corners(img, L)
            _grid_cache[key] = go.find_grid_from_points(pts, gridn) if len(pts) else None
        grid = _grid_cache[key]
        if grid is None:
            continue
        lv = np.full(len(grid), L, np.int8)
        if refine:
            for l in range(L - 1, -1, -1):
                n, grid, lv = po.refine_corners(img, l, grid, lv)
                if n <= 0:
                    break
        return L, grid, lv
    return -1, None, None


@pytest.mark.gpu
def test_find_board_matches_oracle_pipeline():
    api._require_gpu()
    for (w, h, gridn, seed) in ((1280, 960, 10, 1), (1920, 1080, 14, 2), (800, 608, 10, 3), (640, 480, 6, 4)):
        img = synth.board_frame(w, h, gridn, seed=seed)
        for level in (-1, 0, 1, 2):
            L, xy, lv = oracle_board(img, gridn, level)
            got = api.find_board(img, image_pyramid_level=level, gridn=gridn)
            if L < 0:
                assert got is None, (w, h, level)
                continue
            assert got is not None and np.array_equal(got, xy), (w, h, level)
            Lg, xyg, lvg = api.find_chessboard_from_image_array(img, gridn, level, refine=True)
            assert Lg == L and np.array_equal(xyg, xy) and np.array_equal(lvg, lv), (w, h, level)
            Lg, xyg, _ = api.find_chessboard_from_image_array(img, gridn, level, refine=False)
            L0, xy0, _ = oracle_board(img, gridn, level, refine=False)
            assert Lg == L0 and np.array_equal(xyg, xy0)
    # nothing to find
    assert api.find_board(synth.noise_frame(320, 240, seed=1)) is None
    assert api.find_board(np.full((200, 300), 128, np.uint8)) is None
    # blobs: only at level 0 (mrgingham_pywrap.c:258-262)
    dots = synth.circle_grid_frame(1280, 960, 10, seed=5)
    with pytest.raises(RuntimeError):
        api.find_board(dots, blobs=True)
    got = api.find_board(dots, image_pyramid_level=0, blobs=True)
    want = go.find_grid_from_points(po.find_blobs(dots), 10)
    assert want is not None and got is not None and np.array_equal(got, want)


@pytest.mark.gpu
def test_find_boards_batch():
    import torch
    api._require_gpu()
    frames = [synth.board_frame(1024, 768, 10, seed=s) for s in range(5)]
    frames[2] = synth.noise_frame(1024, 768, seed=9)                   # no board in this one
    frames[3] = synth.board_frame(1024, 768, 10, seed=3, noise_sigma=12.0)
    raw = np.stack(frames)
    det = api.Detector(max_frames=2)                                   # several chunks
    for level in (-1, 1, 0):
        want = [oracle_board(f, 10, level) for f in raw]
        for images in (raw, torch.from_numpy(raw).cuda()):
            found, xy, lv = det.find_boards(images, gridn=10, level=level)
            for i, (L, wxy, wlv) in enumerate(want):
                assert found[i] == L, (level, i)
                if L >= 0:
                    assert np.array_equal(xy[i], wxy) and np.array_equal(lv[i], wlv), (level, i)
    assert [w[0] for w in want][2] == -1
    # blobs over a batch
    dots = np.stack([synth.circle_grid_frame(800, 608, 10, seed=s) for s in (1, 2)])
    found, xy, _ = det.find_boards(dots, gridn=10, level=0, blobs=True)
    for i in range(2):
        want = go.find_grid_from_points(po.find_blobs(dots[i]), 10)
        assert (found[i] == 0) == (want is not None)
        if want is not None:
            assert np.array_equal(xy[i], want)
    det.close()


@pytest.mark.gpu
def test_find_boards_pipelined_equals_serial_and_oracle(monkeypatch):
    """Chunks of >= 16 frames with refinement take the pipelined path (corner passes over all frames of the chunk ahead of
    the grid searches, refinement on a second detector beside them): same boards, levels and refined points as the
    pass-by-pass path (MRG_B200_BOARDS_SERIAL=1) and as the oracle pipeline, on frames that find their grid at different
    levels, not at all, or only at level 0, from host and device memory, in chunks that end in a short one."""
    import torch
    api._require_gpu()
    w, h, gridn = 1280, 720, 10
    frames = []
    for s in range(40):
        kind = s % 8
        if kind == 5:
            frames.append(synth.noise_frame(w, h, seed=100 + s))                                  # no board
        elif kind == 3:
            frames.append(synth.board_frame(w, h, gridn, seed=s, noise_sigma=10.0))               # noisy: coarse levels do better
        elif kind == 6:
            frames.append(synth.board_frame(w, h, gridn, seed=s, noise_sigma=0.5, blur=False))    # sharp
        else:
            frames.append(synth.board_frame(w, h, gridn, seed=s))
    raw = np.stack(frames)
    det = api.Detector(max_frames=24, max_points=512)                  # chunks of 24 and 16
    for level in (-1, 2, 0):
        monkeypatch.setenv("MRG_B200_BOARDS_SERIAL", "1")
        f0, xy0, lv0 = det.find_boards(raw, gridn=gridn, level=level)
        monkeypatch.delenv("MRG_B200_BOARDS_SERIAL")
        for images in (raw, torch.from_numpy(raw).cuda()):
            f1, xy1, lv1 = det.find_boards(images, gridn=gridn, level=level)
            assert np.array_equal(f0, f1), level
            ok = f0 >= 0
            assert np.array_equal(xy0[ok], xy1[ok]) and np.array_equal(lv0[ok], lv1[ok]), level
        if level == -1:
            assert len(set(int(v) for v in f0)) >= 3, sorted(set(int(v) for v in f0))   # several outcomes in one chunk
            for i in (0, 3, 5, 6, 27):
                L, wxy, wlv = oracle_board(raw[i], gridn, level)
                assert f1[i] == L, i
                if L >= 0:
                    assert np.array_equal(xy1[i], wxy) and np.array_equal(lv1[i], wlv), i
    det.close()
    # a detector whose configured max_points is smaller than a board's corner count: the corner passes of the pipeline
    # grow their own capacity and run again (the detector's configuration is left alone)
    small = api.Detector(max_frames=16, max_points=32)
    f2, xy2, lv2 = small.find_boards(raw[:16], gridn=gridn, level=-1)
    monkeypatch.setenv("MRG_B200_BOARDS_SERIAL", "1")
    f3, xy3, lv3 = small.find_boards(raw[:16], gridn=gridn, level=-1)
    monkeypatch.delenv("MRG_B200_BOARDS_SERIAL")
    ok = f3 >= 0
    assert ok.sum() >= 10 and np.array_equal(f2, f3) and np.array_equal(xy2[ok], xy3[ok]) and np.array_equal(lv2[ok], lv3[ok])
    assert small.max_points == 32
    small.close()
