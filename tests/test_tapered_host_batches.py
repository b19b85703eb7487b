"""Host-frame batches large enough for the tapered tail of mrg_b200_find_corners_batch (api.cu, enqueue_locked:
the last max_frames of a host batch go up as 256, 128, 128 ... frames so that the kernels of the last chunk
hide behind the copies): every frame's corner list must still equal the oracle's, whatever the launch sizes,
and forcing the taper on or off (MRG_B200_TAPER) must not change a single result."""
import os
import subprocess
import sys

import numpy as np
import pytest

from mrgingham_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames(n):
    base = [synth.board_frame(192, 160, 6, seed=300 + s) for s in range(5)] + [synth.noise_frame(192, 160, seed=310)]
    return np.stack([base[i % len(base)] for i in range(n)]), base


def test_host_batch_with_tapered_tail_matches_oracle():
    from mrgingham_b200 import api
    from oracle import pyoracle as po
    api._require_gpu()
    for n, max_frames in ((512, 512), (700, 512), (300, 1024)):        # 256,128,128 | 512,then 188 whole | 192,then 108 whole
        frames, base = _frames(n)
        want = [po.find_corners(b, 0) for b in base]
        det = api.Detector(max_frames=max_frames, max_points=1024)
        xy, counts = det.find_corners(frames, 0)
        for i in range(n):
            w = want[i % len(base)]
            assert counts[i] == len(w) and np.array_equal(xy[i, :counts[i]], w), (n, max_frames, i)
        det.close()


def test_taper_knob_does_not_change_results(tmp_path):
    # the knob is read once per process: run the same batch in two child processes
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
            "from mrgingham_b200 import api, synth\n"
            "base = [synth.board_frame(192, 160, 6, seed=300 + s) for s in range(5)] + [synth.noise_frame(192, 160, seed=310)]\n"
            "frames = np.stack([base[i %% 6] for i in range(512)])\n"
            "det = api.Detector(max_frames=512, max_points=1024)\n"
            "xy, counts = det.find_corners(frames, 0)\n"
            "np.savez(sys.argv[1], xy=np.stack([np.where(np.arange(1024)[:, None] < c, x, 0) for x, c in zip(xy, counts)]), counts=counts)\n" % ROOT)
    out = []
    for v in ("0", "1"):
        path = str(tmp_path / ("taper%s.npz" % v))
        env = dict(os.environ, MRG_B200_TAPER=v)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
        out.append(np.load(path))
    assert np.array_equal(out[0]["counts"], out[1]["counts"]) and np.array_equal(out[0]["xy"], out[1]["xy"])
    assert out[0]["counts"].max() > 0
