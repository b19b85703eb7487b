"""The oracle restatement (oracle/mrg_oracle.c) against the committed golden vectors, which were
generated from the reference itself (tests/golden/make_golden.py). CPU only."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402


def _names(golden):
    return sorted({k.split("/")[0] for k in golden.files})


def test_golden_has_all_cases(golden):
    assert set(_names(golden)) == set(cases.golden_images().keys())


def test_generator_reproduces_golden_images(golden):
    # the synthetic generator is deterministic: the bytes committed in the fixture are what
    # mrgingham_b200.synth produces today
    for name, img in cases.golden_images().items():
        assert np.array_equal(golden[f"{name}/image"], img), name


def test_response_matches_golden(golden, oracle):
    for name in _names(golden):
        img = golden[f"{name}/image"]
        assert np.array_equal(oracle.chess_response_5(img, fill=0), golden[f"{name}/response"]), name


@pytest.mark.parametrize("level", cases.LEVELS)
def test_corners_match_golden(golden, oracle, level):
    for name in _names(golden):
        img = golden[f"{name}/image"]
        got = oracle.find_corners(img, level)
        want = golden[f"{name}/corners_L{level}"]
        assert got.shape == want.shape and np.array_equal(got, want), (name, level)


@pytest.mark.parametrize("start", (1, 2, 3))
def test_refine_chain_matches_golden(golden, oracle, start):
    for name in _names(golden):
        img = golden[f"{name}/image"]
        xy, levels, counts = cases.refine_chain(oracle.find_corners, oracle.refine_corners, img, start)
        assert np.array_equal(counts, golden[f"{name}/refine_from_L{start}/counts"]), name
        assert np.array_equal(levels, golden[f"{name}/refine_from_L{start}/levels"]), name
        # bit-equal doubles
        assert np.array_equal(xy.view(np.uint64), golden[f"{name}/refine_from_L{start}/xy"].view(np.uint64)), name
