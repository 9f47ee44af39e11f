"""CPU tests: the C restatement of HoughTransform::getLines against the committed golden vectors
(tests/golden/hough/*.npz, produced from the reference's own translation unit by
tests/golden/make_golden_hough.py) - integer work, so everything is exact."""
import glob
import os

import numpy as np
import pytest

from hough_lib import RADIUS, THETA, HoughOracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hough", "*.npz")))


def golden_grid(g, k):
    grid = np.zeros(THETA * RADIUS, np.uint8)
    a, b = int(g["nz_off"][k]), int(g["nz_off"][k + 1])
    grid[g["nz_cell"][a:b]] = g["nz_val"][a:b]
    return grid


@pytest.fixture(scope="module")
def ho(built):
    return HoughOracle()


def test_golden_present():
    assert len(GOLDEN) >= 1


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(ho, path):
    g = np.load(path)
    assert np.array_equal(ho.cos, g["cos"]) and np.array_equal(ho.sin, g["sin"])
    for k in range(len(g["x"])):
        lines, peaks, grid = ho.get_lines(g["x"][k], g["y"][k], g["range"][k], want_grid=True)
        assert np.array_equal(grid, golden_grid(g, k)), "accumulator, scan %d" % k
        assert np.array_equal(peaks, g["peaks"][k]), "peak array, scan %d" % k
        n = int(g["n_lines"][k])
        assert len(lines) == n and np.array_equal(lines, g["lines"][k][:n]), "lines, scan %d" % k


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_feature_stages_match_golden(ho, path):
    g = np.load(path)
    for k in range(len(g["x"])):
        feats, segs, compass, off, _ = ho.get_features(g["x"][k], g["y"][k], g["range"][k], float(g["phi"][k]), float(g["off_in"][k]))
        n, m = int(g["n_feats"][k]), int(g["n_segs"][k])
        assert len(segs) == m and np.array_equal(segs, g["segs"][k][:m]), "segments, scan %d" % k
        assert len(feats) == n and np.array_equal(feats, g["feats"][k][:n]), "features, scan %d" % k
        assert compass == g["compass"][k] and off == g["off_out"][k], "compass, scan %d" % k
