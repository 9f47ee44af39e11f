"""CPU tests of the host-side pieces of the Hough front-end in the product library (no GPU work):
the constructor tables and the peak grouping / line conversion (houghtransform.cpp:5-18, 58-236)."""
import numpy as np
import pytest

import scan_synth
from hough_lib import HoughOracle


@pytest.fixture(scope="module")
def ho(built):
    return HoughOracle()


def test_tables_match_the_reference_constructor(ekf, ho):
    c, s = ekf.hough_tables()
    assert np.array_equal(c, ho.cos) and np.array_equal(s, ho.sin)


def test_host_grouping_matches_the_oracle(ekf, ho):
    X, Y, R = scan_synth.make_scans(30, seed=2)
    for k in range(len(X)):
        lines, peaks, grid = ho.get_lines(X[k], Y[k], R[k], want_grid=True)
        assert np.array_equal(ekf.hough_lines_from_peaks(peaks, grid[peaks]), lines)
