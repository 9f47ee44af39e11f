"""CPU tests: the C restatement of HoughTransform::getLines (oracle/hough_oracle.c) against the
reference's own translation unit (oracle/_ref/libhough_ref.so) - accumulator bytes, the
order-dependent peak array and the lines, all bit for bit."""
import numpy as np
import pytest

import scan_synth
from hough_lib import PEAKS, RADIUS, THETA, HoughOracle, HoughRef


@pytest.fixture(scope="module")
def ho(built):
    return HoughOracle()


@pytest.fixture(scope="module")
def hr(built):
    if not HoughRef.available():
        pytest.skip("oracle/_ref/libhough_ref.so not built (needs /root/reference or a prebuilt binary)")
    return HoughRef()


def test_constants_and_tables(ho, hr):
    assert hr.constants() == (THETA, RADIUS, PEAKS)
    c, s = hr.tables()
    assert np.array_equal(c, ho.cos) and np.array_equal(s, ho.sin)


def test_get_lines_matches_the_reference(ho, hr):
    X, Y, R = scan_synth.make_scans(40, seed=3)
    n_lines = []
    for k in range(len(X)):
        la, pa, ga = ho.get_lines(X[k], Y[k], R[k], want_grid=True)
        lb, pb, gb = hr.get_lines(X[k], Y[k], R[k], want_grid=True)
        assert np.array_equal(ga, gb), "accumulator, scan %d" % k
        assert np.array_equal(pa, pb), "peak array, scan %d" % k
        assert la.shape == lb.shape and np.array_equal(la, lb), "lines, scan %d" % k
        assert np.array_equal(ho.lines_from_peaks(pa, ga[pa]), la)
        n_lines.append(len(la))
    assert min(n_lines) >= 1 and max(n_lines) >= 4


def test_edge_cases(ho, hr):
    """No return in range (nothing accumulated: no line), a single return, one straight wall, many
    returns on the same cell (counts above 127), returns exactly at MAX_DIST."""
    ang = np.deg2rad(np.arange(181) - 90.0)
    cases = []
    r = np.full(181, 8191, np.uint32)
    cases.append((r * np.cos(ang), r * np.sin(ang), r))                            # all out of range
    r = np.full(181, 8191, np.uint32); r[90] = 2500
    cases.append((r * np.cos(ang), r * np.sin(ang), r))                            # one return
    d = 3000.0 / np.maximum(np.cos(ang), 1e-3)                                     # wall x = 3 m
    r = np.where(d < 8000, np.rint(d), 8191).astype(np.uint32)
    cases.append((r * np.cos(ang), r * np.sin(ang), r))
    r = np.full(181, 4000, np.uint32)                                              # 181 returns on one point
    cases.append((np.full(181, 4000.0), np.zeros(181), r))
    r = np.full(181, 8000, np.uint32)                                              # exactly MAX_DIST (kept)
    cases.append((r * np.cos(ang), r * np.sin(ang), r))
    for k, (x, y, r) in enumerate(cases):
        la, pa, ga = ho.get_lines(x, y, r, want_grid=True)
        lb, pb, gb = hr.get_lines(x, y, r, want_grid=True)
        assert np.array_equal(ga, gb) and np.array_equal(pa, pb) and np.array_equal(la, lb), "case %d" % k
    assert len(ho.get_lines(*cases[0])[0]) == 0
    assert ho.get_lines(*cases[3], want_grid=True)[2].max() == 181


def test_feature_stages_match_the_reference(ho, hr):
    """fitLineSegments, extractCorners and getStructCompass (featuredetector.cpp:74-362) behind the
    Hough transform: segments, corner features and the compass value, bit for bit; the compass
    offset is carried from scan to scan as the detector object would."""
    X, Y, R = scan_synth.make_scans(60, seed=17, n_boxes=4)
    rng = np.random.default_rng(0)
    off_a = off_b = 100.0
    n_feats = 0
    for k in range(len(X)):
        phi = float(rng.uniform(-7, 7))
        fa, sa, ca, off_a, _ = ho.get_features(X[k], Y[k], R[k], phi, off_a)
        fb, sb, cb, off_b = hr.get_features(X[k], Y[k], R[k], phi, off_b)
        assert sa.shape == sb.shape and np.array_equal(sa, sb), "segments, scan %d" % k
        assert fa.shape == fb.shape and np.array_equal(fa, fb), "features, scan %d" % k
        assert ca == cb and off_a == off_b, "compass, scan %d" % k
        n_feats += len(fa)
    assert n_feats >= 30
