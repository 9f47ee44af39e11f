"""GPU parity tests of the batched Hough line extractor (`ekf_hough_*`) against
the C restatement of HoughTransform::getLines (oracle/hough_oracle.c, itself pinned to the
reference's translation unit by tests/test_hough_oracle.py). Everything is integer or
integer-derived, so the bar is bit-exact: accumulator bytes, the order-dependent peak array, the
counts at the peaks, and the doubles of the lines."""
import numpy as np
import pytest

import scan_synth
from hough_lib import HoughOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ho(built):
    return HoughOracle()


def _check(ekf, ho, X, Y, R, want_grid):
    hb = ekf.HoughBatch(len(X))
    got = hb.get_lines(X, Y, R, max_lines=200, want_grid=want_grid)
    hb.close()
    for k in range(len(X)):
        lines, peaks, grid = ho.get_lines(X[k], Y[k], R[k], want_grid=True)
        if want_grid:
            assert np.array_equal(got["grid"][k], grid), "accumulator, scan %d" % k
        assert np.array_equal(got["peaks"][k], peaks), "peak array, scan %d" % k
        assert np.array_equal(got["values"][k], grid[peaks].astype(np.int32)), "counts at the peaks, scan %d" % k
        assert got["n_lines"][k] == len(lines), "line count, scan %d" % k
        assert np.array_equal(got["lines"][k], lines), "lines, scan %d" % k
    return got


def test_scans_match_the_oracle_bit_for_bit(ekf, ho):
    X, Y, R = scan_synth.make_scans(64, seed=11)
    got = _check(ekf, ho, X, Y, R, want_grid=True)
    assert got["n_lines"].min() >= 1 and got["n_lines"].max() >= 4


def test_cuda_matches_the_golden_vectors(ekf):
    """tests/golden/hough/*.npz were produced by the reference's own houghtransform.cpp."""
    from test_hough_golden import GOLDEN, golden_grid
    assert GOLDEN
    for path in GOLDEN:
        g = np.load(path)
        hb = ekf.HoughBatch(len(g["x"]))
        got = hb.get_lines(g["x"], g["y"], g["range"], max_lines=200, want_grid=True)
        hb.close()
        for k in range(len(g["x"])):
            n = int(g["n_lines"][k])
            assert np.array_equal(got["grid"][k], golden_grid(g, k)), "accumulator, scan %d" % k
            assert np.array_equal(got["peaks"][k], g["peaks"][k]), "peak array, scan %d" % k
            assert got["n_lines"][k] == n and np.array_equal(got["lines"][k], g["lines"][k][:n]), "lines, scan %d" % k
        hb = ekf.HoughBatch(len(g["x"]))
        f = hb.get_features(g["x"], g["y"], g["range"], cur_phi=g["phi"], offset=g["off_in"], max_feats=64,
                            want_segments=True, max_segs=64)
        hb.close()
        for k in range(len(g["x"])):
            n, m = int(g["n_feats"][k]), int(g["n_segs"][k])
            assert f["n_segs"][k] == m and np.array_equal(f["segments"][k, :m], g["segs"][k][:m]), "segments, scan %d" % k
            assert f["n_feats"][k] == n and np.array_equal(f["feats"][k, :n], g["feats"][k][:n]), "features, scan %d" % k
            assert f["compass"][k] == g["compass"][k] and f["offset"][k] == g["off_out"][k], "compass, scan %d" % k


def test_more_scans_than_sms(ekf, ho):
    """A persistent grid: 400 scans over 148 CTAs; every scan must still come out exact."""
    X, Y, R = scan_synth.make_scans(400, seed=5, n_boxes=5)
    _check(ekf, ho, X, Y, R, want_grid=False)


def test_pipelined_chunks_and_resident_path_agree(ekf, ho):
    """More scans than two waves of CTAs: ekf_hough_get_lines splits them into chunks over three
    streams; upload / run_resident / download is one launch. Both must equal the oracle."""
    base = scan_synth.make_scans(96, seed=9)
    reps = 20                                            # 1,920 scans: two chunks on a 148-SM part
    X, Y, R = (np.ascontiguousarray(np.tile(a, (reps, 1))) for a in base)
    hb = ekf.HoughBatch(len(X))
    got = hb.get_lines(X, Y, R, max_lines=40)
    hb.upload(X, Y, R)
    hb.run_resident(40)
    lines2, n2 = hb.download(40)
    hb.close()
    want = [ho.get_lines(base[0][k], base[1][k], base[2][k]) for k in range(96)]
    for s in range(len(X)):
        lines, peaks, _ = want[s % 96]
        assert np.array_equal(got["peaks"][s], peaks), "scan %d" % s
        assert got["n_lines"][s] == len(lines) == n2[s]
        assert np.array_equal(got["lines"][s], lines[:40]) and np.array_equal(lines2[s, :len(lines[:40])], lines[:40])


def test_edge_cases(ekf, ho):
    ang = np.deg2rad(np.arange(181) - 90.0)
    rows = []
    r = np.full(181, 8191, np.uint32)
    rows.append((r * np.cos(ang), r * np.sin(ang), r))                             # nothing in range
    r = np.full(181, 8191, np.uint32); r[90] = 2500
    rows.append((r * np.cos(ang), r * np.sin(ang), r))                             # one return
    d = 3000.0 / np.maximum(np.cos(ang), 1e-3)
    r = np.where(d < 8000, np.rint(d), 8191).astype(np.uint32)
    rows.append((r * np.cos(ang), r * np.sin(ang), r))                             # one wall
    r = np.full(181, 4000, np.uint32)
    rows.append((np.full(181, 4000.0), np.zeros(181), r))                          # 181 votes in one cell
    r = np.full(181, 8000, np.uint32)
    rows.append((r * np.cos(ang), r * np.sin(ang), r))                             # exactly MAX_DIST
    X = np.stack([a for a, _, _ in rows]); Y = np.stack([b for _, b, _ in rows]); R = np.stack([c for _, _, c in rows])
    got = _check(ekf, ho, X, Y, R, want_grid=True)
    assert got["n_lines"][0] == 0 and got["grid"][3].max() == 181
    # fewer readings per scan than an LMS-200 delivers, and the largest accepted count
    for P in (1, 37, 200):
        Xs, Ys, Rs = scan_synth.make_scans(3, seed=P)
        idx = np.arange(P) % 181
        _check(ekf, ho, Xs[:, idx], Ys[:, idx], Rs[:, idx], want_grid=True)


def test_feature_stages_match_the_oracle(ekf, ho):
    """fitLineSegments / extractCorners / getStructCompass on the GPU (ekf_hough_get_features) against the
    restatement that is pinned to the reference's featuredetector.cpp: segments, corner features and
    the compass value. (The device's sin/cos of a line angle, rounded to float, could in principle
    differ from glibc's at a float rounding boundary, about 1e-8 per value; none does here.)"""
    X, Y, R = scan_synth.make_scans(300, seed=21, n_boxes=4)
    rng = np.random.default_rng(3)
    phi = rng.uniform(-7, 7, len(X))
    off = np.where(rng.random(len(X)) < 0.5, 100.0, rng.uniform(-1.5, 0.0, len(X)))
    hb = ekf.HoughBatch(len(X))
    got = hb.get_features(X, Y, R, cur_phi=phi, offset=off, max_feats=32, want_segments=True, max_segs=64, max_lines=64)
    hb.close()
    total = 0
    for k in range(len(X)):
        feats, segs, compass, new_off, lines = ho.get_features(X[k], Y[k], R[k], float(phi[k]), float(off[k]))
        assert got["n_lines"][k] == len(lines) and np.array_equal(got["lines"][k, :len(lines)], lines), "lines, scan %d" % k
        assert got["n_segs"][k] == len(segs) and np.array_equal(got["segments"][k, :len(segs)], segs), "segments, scan %d" % k
        assert got["n_feats"][k] == len(feats) and np.array_equal(got["feats"][k, :len(feats)], feats), "features, scan %d" % k
        assert got["compass"][k] == compass and got["offset"][k] == new_off, "compass, scan %d" % k
        total += len(feats)
    assert total >= 150


def test_argument_validation(ekf):
    hb = ekf.HoughBatch(4)
    X, Y, R = scan_synth.make_scans(5, seed=1)
    with pytest.raises(ekf.EkfError) as e:
        hb.get_lines(X, Y, R)                      # more scans than the handle was created for
    assert e.value.code == ekf.ERR_BAD_ARG
    idx = np.arange(201) % 181
    with pytest.raises(ekf.EkfError):
        hb.get_lines(X[:2, idx], Y[:2, idx], R[:2, idx])
    hb.close()
