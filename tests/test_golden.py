"""CPU tests (not gpu): the oracle against the committed golden vectors (tests/golden/*.npz,
generated from the reference build by tests/golden/make_golden.py). Decisions and indices must be
exact; floating-point values are compared to 1e-12 because the golden run and this run may use
different glibc sin/cos variants (ifunc dispatch by CPU)."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def check_against_golden(g, got, get_state, tol):
    assert np.array_equal(got["decision"], g["decision"])
    assert np.array_equal(got["index"], g["index"])
    assert np.array_equal(got["final_nlm"], g["final_nlm"])
    assert np.abs(got["mahal"] - g["mahal"]).max() <= tol * max(1.0, np.abs(g["mahal"][g["mahal"] < 1e11]).max())
    assert np.abs(got["pose_trace"] - g["pose_trace"]).max() <= tol * np.abs(g["pose_trace"]).max()
    for f in range(len(g["final_nlm"])):
        n = 3 + 2 * int(g["final_nlm"][f])
        x, P = get_state(f, n)
        xr, Pr = g["final_x"][f, :n], g["final_P"][f, :n, :n].T
        assert np.abs(x - xr).max() <= tol * np.abs(xr).max()
        assert np.abs(P - Pr).max() <= tol * np.abs(Pr).max()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(oracle, path):
    g = np.load(path)
    rec = np.ascontiguousarray(g["records"])
    got = oracle.run_batch(rec, int(g["max_meas"]), int(g["cap"]), pose_trace=True, final_state=True)
    assert not got["bad"]
    check_against_golden(g, got, lambda f, n: (got["final_x"][f, :n], got["final_P"][f, :n, :n].T), 1e-12)


def test_golden_files_present():
    assert len(GOLDEN) >= 2
