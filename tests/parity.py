"""Parity metrics and fixtures shared by the tests (SURVEY.md 8c "Parity metric").

decisions + landmark indices: exact equality per (filter, step, measurement);
state:       max|dx| / max(|x_ref|_inf, 1e-300)        <= 1e-9
covariance:  |dP|_max / |P_ref|_max                    <= 1e-9   (norm-wise)
"""
import numpy as np

TOL = 1e-9   # BASELINE.json north_star: "within 1e-9 relative per step in FP64"


def rel_state(x, x_ref):
    x, x_ref = np.asarray(x), np.asarray(x_ref)
    return float(np.abs(x - x_ref).max() / max(np.abs(x_ref).max(), 1e-300))


def rel_cov(P, P_ref):
    P, P_ref = np.asarray(P), np.asarray(P_ref)
    return float(np.abs(P - P_ref).max() / max(np.abs(P_ref).max(), 1e-300))


def assert_trace_equal(got, want, what=""):
    """Decisions and indices exact; Mahalanobis distances to TOL."""
    assert np.array_equal(got["decision"], want["decision"]), what + " decisions differ"
    assert np.array_equal(got["index"], want["index"]), what + " landmark indices differ"
    m, mr = got["mahal"], want["mahal"]
    ok = np.abs(m - mr) <= TOL * np.maximum(1.0, np.abs(mr))
    assert ok.all(), what + " Mahalanobis distances differ: max abs %g" % np.abs(m - mr).max()


def assert_state_close(x, P, x_ref, P_ref, what="", tol=TOL):
    assert x.shape == x_ref.shape and P.shape == P_ref.shape, what + " dimension differs"
    ex, eP = rel_state(x, x_ref), rel_cov(P, P_ref)
    assert ex <= tol, "%s state rel err %g" % (what, ex)
    assert eP <= tol, "%s covariance rel err %g" % (what, eP)


def injected_state(world_xy, seed, pose=(0.0, 0.0, 0.0), lm_sigma=0.02, rank=8):
    """A plausible large-map state without running N 'New' updates:
    x from the world (+ small errors), P = D + U U^T, SPD and bit-symmetric."""
    rng = np.random.default_rng(seed)
    N = len(world_xy)
    n = 3 + 2 * N
    x = np.zeros(n)
    x[:3] = pose
    x[3:] = (np.asarray(world_xy) + rng.normal(0, lm_sigma, (N, 2))).reshape(-1)
    U = rng.normal(0, 0.01, (n, rank))
    D = rng.uniform(0.5, 1.5, n) * (lm_sigma ** 2)
    D[:3] = [1e-4, 1e-4, 1e-5]
    P = U @ U.T
    P[np.arange(n), np.arange(n)] += D
    B = 2048                                   # mirror the upper triangle in place (n can be 20,003)
    for j0 in range(0, n, B):
        j1 = min(n, j0 + B)
        P[j0:j1, :j0] = P[:j0, j0:j1].T
        blk = P[j0:j1, j0:j1]
        P[j0:j1, j0:j1] = np.triu(blk) + np.triu(blk, 1).T
    if n <= 5000:
        assert np.array_equal(P, P.T)
    return x, P
