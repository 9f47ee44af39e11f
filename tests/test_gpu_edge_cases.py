"""GPU parity on hand-built adversarial inputs, through EVERY engine
that implements doUpdate: the three fused batch kernels, the per-call kernels of both regimes, the
fused large regime and the sharded map. The synthetic world never produces these situations
(0 cond-skips, no ties, no NaN), so they are constructed here:

  * the condition gate `cond >= 80` (Update.cpp:127-131) swept across its threshold in steps down
    to 1e-13 relative - the CUDA gating takes a shortcut away from the threshold and the exact
    singular-value path inside a 1e-9 band, the decision must flip exactly where the oracle's does;
  * exact Mahalanobis ties (Update.cpp:140, strict '>': the lowest index wins);
  * the Ignore band Gamma_min <= d^2 <= Gamma_max (Update.cpp:152,181,191);
  * every landmark skipped by the gate -> New although a landmark sits under the measurement;
  * NaN / inf measurements (NaN is never selected, NaN cond is not skipped).
"""
import numpy as np
import pytest

from parity import TOL

pytestmark = pytest.mark.gpu

ENGINES = ["fused-smem", "fused-tile", "fused-stile", "fused-dtile", "fused-large", "percall-batch", "percall-large", "sharded3"]


def _records(steps, M):
    L = 8 + 6 * M
    rec = np.zeros((1, len(steps), L))
    for t, (vel, rot, dt, meas) in enumerate(steps):
        rec[0, t, 0:3] = (vel, rot, dt)
        rec[0, t, 5] = len(meas)
        for m, (z, R) in enumerate(meas):
            rec[0, t, 8 + 6 * m:10 + 6 * m] = z
            rec[0, t, 10 + 6 * m:14 + 6 * m] = R
    return rec


def _run_engine(ekf, engine, x0, P0, cap, steps, M):
    """-> decision [T][M], index [T][M], mahal [T][M], x, P"""
    T = len(steps)
    if engine.startswith("fused") or engine == "sharded3":
        rec = _records(steps, M)
        if engine == "sharded3":
            obj = ekf.ShardedMap([0, 0, 0], cap)
            obj.set_state(x0, P0)
            out = obj.run(rec, M, trace=True, allow_capacity=True)
            x, P = obj.get_state()
        else:
            kern = {"fused-smem": 1, "fused-tile": 2, "fused-stile": 3, "fused-dtile": 4, "fused-large": 0}[engine]
            obj = ekf.FilterBatch(1, cap, regime=2 if engine == "fused-large" else 1, batch_kernel=kern)
            obj.set_state(0, x0, P0)
            out = obj.run(rec, M, trace=True, allow_capacity=True)
            x, P = obj.get_state(0)
        obj.close()
        return out["decision"][0], out["index"][0], out["mahal"][0], x, P
    fb = ekf.FilterBatch(1, cap, regime=1 if engine == "percall-batch" else 2)
    fb.set_state(0, x0, P0)
    dec = np.full((T, M), -1, np.int32)
    idx = np.full((T, M), -1, np.int32)
    mah = np.zeros((T, M))
    for t, (vel, rot, dt, meas) in enumerate(steps):
        fb.propagate(vel, rot, dt)
        for m, (z, R) in enumerate(meas):
            try:
                d, i, v = fb.update(np.asarray(z)[None], np.asarray(R)[None])
                dec[t, m], idx[t, m], mah[t, m] = d[0, 0], i[0, 0], v[0, 0]
            except ekf.EkfError as e:
                assert e.code == ekf.ERR_CAPACITY
                dec[t, m], idx[t, m] = ekf.DECISION_DROPPED, -1
    x, P = fb.get_state(0)
    fb.close()
    return dec, idx, mah, x, P


def _run_oracle(oracle, x0, P0, cap, steps, M):
    T = len(steps)
    of = oracle.new_filter(cap).set_state(x0, P0)
    dec = np.full((T, M), -1, np.int32)
    idx = np.full((T, M), -1, np.int32)
    mah = np.zeros((T, M))
    for t, (vel, rot, dt, meas) in enumerate(steps):
        of.propagate(vel, rot, dt)
        for m, (z, R) in enumerate(meas):
            n_before = of.n
            try:
                tr = of.update(np.asarray(z, float), np.asarray(R, float))
            except OverflowError:
                dec[t, m], idx[t, m] = 3, -1
                continue
            dec[t, m] = tr.decision
            idx[t, m] = n_before if tr.decision == 0 else tr.opt_i
            mah[t, m] = tr.mahal
    x, P = of.get_state()
    return dec, idx, mah, x, P


def _check(got, want, what, compare_mahal_of_dropped=False):
    gd, gi, gm, gx, gP = got
    wd, wi, wm, wx, wP = want
    assert np.array_equal(gd, wd), "%s: decisions %s vs oracle %s" % (what, gd.ravel(), wd.ravel())
    assert np.array_equal(gi, wi), "%s: indices %s vs oracle %s" % (what, gi.ravel(), wi.ravel())
    live = wd != 3
    both_nan = np.isnan(gm) & np.isnan(wm)
    ok = both_nan | (np.abs(gm - wm) <= TOL * np.maximum(1.0, np.abs(wm)))
    assert ok[live].all(), "%s: Mahalanobis %s vs %s" % (what, gm.ravel(), wm.ravel())
    assert gx.shape == wx.shape and gP.shape == wP.shape, what + ": dimension"
    assert np.array_equal(np.isnan(gx), np.isnan(wx)) and np.array_equal(np.isnan(gP), np.isnan(wP)), what + ": NaN pattern"
    fx, fP = ~np.isnan(wx), ~np.isnan(wP)
    if fx.any():
        assert np.abs(gx[fx] - wx[fx]).max() <= TOL * max(np.abs(wx[fx]).max(), 1e-300), what + ": state"
    if fP.any():
        assert np.abs(gP[fP] - wP[fP]).max() <= TOL * max(np.abs(wP[fP]).max(), 1e-300), what + ": covariance"


def _map4(p_diag=1e-14, pose=(0.0, 0.0, 0.0)):
    lms = [(2.0, 0.0), (0.0, 2.0), (-2.0, 0.0), (0.0, -2.0)]
    x = np.array(list(pose) + [c for lm in lms for c in lm])
    P = np.eye(len(x)) * p_diag
    return x, P


@pytest.mark.parametrize("engine", ENGINES)
def test_condition_gate_threshold_sweep(ekf, oracle, engine):
    """S ~= R (P is 1e-14): cond(S) = 80 (1 + delta). The decision flips between Old (not skipped)
    and New (every landmark skipped) exactly where the oracle's `cond >= 80` does."""
    deltas = [0.0]
    for e in (1e-13, 1e-12, 1e-11, 1e-10, 1.5e-10, 2e-10, 2.5e-10, 3e-10, 4e-10, 5e-10, 7e-10, 1e-9, 3e-9, 1e-8, 1e-6, 1e-3):
        deltas += [e, -e]
    _sweep(ekf, oracle, engine, _map4(), deltas)


@pytest.mark.parametrize("engine", ENGINES)
def test_condition_gate_threshold_sweep_at_ulp_scale(ekf, oracle, engine):
    """P = 0 exactly, so S = R bit for bit and cond(S) sits within a few ulp of 80."""
    deltas = [0.0]
    for e in (2.3e-16, 4.5e-16, 9e-16, 2e-15, 1e-14, 1e-13):
        deltas += [e, -e]
    _sweep(ekf, oracle, engine, _map4(p_diag=0.0), deltas)


def _sweep(ekf, oracle, engine, state, deltas):
    x0, P0 = state
    flips = set()
    for k, delta in enumerate(deltas):
        r1 = 1e-4
        r0 = r1 * 80.0 * (1.0 + delta)
        # alternate which axis is the large one and add an off-diagonal rotation on some cases
        if k % 3 == 0:
            R = [r0, 0.0, 0.0, r1]
        elif k % 3 == 1:
            R = [r1, 0.0, 0.0, r0]
        else:
            c, s = np.cos(0.3), np.sin(0.3)
            Rm = np.array([[c, -s], [s, c]]) @ np.diag([r0, r1]) @ np.array([[c, s], [-s, c]])
            off = 0.5 * (Rm[0, 1] + Rm[1, 0])
            R = [Rm[0, 0], off, off, Rm[1, 1]]
        steps = [(0.0, 0.0, 0.1, [((2.0, 0.0), R)])]
        want = _run_oracle(oracle, x0, P0, 6, steps, 1)
        got = _run_engine(ekf, engine, x0, P0, 6, steps, 1)
        _check(got, want, "%s delta=%g" % (engine, delta))
        flips.add(int(want[0][0, 0]))
    assert flips == {0, 1}, "the sweep must cross the gate (saw decisions %s)" % flips


@pytest.mark.parametrize("engine", ENGINES)
def test_exact_ties_pick_the_lowest_index(ekf, oracle, engine):
    """Landmarks 2 and 4 are bitwise copies of each other: equal d^2, Opt_i must be the lower one."""
    x0, P0 = _map4(p_diag=1e-4)
    x0[5:7] = (1.0, 1.0)
    x0[9:11] = (1.0, 1.0)
    R = [1e-2, 0.0, 0.0, 1e-2]
    steps = [(0.0, 0.0, 0.1, [((1.0, 1.02), R)]), (0.0, 0.0, 0.1, [((1.0, 0.99), R)])]
    want = _run_oracle(oracle, x0, P0, 6, steps, 1)
    assert want[0][0, 0] == 1 and want[1][0, 0] == 5
    got = _run_engine(ekf, engine, x0, P0, 6, steps, 1)
    _check(got, want, engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_ignore_band_and_far_measurement(ekf, oracle, engine):
    """sigma ~ 0.1: an offset of 0.45 gives 10 < d^2 < 50 (Ignore, state untouched), 0.9 gives New."""
    x0, P0 = _map4(p_diag=1e-4)
    R = [1e-2, 0.0, 0.0, 1e-2]
    steps = [(0.0, 0.0, 0.1, [((2.45, 0.0), R), ((2.9, 0.0), R)]),
             (300.0, 5.0, 0.1, [((0.0, 2.0), R), ((2.03, 0.02), R)])]
    want = _run_oracle(oracle, x0, P0, 6, steps, 2)
    assert list(want[0][0]) == [2, 0]
    got = _run_engine(ekf, engine, x0, P0, 6, steps, 2)
    _check(got, want, engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_every_landmark_skipped_by_the_gate_gives_new(ekf, oracle, engine):
    """A corner 0.22 m away: R from slam.cpp:158-167 has cond 25/d^2 >> 80, every S is skipped, and
    the measurement starts a new landmark although one sits exactly there."""
    x0, P0 = _map4(p_diag=1e-8)
    x0[3:5] = (0.2, 0.1)
    z, R = oracle.measurement_from_feature(200.0, 100.0)
    steps = [(0.0, 0.0, 0.1, [(z, R)])]
    want = _run_oracle(oracle, x0, P0, 6, steps, 1)
    assert want[0][0, 0] == 0 and want[1][0, 0] == 11
    got = _run_engine(ekf, engine, x0, P0, 6, steps, 1)
    _check(got, want, engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_nan_and_inf_measurements(ekf, oracle, engine):
    """A NaN measurement is never associated (strict '>' on NaN is false) and is appended as a New
    landmark by the reference; later finite measurements still associate with the finite landmarks
    the NaN did not touch. Parity is on decisions, indices and the NaN pattern of the state."""
    x0, P0 = _map4(p_diag=1e-4)
    R = [1e-2, 0.0, 0.0, 1e-2]
    for bad in ((np.nan, 1.0), (np.inf, 0.0)):
        steps = [(0.0, 0.0, 0.1, [(bad, R)]), (0.0, 0.0, 0.1, [((2.0, 0.01), R)])]
        want = _run_oracle(oracle, x0, P0, 7, steps, 1)
        got = _run_engine(ekf, engine, x0, P0, 7, steps, 1)
        _check(got, want, "%s z=%s" % (engine, bad))


@pytest.mark.parametrize("engine", ENGINES)
def test_capacity_drop_leaves_state_untouched(ekf, oracle, engine):
    """A New beyond max_landmarks is reported as DROPPED and changes nothing; the next measurement
    is processed normally."""
    x0, P0 = _map4(p_diag=1e-4)
    R = [1e-2, 0.0, 0.0, 1e-2]
    steps = [(0.0, 0.0, 0.1, [((5.0, 5.0), R)]), (100.0, 2.0, 0.1, [((2.0, 0.01), R)])]
    want = _run_oracle(oracle, x0, P0, 4, steps, 1)
    assert want[0][0, 0] == 3 and want[0][1, 0] == 1
    got = _run_engine(ekf, engine, x0, P0, 4, steps, 1)
    _check(got, want, engine)
