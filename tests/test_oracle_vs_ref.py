"""CPU tests (not gpu): the C restatement (oracle/ekf_oracle.c) must be BIT-IDENTICAL to the
reference's own translation units compiled over the stand-in Eigen (oracle/_ref). This is what
pins the oracle (oracle/ekf_oracle.c header, "PARITY STATUS")."""
import numpy as np
import pytest


def _records(ekf, N, F, T, laps=1, **kw):
    syn = ekf.Synth(N, steps_per_lap=T, **kw)
    lap = syn.generate(F, T)
    return np.ascontiguousarray(np.concatenate([lap] * laps, axis=1))


@pytest.mark.parametrize("N,F,T,laps,M,compass", [(20, 2, 1000, 1, 1, 0), (50, 1, 1000, 2, 1, 0),
                                                  (12, 3, 300, 2, 3, 4), (8, 2, 120, 3, 2, 1)])
def test_sequences_bit_identical(ekf, oracle, ref, N, F, T, laps, M, compass):
    rec = _records(ekf, N, F, T, laps, max_meas=M, compass_every=compass)
    a = oracle.run_batch(rec, M, N + 6, pose_trace=True, final_state=True)
    b = ref.run_batch(rec, M, N + 6, pose_trace=True, final_state=True)
    assert not a["bad"] and not b["bad"]      # b["bad"] would mean a harness cross-check failed
    for k in ("decision", "index", "mahal", "pose_trace", "final_pose", "final_nlm", "final_x", "final_P"):
        assert np.array_equal(a[k], b[k]), k
    assert (a["final_nlm"] == N).all()
    d = a["decision"]
    assert (d == 0).sum() == N * F, "every landmark is initialised exactly once"


def test_single_calls_bit_identical_with_trace(ekf, oracle, ref):
    """Per-call surface, including the observed gating values (cond skips, margins, Opt_i)."""
    N, T = 10, 200
    rec = _records(ekf, N, 1, T, 2, max_meas=2, compass_every=3)[0]
    fo, fr = oracle.new_filter(N + 2), ref.new_filter()
    min_margin = 1e300
    for r in rec:
        fo.propagate(r[0], r[1], r[2])
        fr.propagate(r[0], r[1], r[2])
        if r[6]:
            fo.update_compass(r[3], r[4])
            fr.update_compass(r[3], r[4])
        for m in range(int(r[5])):
            z, R = r[8 + 6 * m:10 + 6 * m], r[10 + 6 * m:14 + 6 * m]
            to, tr = fo.update(z, R), fr.update(z, R)
            for fld in ("decision", "opt_i", "mahal", "n_cond_skipped", "k_col", "margin_gmin", "margin_gmax",
                        "margin_cond"):
                assert getattr(to, fld) == getattr(tr, fld), fld
            if to.opt_i:
                min_margin = min(min_margin, to.margin_gmin, to.margin_gmax)
        xo, Po = fo.get_state()
        xr, Pr = fr.get_state()
        assert np.array_equal(xo, xr) and np.array_equal(Po, Pr)
        assert np.array_equal(Po, Po.T), "P is bit-symmetric at call boundaries"
    assert min_margin > 1e-6, "no knife-edge threshold decisions in the synthetic world"


def _chunk(r):
    nz = int(r[5])
    zs = np.array([r[8 + 6 * m:10 + 6 * m] for m in range(nz)])                   # [n_z][2]
    Rs = np.array([r[10 + 6 * m:14 + 6 * m] for m in range(nz)])                  # [n_z][4] column-major 2x2
    return nz, zs, Rs


def test_private_update_with_nz_gt_1(ekf, oracle, ref):
    """One doUpdate(z_chunk) with n_z > 1 (Update.cpp:80-195): measurements are applied sequentially,
    but the gating bound n_lm is read once at call entry (Update.cpp:26). The restatement's
    update_chunk follows the reference's private Update bit for bit from an empty map onwards
    (chunks that add several landmarks included)."""
    N, T = 8, 40
    rec = _records(ekf, N, 1, T, 2, max_meas=3)[0]
    fo = oracle.new_filter(N + 8)
    seen_new_in_chunk = 0
    for r in rec[:T]:
        fo.propagate(r[0], r[1], r[2])
        nz, zs, Rs = _chunk(r)
        if nz == 0:
            continue
        x0, P0 = fo.get_state()
        x1, P1 = ref.call_update(x0, P0, zs.T, np.concatenate([R.reshape(2, 2).T for R in Rs], axis=1))
        trs = fo.update_chunk(zs, Rs)
        seen_new_in_chunk += sum(t.decision == 0 for t in trs) > 1
        x2, P2 = fo.get_state()
        assert np.array_equal(x1, x2) and np.array_equal(P1, P2)
    assert seen_new_in_chunk > 0 and fo.num_landmarks == N


def test_chunk_gating_bound_is_frozen_at_call_entry(oracle, ref):
    """The quirk itself: the same corner twice in ONE call adds two landmarks (the second copy is
    gated against the map as it was at call entry); in two calls the second is an Old update."""
    z, R = np.array([2.0, 1.0]), np.array([0.01, 0.0, 0.0, 0.02])
    x0, P0 = np.zeros(3), np.zeros((3, 3))
    x1, _ = ref.call_update(x0, P0, np.stack([z, z], axis=1), np.concatenate([R.reshape(2, 2)] * 2, axis=1))
    assert len(x1) == 7
    fo = oracle.new_filter(4)
    assert [t.decision for t in fo.update_chunk([z, z], [R, R])] == [0, 0] and fo.num_landmarks == 2
    x2, _ = fo.get_state()
    assert np.array_equal(x1, x2)
    fs = oracle.new_filter(4)
    assert [fs.update(z, R).decision, fs.update(z, R).decision] == [0, 1] and fs.num_landmarks == 1


def test_measurement_covariance_from_feature(ekf, oracle, ref):
    """slam.cpp:158-167: the synthetic driver, the restatement and the reference expressions agree bitwise."""
    rng = np.random.default_rng(7)
    for _ in range(200):
        fx, fy = rng.uniform(-8000, 8000, 2)
        z0, R0 = ekf.measurement_from_feature(fx, fy)
        z1, R1 = oracle.measurement_from_feature(fx, fy)
        z2, R2 = ref.measurement_from_feature(fx, fy)
        assert np.array_equal(z0, z1) and np.array_equal(z1, z2)
        assert np.array_equal(R0, R1) and np.array_equal(R1, R2)
        assert R0[1] == R0[2] or abs(R0[1] - R0[2]) < 1e-18


def test_edge_cases(oracle, ref):
    """Empty map -> New; first landmark covariance C R C^T; far measurement -> New; exact repeat -> Old;
    NaN measurement never associates."""
    for make in (lambda: oracle.new_filter(8), lambda: ref.new_filter()):
        f = make()
        z, R = np.array([2.0, 1.0]), np.array([0.01, 0.0, 0.0, 0.02])
        t = f.update(z, R)
        assert t.decision == 0 and t.opt_i == 0 and f.num_landmarks == 1
        x, P = f.get_state()
        assert np.allclose(x[3:], z) and np.allclose(P[3:, 3:], [[0.01, 0], [0, 0.02]]) and not P[:3].any()
        t = f.update(z, R)
        assert t.decision == 1 and t.opt_i == 3 and t.mahal == 0.0
        t = f.update(np.array([2.6, 1.0]), R)          # d^2 = 0.36/0.02... between the thresholds
        assert t.decision in (1, 2)
        t = f.update(np.array([-50.0, 40.0]), R)
        assert t.decision == 0 and f.num_landmarks == 2
        t = f.update(np.array([np.nan, 1.0]), R)       # NaN distance is never selected -> New (Opt_i == 0)
        assert t.decision == 0 and t.opt_i == 0
