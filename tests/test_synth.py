"""CPU tests (not gpu) of the synthetic odometry / measurement driver (host logic)."""
import numpy as np


def test_generation_is_deterministic_and_range_addressable(ekf):
    syn = ekf.Synth(20, steps_per_lap=200, max_meas=2, compass_every=4)
    a = syn.generate(6, 200)
    b = syn.generate(6, 200, n_threads=1)
    assert np.array_equal(a, b)
    part = syn.generate(2, 50, f0=3, t0=100)            # any sub-range reproduces the same bits:
    assert np.array_equal(part, a[3:5, 100:150])        # this is what makes per-rank generation exact


def test_records_respect_the_reference_sensor_envelope(ekf):
    syn = ekf.Synth(50, steps_per_lap=1000)
    rec, ids = syn.generate(3, 1000, want_ids=True)
    assert (rec[:, :, 5] == 1).all() and (ids >= 0).all()
    z = rec[:, :, 8:10]
    d = np.hypot(z[..., 0], z[..., 1])
    assert d.min() > 1.0 and d.max() < 8.0               # featuredetector.h:28, houghtransform.h:20
    assert (z[..., 0] > 0).all()                         # 180 degree field of view (slam.cpp:90)
    assert len(np.unique(ids)) == 50, "every landmark is observed within one lap"
    R = rec[:, :, 10:14]
    assert np.allclose(R[..., 1], R[..., 2], atol=1e-18)
    det = R[..., 0] * R[..., 3] - R[..., 1] * R[..., 2]
    assert (det > 0).all()


def test_truth_is_the_filter_motion_model_and_closes(ekf):
    syn = ekf.Synth(20, steps_per_lap=500)
    p0, pT = syn.true_pose(0), syn.true_pose(500)
    assert np.allclose(p0, [0, 0, 0], atol=1e-12)
    assert np.allclose(pT[:2], p0[:2], atol=1e-9) and abs(pT[2] - 2 * np.pi) < 1e-12
    # Euler unicycle step (Propagate.cpp:33-37) with the true controls lands on the next vertex
    cfg = syn.cfg
    v = 2 * cfg.radius * np.sin(np.pi / 500) / cfg.dt
    w = 2 * np.pi / 500 / cfg.dt
    for t in (0, 7, 123, 499):
        a, b = syn.true_pose(t), syn.true_pose(t + 1)
        pred = a + cfg.dt * np.array([v * np.cos(a[2]), v * np.sin(a[2]), w])
        assert np.allclose(pred, b, atol=1e-9)


def test_odometry_units_are_what_the_robot_reports(ekf):
    syn = ekf.Synth(20, steps_per_lap=1000)
    rec = syn.generate(200, 50)
    v_m = rec[:, :, 0] / 1000.0                           # mm/s  (kalmanfilter.cpp:18,26)
    w_m = rec[:, :, 1] * 3.141592654 / 180.0              # deg/s (kalmanfilter.cpp:19)
    v = 2 * syn.cfg.radius * np.sin(np.pi / 1000) / syn.cfg.dt
    w = 2 * np.pi / 1000 / syn.cfg.dt
    assert abs(v_m.mean() - v) < 3e-4 and abs(w_m.mean() - w) < 1e-3
    assert abs(v_m.std() - 0.01 * v) < 2e-4 and abs(w_m.std() - 0.04 * v) < 1e-3
