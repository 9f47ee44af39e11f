"""GPU test of the whole measurement-to-filter chain, the loop body of slam.cpp:130-182 with every
computation on the GPU: a laser scan goes through ekf_hough_get_features (Hough lines, segments,
corners, structural compass), each corner becomes (z, R) as slam.cpp:150-167 does, and the filter
takes doPropagation / doUpdateCompass / doUpdate through the per-call C ABI. The same loop on the
CPU oracles (hough_oracle / features_oracle / ekf_oracle) must give the same associations and the
same state."""
import numpy as np
import pytest

import scan_synth
from hough_lib import HoughOracle
from parity import TOL, assert_state_close

pytestmark = pytest.mark.gpu


def test_scans_to_filter_updates(ekf, oracle):
    ho = HoughOracle()
    room = scan_synth.make_room(seed=4, n_boxes=2)
    rng = np.random.default_rng(8)
    T, dt = 120, 0.2
    cx, cy, rad = 4500.0, 3500.0, 1500.0                   # the robot drives a circle in the middle of the room
    v_mm_s = 300.0                                         # movementcontroller.cpp:36
    w = v_mm_s / rad
    fb = ekf.FilterBatch(1, 60)
    hb = ekf.HoughBatch(1)
    of = oracle.new_filter(60)
    off_gpu = np.array([100.0])
    off_cpu = 100.0
    n_updates = n_old = n_compass = 0
    for t in range(T):
        a = w * dt * t
        px, py, phi = cx + rad * np.sin(a), cy - rad * np.cos(a), a
        x, y, r = scan_synth.scan_from_pose(room, px, py, phi, rng)
        vel = v_mm_s * (1 + 0.01 * rng.normal())
        rot = np.rad2deg(w) * (1 + 0.01 * rng.normal())
        # --- GPU
        fb.propagate(vel, rot, dt)
        pose, _ = fb.get_pose()
        got = hb.get_features(x[None], y[None], r[None], cur_phi=pose[:, 2].copy(), offset=off_gpu, max_feats=16)
        off_gpu = got["offset"]
        # --- CPU oracles
        of.propagate(vel, rot, dt)
        feats, _, compass, off_cpu, _ = ho.get_features(x, y, r, float(of.pose()[2]), off_cpu)
        assert got["n_feats"][0] == len(feats) and np.array_equal(got["feats"][0, :len(feats)], feats), "features, step %d" % t
        assert got["compass"][0] == compass and off_gpu[0] == off_cpu, "compass, step %d" % t
        if compass != 100.0:                               # slam.cpp:144-147
            fb.update_compass(compass, 0.0005)
            of.update_compass(compass, 0.0005)
            n_compass += 1
        for fx, fy in feats:                               # slam.cpp:150-171
            z, R = ekf.measurement_from_feature(fx, fy)
            dec, idx, mah = fb.update(z, R)
            n_before = of.n
            tr = of.update(z, R)
            assert dec[0, 0] == tr.decision and idx[0, 0] == (n_before if tr.decision == 0 else tr.opt_i), "step %d" % t
            assert abs(mah[0, 0] - tr.mahal) <= TOL * max(1.0, abs(tr.mahal))
            n_updates += 1
            n_old += tr.decision == 1
    xg, Pg = fb.get_state(0)
    xo, Po = of.get_state()
    assert_state_close(xg, Pg, xo, Po, "after %d steps" % T)
    assert n_updates >= T and n_old >= n_updates // 3 and n_compass >= T // 2
    fb.close()
    hb.close()
