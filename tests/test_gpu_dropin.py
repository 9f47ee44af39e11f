"""GPU test: the reference's SLAM loop (slam.cpp:127-182 call sequence, examples/slam_synthetic.cpp)
driving the CUDA core through the drop-in KalmanFilter class (2d-ekf-slam_b200/host/kalmanfilter.h),
checked line by line against the oracle on the same records."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOKENS = {"New": 0, "Old": 1, "Ignore": 2, "Full": 3}


def test_slam_loop_through_the_dropin_class(ekf, oracle, tmp_path):
    exe = str(tmp_path / "slam_synthetic")
    lib = os.path.join(ROOT, "2d-ekf-slam_b200", "lib")
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "oracle", "shim"),
                    "-I" + os.path.join(ROOT, "2d-ekf-slam_b200", "host"), "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "slam_synthetic.cpp"), "-L" + lib, "-lekf_slam_b200",
                    "-lekf_synth", "-Wl,-rpath," + lib, "-o", exe], check=True)
    N, T = 12, 240
    for cap in (16, 300):          # batch regime kernels, then the large-map (HBM) kernels
        out = subprocess.run([exe, str(N), str(T), str(cap)], stdout=subprocess.PIPE, text=True, check=True).stdout
        syn = ekf.Synth(N, steps_per_lap=T, max_meas=2, compass_every=10)
        rec = syn.generate(1, T)
        want = oracle.run_batch(rec, 2, N + 4, pose_trace=True)
        lines = out.strip().split("\n")
        dec, nlm, odom = [], [], []
        for ln in lines:
            if ln.startswith("Update:"):
                m = re.match(r"Update: (\w+) (\d+)$", ln)
                dec.append(TOKENS[m.group(1)])
                nlm.append(int(m.group(2)))
            else:
                odom.append([float(v) for v in ln.split()[1:]])
        wd = want["decision"][0].reshape(-1)
        assert dec == [d for d in wd if d >= 0]
        assert nlm[-1] == N
        odom = np.array(odom)
        assert odom.shape == (T, 3)
        err = np.abs(odom - want["pose_trace"][0]).max() / np.abs(want["pose_trace"][0]).max()
        assert err <= 1e-9, err


def test_log_files_match_the_reference_formats(ekf, ref, tmp_path):
    """odomRun / featuresRun / covRun / knownfeaturesRun / scanRun written through the
    drop-in class are the reference's text, line for line (default ostream formatting, the
    reference's own knownfeatures index stride, one laser scan per second of loop time as
    slam.cpp:184-203), so plot.py / RealTimePlotting.m work unchanged."""
    exe = str(tmp_path / "slam_synthetic")
    lib = os.path.join(ROOT, "2d-ekf-slam_b200", "lib")
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "oracle", "shim"),
                    "-I" + os.path.join(ROOT, "2d-ekf-slam_b200", "host"), "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "slam_synthetic.cpp"), "-L" + lib, "-lekf_slam_b200",
                    "-lekf_synth", "-Wl,-rpath," + lib, "-o", exe], check=True)
    N, T = 12, 240
    ours, theirs = tmp_path / "ours", tmp_path / "ref"
    ours.mkdir()
    theirs.mkdir()
    subprocess.run([exe, str(N), str(T), "16", str(ours)], stdout=subprocess.DEVNULL, check=True)
    syn = ekf.Synth(N, steps_per_lap=T, max_meas=2, compass_every=10)
    scans = [np.stack(v) for v in zip(*[syn.scan(t + 1) for t in range(T)])]     # the robot has made t+1 moves
    ref.run_logged(syn.generate(1, T)[0], 2, theirs, scans=scans)
    for name, min_lines in (("odomRun.txt", T), ("featuresRun.txt", T), ("covRun.txt", T), ("knownfeaturesRun.txt", T),
                            ("scanRun.txt", 200)):
        a = (ours / name).read_text().split("\n")
        b = (theirs / name).read_text().split("\n")
        assert len(a) == len(b) and len(b) > min_lines, name
        diff = 0
        for la, lb in zip(a, b):
            if la == lb:
                continue
            fa, fb = la.split(), lb.split()
            assert len(fa) == len(fb), name
            for va, vb in zip(fa, fb):      # 6 significant digits: a 1e-12 difference may flip the last one
                assert abs(float(va) - float(vb)) <= 2e-6 * max(abs(float(vb)), 1e-30), (name, la, lb)
            diff += 1
        assert diff <= len(b) // 100, "%s: %d of %d lines differ in the last printed digit" % (name, diff, len(b))


def test_hough_example_through_the_c_header(tmp_path):
    """examples/hough_synthetic.cpp: include/ekf_hough_b200.h from plain C++, a room corner 3 m ahead
    and 2.5 m to the left must come out as a corner feature near (3000, 2500) rotated by the turn."""
    exe = str(tmp_path / "hough_synthetic")
    lib = os.path.join(ROOT, "2d-ekf-slam_b200", "lib")
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "hough_synthetic.cpp"), "-L" + lib, "-lekf_slam_b200",
                    "-Wl,-rpath," + lib, "-o", exe], check=True)
    out = subprocess.run([exe, "4"], stdout=subprocess.PIPE, text=True, check=True).stdout.strip().split("\n")
    scans = [ln for ln in out if ln.startswith("scan")]
    assert len(scans) == 4
    found = 0
    s = -1
    for ln in out:
        if ln.startswith("scan"):
            s += 1
            continue
        fx, fy = (float(v) for v in ln.split()[1:])
        t = np.deg2rad(5.0 * s)
        cx, cy = 3000 * np.cos(t) + 2500 * np.sin(t), -3000 * np.sin(t) + 2500 * np.cos(t)   # the corner in the robot frame
        if np.hypot(fx - cx, fy - cy) < 60:
            found += 1
    assert found >= 3, out


def _build_example(tmp_path):
    exe = str(tmp_path / "slam_synthetic")
    lib = os.path.join(ROOT, "2d-ekf-slam_b200", "lib")
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "oracle", "shim"),
                    "-I" + os.path.join(ROOT, "2d-ekf-slam_b200", "host"), "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "slam_synthetic.cpp"), "-L" + lib, "-lekf_slam_b200",
                    "-lekf_synth", "-Wl,-rpath," + lib, "-o", exe], check=True)
    return exe


def test_dropin_class_grows_like_the_reference_and_runs_sharded(tmp_path):
    """Update.cpp:158-177 never refuses a New association. (i) The drop-in class started with room for 4
    landmarks in a 20-landmark world must print exactly what it prints with room for 24 (ekf_resize doubles
    the capacity before an update that could overflow: 4 -> 8 -> 16 -> 32, crossing nothing but memory);
    (ii) the same loop with the filter as ONE map column-sharded over three shards prints the same tokens
    and the same odometry to 1e-9."""
    exe = _build_example(tmp_path)
    run = lambda *a: subprocess.run([exe] + [str(v) for v in a], stdout=subprocess.PIPE, check=True, text=True).stdout
    big = run(20, 300, 24)
    small = run(20, 300, 4)
    assert "Full" not in small and small == big
    sharded = run(20, 300, 24, "-", "shards=3")
    la, lb = big.split("\n"), sharded.split("\n")
    assert len(la) == len(lb)
    for a, b in zip(la, lb):
        if a.startswith("odom"):
            for va, vb in zip(a.split()[1:], b.split()[1:]):
                assert abs(float(va) - float(vb)) <= 1e-9 * max(1.0, abs(float(va)))
        else:
            assert a == b
