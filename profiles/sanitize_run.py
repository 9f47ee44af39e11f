"""Tiny end-to-end workload for compute-sanitizer (memcheck / racecheck): every kernel family once.
    compute-sanitizer --tool memcheck python profiles/sanitize_run.py"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("ekf_b200", os.path.join(ROOT, "2d-ekf-slam_b200", "ekf_b200.py"))
ekf = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ekf)
from oracle_lib import Oracle  # noqa: E402

N, F, T, M, cap = 10, 5, 50, 2, 12
syn = ekf.Synth(N, steps_per_lap=T, max_meas=M, compass_every=4)
rec = np.ascontiguousarray(np.concatenate([syn.generate(F, T)] * 2, axis=1))
want = Oracle().run_batch(rec, M, cap, pose_trace=True)
for kern, name in ((1, "smem"), (2, "tile"), (3, "stile")):
    fb = ekf.FilterBatch(F, cap, batch_kernel=kern)
    got = fb.run(rec, M, pose_trace=True)
    assert np.array_equal(got["decision"], want["decision"]), name
    assert np.abs(got["pose_trace"] - want["pose_trace"]).max() < 1e-9, name
    fb.close()
    print("fused", name, "ok")
# per-call surface, batch regime and large regime (plain and TMA downdate)
for regime, name in ((1, "batch per-call"), (2, "large per-call")):
    fb = ekf.FilterBatch(2, cap, regime=regime)
    for t in range(12):
        r = rec[:2, t]
        fb.propagate(r[:, 0], r[:, 1], r[:, 2])
        if r[0, 6]:
            fb.update_compass(r[:, 3], r[:, 4])
        nz = int(r[0, 5])
        if nz:
            zr = r[:, 8:8 + 6 * nz].reshape(2, nz, 6)
            fb.update(zr[:, :, :2], zr[:, :, 2:])
    fb.get_state(0)
    fb.close()
    print(name, "ok")
fb = ekf.FilterBatch(1, cap, regime=2)
fb.run(np.ascontiguousarray(rec[:1, :30]), M)
fb.close()
print("large fused ok")
