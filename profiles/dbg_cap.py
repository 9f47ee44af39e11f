"""Stress / debugging aid for ekf_sharded_run: a small map (14 landmarks), many operations.

    python profiles/dbg_cap.py <compass_every> <devices> <capacity> <steps>
"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
ekf = bench.load_product()
N, T, cap, M = 14, int(sys.argv[4]) if len(sys.argv) > 4 else 160, int(sys.argv[3]) if len(sys.argv) > 3 else 11, 2
ce = int(sys.argv[1]) if len(sys.argv) > 1 else 6
devs = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 0, 0]
syn = ekf.Synth(N, steps_per_lap=min(T, 400), max_meas=M, compass_every=ce)
rec = syn.generate(1, T)
sm = ekf.ShardedMap(devs, cap)
print(sm.run_mode())
t0 = time.time()
try:
    got = sm.run(rec, M, trace=True, pose_trace=True, allow_capacity=True)
    print("ok", np.bincount(got["decision"].ravel() + 1))
except Exception as e:
    print("ERR", e)
print("elapsed %.2f s" % (time.time() - t0))
