import os, sys, time, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench
ekf = bench.load_product()
N, T, cap, M = 14, 160, int(sys.argv[3]) if len(sys.argv) > 3 else 11, 2
ce = int(sys.argv[1]) if len(sys.argv) > 1 else 6
devs = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 0, 0]
syn = ekf.Synth(N, steps_per_lap=T, max_meas=M, compass_every=ce)
rec = syn.generate(1, T)
sm = ekf.ShardedMap(devs, cap)
t0 = time.time()
try:
    got = sm.run(rec, M, trace=True, pose_trace=True, allow_capacity=True)
    print("ok", np.bincount(got["decision"].ravel() + 1))
except Exception as e:
    print("ERR", e)
print("elapsed %.2f s" % (time.time() - t0))
