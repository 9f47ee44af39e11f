"""Per-call path of the batch regime (one launch per reference call, covariance through HBM):
SURVEY.md 8d "per-call path bytes". F filters x 50 landmarks, complete maps, then timed
doPropagation + doUpdate calls through the C ABI with device-resident state.

    python profiles/percall_run.py [F] [steps]    -> one JSON line
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

if __name__ == "__main__":
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    ekf = bench.load_product()
    N, T = 50, 1000
    syn = ekf.Synth(N, steps_per_lap=T)
    rec = syn.generate(F, T)
    fb = ekf.FilterBatch(F, N)
    fb.run(rec, 1, trace=False, allow_capacity=True)            # build the maps with the fused kernel
    n = 3 + 2 * N
    for t in range(3):                                           # warm-up calls
        r = rec[:, t]
        fb.propagate(r[:, 0], r[:, 1], r[:, 2])
        fb.update(r[:, 8:10], r[:, 10:14], want=False)
    fb.sync(allow_capacity=True)
    # the caller's per-step input arrays (contiguous, as a driver loop would hold them)
    ins = [tuple(np.ascontiguousarray(a) for a in (rec[:, t, 0], rec[:, t, 1], rec[:, t, 2], rec[:, t, 8:10], rec[:, t, 10:14]))
           for t in range(3, 3 + steps)]
    fb.kernel_time()
    t0 = time.perf_counter()
    for v, w, d, z, R in ins:
        fb.propagate(v, w, d)
        fb.update(z, R, want=False)
    fb.sync(allow_capacity=True)
    dt = (time.perf_counter() - t0) / steps
    kms, kn = fb.kernel_time()
    alg = F * (2 * 8 * n * n + 6 * 8 * n)                        # update: read + write P; propagate: the 3 x n strip
    print(json.dumps({"workload": "%d filters x %d landmarks, per-call doPropagation + doUpdate (host arrays of inputs per call)" % (F, N),
                      "ms_per_step": dt * 1e3, "filter_steps_per_s": F / dt, "algorithmic_GBs": alg / dt / 1e9,
                      "frac_of_hbm_peak_6560": alg / dt / 1e9 / 6560.0,
                      "fused_pair_kernel_ms": kms, "kernel_launches_timed": kn,
                      "kernel_algorithmic_GBs": alg / (kms * 1e-3) / 1e9 if kn else None,
                      "kernel_frac_of_hbm_peak_6560": alg / (kms * 1e-3) / 1e9 / 6560.0 if kn else None}))
    fb.close()
