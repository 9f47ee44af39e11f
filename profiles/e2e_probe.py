"""Where the end-to-end lap of ekf_run() differs from the device-resident lap: one rank's share of the
8-GPU bench (8,192 filters x 50 landmarks x 1,000 steps by default) on one GPU.

    python profiles/e2e_probe.py [F]     -> one JSON line (ms per lap)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

if __name__ == "__main__":
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    ekf = bench.load_product()
    T, N, CAP = 1000, 50, 56
    syn = ekf.Synth(N, steps_per_lap=T)
    pinned = ekf.PinnedArray((F, T, 14))
    syn.generate(F, T, out=pinned.array)
    rec = pinned.array
    fb = ekf.FilterBatch(F, CAP)
    outs = fb.alloc_outputs(T, 1, trace=True, pinned=True)
    outs_small = fb.alloc_outputs(T, 1, trace=False, pinned=True)
    res = {"filters": F}

    def lap(name, fn, reps=4):
        fn()
        fb.sync(allow_capacity=True)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        fb.sync(allow_capacity=True)
        res[name] = (time.perf_counter() - t0) / reps * 1e3

    fb.run(rec, 1, outputs=outs, allow_capacity=True)           # builds the maps
    fb.upload_records(rec, 1)
    lap("resident_ms", lambda: fb.run_resident(trace=True))
    fb.kernel_time()
    lap("e2e_full_outputs_ms", lambda: fb.run(rec, 1, outputs=outs, allow_capacity=True))
    kms, kn = fb.kernel_time()
    res["e2e_kernel_sum_ms"] = kms * kn / 5
    res["e2e_chunks"] = kn / 5
    lap("e2e_final_state_only_ms", lambda: fb.run(rec, 1, outputs=outs_small, allow_capacity=True))
    lap("upload_only_ms", lambda: fb.upload_records(rec, 1))
    print(json.dumps(res))
    fb.close()
