"""Short, fixed workloads for ncu (one GPU). Usage:
    python profiles/prof_run.py batch [F] [T]      # fused batch kernel: F filters x T steps, 3 launches
    python profiles/prof_run.py large [N] [steps]  # regime B: injected N-landmark map, update steps
The first launch(es) build the maps; profile the LAST launch (ncu -s/-c as in profiles/README.md).
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("ekf_b200", os.path.join(ROOT, "2d-ekf-slam_b200", "ekf_b200.py"))
ekf = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ekf)


def batch(F=592, T=1000, cap=50):
    syn = ekf.Synth(50, steps_per_lap=T)
    rec = syn.generate(F, T)
    fb = ekf.FilterBatch(F, cap, batch_kernel=int(os.environ.get("EKF_KERNEL", "0")))
    fb.upload_records(rec, 1)
    if os.environ.get("EKF_BUILD_WITH_SMEM"):   # experiment builds of the tile kernel without the New path
        fb.set_batch_kernel(1)
    for lap in range(3):                   # lap 0 builds the maps, laps 1-2 are full size
        if lap == 1:
            fb.set_batch_kernel(int(os.environ.get("EKF_KERNEL", "0")))
        if lap == 2 and os.environ.get("EKF_PHASES"):
            fb.sync()
            ekf.debug_phase_cycles(read=False)
        fb.run_resident(trace=True)
    out = fb.download_outputs(trace=True)
    if os.environ.get("EKF_PHASES"):
        cyc = ekf.debug_phase_cycles()
        names = ["scalar chains", "cov propagate+publish", "gating", "decision+column publish", "gain rows",
                 "downdate+publish", "step epilogue", "-"]
        per = -(-F // (2 * 148))           # filters CTA 0 processed
        print("phase cycles per step (CTA 0, %d filter(s) x %d steps):" % (per, T))
        for n, c in zip(names, cyc[:8]):
            print("  %-26s %8.0f" % (n, c / (per * T)))
        print("  %-26s %8.0f" % ("total", sum(cyc[:8]) / (per * T)))
        if any(cyc[8:]):
            fine = ["gate: loads+prelude", "gate: S terms", "gate: group barrier", "gate: finish", "gate: warp argmin",
                    "gate: candidate store", "gate: CTA barrier", "decision", "col publish (own)", "barrier (post thread)",
                    "gain row (own)", "barrier", "downdate (own tile)", "publish (own)", "barrier", "trace stores",
                    "cp.async wait", "end-of-step barrier", "step start (own)", "barrier (scalar chains)",
                    "propagate (own)", "barrier (robot block/strip)", "-", "-"]
            for n, c in zip(fine, cyc[8:]):
                print("    %-24s %8.0f" % (n, c / (per * T)))
    if os.environ.get("EKF_STILE_TS"):      # -DEKF_STILE_TIMING builds: barrier arrivals of one step
        ts = ekf.debug_stile_timestamps()
        names = ["step start", "B1 propagate done", "B2 gating terms (named)", "B3 gating done", "B4 P*H^T rows / S^-1",
                 "B5 gain, x, W", "B6 downdate", "B7 end of step", "after B7", "  (decision known)", "  (rows / S^-1 done)"]
        t0 = ts[:, 0][ts[:, 0] > 0].min()
        print("barrier arrivals of CTA 0, first filter, step 500 (cycles since the first warp entered the step):")
        print("  %-26s %s   phase (max-to-max)" % ("", " ".join("   w%d" % w for w in range(8))))
        prev = None
        for kk, nm in enumerate(names):
            row = ts[:, kk]
            cells = " ".join("%5d" % (v - t0) if v > 0 else "    -" for v in row)
            mx = row[row > 0].max() - t0 if (row > 0).any() else 0
            print("  %-26s %s   %s" % (nm, cells, "" if prev is None else "%5d" % (mx - prev)))
            prev = mx
    if os.environ.get("EKF_DTILE_TS"):      # -DEKF_DTILE_TIMING builds (make timing; EKF_B200_LIB=.../libekf_slam_b200_timing.so)
        ts = ekf.debug_dtile_timestamps()
        names = ["step start (before B1; warp 2 = helper chain done)", "after B1", "before gating (after the sweep, if any)",
                 "gating done (candidate written)", "after B2", "-", "gain phase / pose rows done", "-", "step end (warp 0: record scalars done)",
                 "-", "-"]
        t0 = ts[:, 0][ts[:, 0] > 0].min()
        print("stamps of CTA 0, first filter, step 501 (cycles since the first warp entered the step); front warp = the one with stamps 2,3:")
        print("  %-40s %s" % ("", " ".join("    w%d" % w for w in range(4))))
        for kk, nm in enumerate(names):
            row = ts[:, kk]
            ref = t0 if kk < 9 else ts[:, 9][ts[:, 9] > 0].min() if (ts[:, 9] > 0).any() else t0
            print("  %-40s %s" % (nm, " ".join("%6d" % (v - ref) if v > 0 else "     -" for v in row)))
    assert (out["final_nlm"] == 50).mean() > 0.99
    ms, n = fb.kernel_time()
    print("batch: F=%d T=%d avg kernel %.3f ms over %d launches -> %.3e filter-steps/s"
          % (F, T, ms, n, F * T / (ms * 1e-3)))
    fb.close()


def large(N=2000, steps=20):
    from parity import injected_state
    syn = ekf.Synth(N, steps_per_lap=10 ** 7, max_meas=1)
    rec = syn.generate(1, steps)
    x0, P0 = injected_state(syn.world(), seed=N)
    fb = ekf.FilterBatch(1, N + 2)
    fb.set_state(0, x0, P0, symmetric=True)
    fb.upload_records(rec, 1)
    fb.run_resident(trace=True)
    out = fb.download_outputs(trace=True)
    ms, n = fb.kernel_time()
    print("large: N=%d steps=%d old=%d avg downdate %.4f ms over %d launches"
          % (N, steps, int((out["decision"] == 1).sum()), ms, n))
    fb.close()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "batch"
    a = [int(v) for v in sys.argv[2:]]
    (batch if mode == "batch" else large)(*a)
