"""How often does a complete 50-landmark map see a spurious New association (profiles/README.md)?
    python profiles/drop_count.py [F] [laps] [cap]
Runs F filters of the bench's synthetic world for `laps` laps (lap 0 builds the maps) and counts the New /
DROPPED decisions of every later lap, per filter."""
import importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ekf_b200", os.path.join(ROOT, "2d-ekf-slam_b200", "ekf_b200.py"))
ekf = importlib.util.module_from_spec(spec); spec.loader.exec_module(ekf)
F = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
laps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 50
t0 = time.perf_counter()
syn = ekf.Synth(50, steps_per_lap=1000)
pin = ekf.PinnedArray((F, 1000, syn.record_len))
syn.generate(F, 1000, out=pin.array)
print("generated %d filters in %.1f s" % (F, time.perf_counter() - t0), flush=True)
fb = ekf.FilterBatch(F, cap)
fb.upload_records(pin.array, 1)
outs = fb.alloc_outputs(1000, 1, trace=True, pinned=True)
tot_new = tot_drop = 0
bad = set()
for lap in range(laps):
    fb.timer_start()
    fb.run_resident(trace=True)
    ms = fb.timer_stop()
    o = fb.download_outputs(trace=True, outputs=outs, allow_capacity=True)
    d = o["decision"]
    n_new = int((d == 0).sum()); n_drop = int((d == 3).sum())
    if lap > 0:
        tot_new += n_new; tot_drop += n_drop
        bad.update(np.nonzero(((d == 0) | (d == 3)).any(axis=(1, 2)))[0].tolist())
    print("lap %d: %.1f ms  %.3e filter-steps/s  New %d  Dropped %d  max n_lm %d" % (lap, ms, F * 1000 / (ms * 1e-3), n_new, n_drop, int(o["final_nlm"].max())), flush=True)
print("laps 1..%d: %d New, %d dropped, %d filters affected of %d" % (laps - 1, tot_new, tot_drop, len(bad), F))
fb.close()
