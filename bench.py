#!/usr/bin/env python
"""bench.py - filter-steps/s of the batched EKF-SLAM filter core (N=50 landmarks) on B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA core
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU code (oracle/_ref)

The workload is BASELINE configs[2], the one the metric and the north-star target are quoted on:
65,536 filters x 50 landmarks x 1,000 steps - all on one GPU at N=1, the same 65,536 split into
disjoint filter ranges over the ranks under torchrun (STRONG scaling; a weak run with
--weak-filters-per-gpu filters per GPU rides along as an extra key). One bench "step" is one pass of
the hot path over the batch: every filter of the rank's range runs one full lap (1,000 iterations of
the slam.cpp:130-182 loop: doPropagation + one doUpdate) with its 50-landmark map already built
during warm-up, so every timed filter-step is a full-size (n = 103) propagate + gating + update.
`value` = filter-steps/s with the step records resident in HBM; `e2e` = the same through ekf_run()
with pinned HOST buffers (H2D of the records and D2H of the decisions inside the timed region). No
data-path collective; torch.distributed is only used for the barrier and the max / sum reductions.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "filter-steps/sec (batched EKF, N=50)"
UNIT = "filter-steps/s"
N_LM = 50
CAP_LM = 56          # landmark capacity per filter. The world has 50 landmarks; a few filters of a large
                     # population see a spurious New association later (an outlier beyond Gamma_max) and, like the
                     # reference (Update.cpp:158-177), grow their map to 51+. Every filter starts in the
                     # four-filters-per-SM kernel instance (tiles for 50 landmarks); the ones that outgrow it are
                     # finished by the instance sized for CAP_LM. Nothing is dropped (asserted over all ranks and laps).
FAST_TILES_LM = 50
T_LAP = 1000
MAX_MEAS = 1
COMPASS_EVERY = 0


def load_product():
    spec = importlib.util.spec_from_file_location("ekf_b200", os.path.join(ROOT, "2d-ekf-slam_b200", "ekf_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ekf_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


# ---- algorithmic work (DESIGN.md "Algorithmic flops / bytes"; SURVEY.md 8d F_min) ------------------
def fmin_flops(n_lm, decision):
    """Minimum FP64 flops of one filter-step with n_lm landmarks in the map."""
    n = 3 + 2 * n_lm
    f = 8 * n_lm + 170                      # propagate
    f += 150 * n_lm                         # gating: H, S, cond, Mahalanobis per landmark
    if decision == 1:                       # Old: gain + state + symmetric rank-2 downdate
        f += 2 * n * n + 42 * n
    return f


def large_bytes(n_lm):
    n = 3 + 2 * n_lm
    return 16 * n * n + 56 * n + 96 * n_lm   # per Old update, regime B


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize(local)


def max_over_ranks(dist, local, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cpu_baseline(records_one_lap, kind_pref="reference", seconds_target=12.0):
    """The reference's CPU filter on the box's host cores: a bounded sample of the same workload
    (same records), one filter per task, all hardware threads, timed on the full-map lap only."""
    from oracle_lib import Oracle, Ref
    if kind_pref == "reference" and Ref.available():
        chk, kind = Ref(), "reference"
    else:
        chk, kind = Oracle(), "port"
    cores = os.cpu_count() or 1
    T = records_one_lap.shape[1]
    # calibrate on one filter, then size the sample for ~seconds_target of wall time
    one = np.ascontiguousarray(np.concatenate([records_one_lap[:1]] * 2, axis=1))
    t0 = time.perf_counter()
    chk.run_batch(one, MAX_MEAS, CAP_LM, n_threads=1, trace=False)
    per_filter = max(time.perf_counter() - t0, 1e-3)          # 2 laps, 1 thread
    n_f = int(max(cores, min(records_one_lap.shape[0], cores * max(1.0, seconds_target / per_filter))))
    n_f = max(cores, (n_f // cores) * cores)
    n_f = min(n_f, records_one_lap.shape[0])
    rec = np.ascontiguousarray(np.concatenate([records_one_lap[:n_f]] * 2, axis=1))
    r = chk.run_batch(rec, MAX_MEAS, CAP_LM, n_threads=cores, trace=False, warm_steps=T)
    assert not r["bad"]
    assert (r["final_nlm"] >= N_LM).all()
    val = n_f * T / r["seconds"]
    return {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d filters x %d full-map steps (after an untimed %d-step map-building lap), "
                      "one filter per task over %d host threads; %s" %
                      (n_f, T, T, cores,
                       "reference odometry/*.cpp compiled unmodified over the stand-in Eigen, -O2" if kind == "reference"
                       else "C restatement oracle/ekf_oracle.c, -O2")}, chk


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ekf = load_product()
    cores = os.cpu_count() or 1
    syn = ekf.Synth(N_LM, steps_per_lap=T_LAP, max_meas=MAX_MEAS, compass_every=COMPASS_EVERY)
    n_f = max(cores, 2 * cores)
    lap = syn.generate(n_f, T_LAP)
    from oracle_lib import Oracle, Ref
    chk, kind = (Ref(), "reference") if Ref.available() else (Oracle(), "port")
    rec = np.ascontiguousarray(np.concatenate([lap] * 2, axis=1))
    times = []
    for i in range(args.warmup + args.steps):
        r = chk.run_batch(rec, MAX_MEAS, CAP_LM, n_threads=cores, trace=False, warm_steps=T_LAP)
        assert not r["bad"]
        if i >= args.warmup:
            times.append(r["seconds"])
    total = sum(times)
    val = n_f * T_LAP * args.steps / total
    sample = ("each step: %d filters x %d full-map steps (after an untimed map-building lap) over %d host "
              "threads" % (n_f, T_LAP, cores))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.filters_per_gpu <= 0 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, 1), "filters_per_step_sample": n_f,
                       "landmarks": N_LM, "steps_per_lap": T_LAP},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(args, world):
    if args.filters_per_gpu:
        return ("%d filters x %d landmarks x %d steps per GPU, chip-resident covariance (weak scaling%s)"
                % (args.filters_per_gpu, N_LM, T_LAP, "" if world == 1 else ", %d GPUs, disjoint filter ranges" % world))
    return ("%d filters x %d landmarks x %d steps, chip-resident covariance (BASELINE configs[2]%s)"
            % (args.filters, N_LM, T_LAP,
               ", all on 1 GPU" if world == 1 else ", split over %d GPUs: %d filters each, disjoint ranges" % (world, args.filters // world)))


def sum_over_ranks(dist, local, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def profile_facts():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_latest.json")))
    except Exception:
        return {}


def verify_large_map(fb, rec, x0, P0, k):
    """The first k update steps of the large-map leg against the CPU oracle (checker only): decisions and
    indices exact, state and covariance (diagonal + 64 sampled columns) within 1e-9 norm-wise."""
    from oracle_lib import Oracle
    from parity import TOL
    n = len(x0)
    of = Oracle().new_filter((n - 3) // 2).set_state(x0, P0)
    got = fb.run(np.ascontiguousarray(rec[:, :k]), 1, trace=True)
    n_old = 0
    for t in range(k):
        r = rec[0, t]
        of.propagate(r[0], r[1], r[2])
        n_before = of.n
        tr = of.update(r[8:10], r[10:14])
        assert got["decision"][0, t, 0] == tr.decision, "large-map decision differs from the oracle at step %d" % t
        assert got["index"][0, t, 0] == (n_before if tr.decision == 0 else tr.opt_i), "large-map index differs at step %d" % t
        n_old += tr.decision == 1
    x, P = fb.get_state(0)
    xr, Pr = of.get_state()
    cols = np.unique(np.concatenate([np.random.default_rng(1).integers(0, n, 64), [0, 1, 2, n - 2, n - 1]]))
    scale = float(np.abs(np.diagonal(Pr)).max())
    err = max(float(np.abs(P[:, cols] - Pr[:, cols]).max()), float(np.abs(np.diagonal(P) - np.diagonal(Pr)).max()))
    ex = float(np.abs(x - xr).max() / np.abs(xr).max())
    assert err <= TOL * scale and ex <= TOL, "large-map state differs from the oracle: P %g of %g, x %g" % (err, scale, ex)
    return {"steps": k, "old_updates": int(n_old), "decisions_and_indices": "exact",
            "cov_max_abs_err_over_scale": err / scale, "state_rel_err": ex,
            "compared": "diagonal + %d sampled columns of P, full x" % len(cols)}


def large_map_leg(ekf, n_lm, steps, hbm_peak, device, verify_steps=0):
    """BASELINE configs[3]/[4]: one large map, covariance in HBM, Old-updates from an injected state."""
    from parity import injected_state
    # a (nearly) stationary robot: the same few visible landmarks are re-observed, every update is
    # an Old-update over the full n x n covariance
    syn = ekf.Synth(n_lm, steps_per_lap=10 ** 7, max_meas=1)
    rec = syn.generate(1, steps)
    x0, P0 = injected_state(syn.world(), seed=n_lm)
    fb = ekf.FilterBatch(1, n_lm + 2, device=device)
    fb.set_state(0, x0, P0, symmetric=True)
    verified = None
    if verify_steps > 0:                       # parity at size, before anything is timed
        verified = verify_large_map(fb, rec, x0, P0, verify_steps)
        fb.set_state(0, x0, P0, symmetric=True)
    sweep = fb.large_downdate_kernel()
    fb.upload_records(rec, 1)
    fb.run_resident(trace=True)            # warm-up pass
    fb.sync()
    fb.set_state(0, x0, P0, symmetric=True)   # timed pass starts from the same injected state
    del P0
    fb.kernel_time()
    l0 = fb.kernel_launches()
    fb.timer_start()
    fb.run_resident(trace=True)
    ms = fb.timer_stop()
    l1 = fb.kernel_launches()
    out = fb.download_outputs(trace=True)
    kms, kn = fb.kernel_time()
    n_old = int((out["decision"] == 1).sum())
    fb.close()
    alg = large_bytes(n_lm)
    res = {"workload": "1 map x %d landmarks (n=%d, P=%.2f GB), %d update steps" %
                       (n_lm, 3 + 2 * n_lm, 8.0 * (3 + 2 * n_lm) ** 2 / 1e9, steps),
           "old_updates": n_old, "steps": steps, "ms_per_step": ms / steps,
           "update_steps_per_s": steps / (ms * 1e-3), "gpu_launches": int(l1 - l0),
           "verified_against_oracle": verified,
           "roofline": {"bound": "hbm", "kernel": sweep + "<2,0>", "achieved": alg / (kms * 1e-3) / 1e9 if kn else None,
                        "peak": hbm_peak, "unit": "GB/s", "frac": (alg / (kms * 1e-3) / 1e9) / hbm_peak if kn else None,
                        "traffic": profile_facts().get("large_traffic_bytes_per_launch", {}).get(sweep, {}).get(str(n_lm)),
                        "algorithmic_bytes_per_launch": alg, "avg_kernel_ms": kms,
                        "launches_timed": kn},
           "step_gbs": alg * n_old / (ms * 1e-3) / 1e9}
    return res


def sharded_map_leg(ekf, n_lm, steps, hbm_peak, devices):
    """the same large-map workload with the covariance column-sharded over
    several GPUs (one process, peer stores over NVLink for the one exchange the path has)."""
    from parity import injected_state
    syn = ekf.Synth(n_lm, steps_per_lap=10 ** 7, max_meas=1)
    rec = syn.generate(1, steps)
    x0, P0 = injected_state(syn.world(), seed=n_lm)
    sm = ekf.ShardedMap(devices, n_lm + 2)
    mode = sm.run_mode()
    sm.set_state(x0, P0, symmetric=True)
    sm.run(rec, 1, trace=True)                 # warm-up pass
    sm.set_state(x0, P0, symmetric=True)
    del P0
    l0 = sm.kernel_launches()
    out = sm.run(rec, 1, trace=True)
    ms, dms = sm.last_run_ms()
    l1 = sm.kernel_launches()
    n_old = int((out["decision"] == 1).sum())
    cols = [sm.columns(s) for s in range(len(devices))]
    sm.close()
    alg = large_bytes(n_lm)
    n = 3 + 2 * n_lm
    alg0 = 16.0 * n * (cols[0][1] - cols[0][0])      # shard 0's share of the downdate traffic
    G = len(devices)
    return {"workload": "1 map x %d landmarks (n=%d, P=%.2f GB) column-sharded over devices %s, %d update steps" %
                        (n_lm, n, 8.0 * n * n / 1e9, list(devices), steps),
            "shards": G, "old_updates": n_old, "steps": steps, "ms_per_step": ms / steps,
            "update_steps_per_s": steps / (ms * 1e-3), "gpu_launches": int(l1 - l0),
            "step_gbs_aggregate": alg * n_old / (ms * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": mode + " (shard 0, one launch sampled mid-run)",
                         "achieved": alg0 / (dms * 1e-3) / 1e9 if dms > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                         "frac": (alg0 / (dms * 1e-3) / 1e9) / hbm_peak if dms > 0 else None, "traffic": None,
                         "algorithmic_bytes_per_launch": alg0, "avg_kernel_ms": dms, "launches_timed": 1 if dms > 0 else 0}}


def hough_leg(ekf, n_scans, hbm_peak, device, with_cpu=True):
    """HoughTransform::getLines for a batch of synthetic LMS-200 scans."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    import scan_synth
    from hough_lib import HoughRef, HoughOracle
    base = 256
    X0, Y0, R0 = scan_synth.make_scans(base, seed=7)
    reps = -(-n_scans // base)
    X = np.ascontiguousarray(np.tile(X0, (reps, 1))[:n_scans])
    Y = np.ascontiguousarray(np.tile(Y0, (reps, 1))[:n_scans])
    R = np.ascontiguousarray(np.tile(R0, (reps, 1))[:n_scans])
    hb = ekf.HoughBatch(n_scans, device=device)
    hb.upload(X, Y, R)
    for _ in range(3):
        hb.run_resident()
    hb.sync()
    hb.kernel_time()
    K = 5
    for _ in range(K):
        hb.run_resident()
    hb.sync()
    ms, n = hb.kernel_time()
    lines, n_lines = hb.download()
    # end to end: pinned host buffers in, lines out (chunked H2D / kernels / D2H pipeline inside the call)
    pin = [ekf.PinnedArray(X.shape, np.float64), ekf.PinnedArray(Y.shape, np.float64), ekf.PinnedArray(R.shape, np.uint32),
           ekf.PinnedArray((n_scans, 32, 3), np.float64), ekf.PinnedArray((n_scans,), np.int32)]
    pin[0].array[:] = X
    pin[1].array[:] = Y
    pin[2].array[:] = R
    hb.get_lines(pin[0].array, pin[1].array, pin[2].array, max_lines=32, want_peaks=False, split=False,
                 out=(pin[3].array, pin[4].array))
    t0 = time.perf_counter()
    for _ in range(5):
        got = hb.get_lines(pin[0].array, pin[1].array, pin[2].array, max_lines=32, want_peaks=False, split=False,
                           out=(pin[3].array, pin[4].array))
    e2e_s = (time.perf_counter() - t0) / 5
    assert np.array_equal(pin[4].array, n_lines)
    got = {"lines": pin[3].array.copy(), "n_lines": pin[4].array.copy()}
    # the whole front end (lines, segments, corners, structural compass), host arrays in, features out
    feat = hb.get_features(pin[0].array, pin[1].array, pin[2].array, max_feats=16, want_lines=False)
    t0 = time.perf_counter()
    for _ in range(3):
        feat = hb.get_features(pin[0].array, pin[1].array, pin[2].array, max_feats=16, want_lines=False)
    feat_s = (time.perf_counter() - t0) / 3
    hb.close()
    for pa in pin:
        pa.free()
    points = X.shape[1]
    alg = n_scans * (points * 20 + 2 * 2 * 200 * 4 + float(n_lines.mean()) * 24 + 4)   # readings in; peak slots + counts out and in again; lines out
    res = {"workload": "%d scans x %d readings (synthetic LMS-200, 180 x 1601 accumulator, 200 peaks)" % (n_scans, points),
           "scans_per_s": n_scans / (ms / n * 1e-3), "ms_per_batch": ms / n, "gpu_launches": n,
           "e2e": {"value": n_scans / e2e_s, "unit": "scans/s", "h2d_bytes_per_step": int(X.nbytes + Y.nbytes + R.nbytes),
                   "d2h_bytes_per_step": int(got["lines"].nbytes + got["n_lines"].nbytes)},
           "mean_lines_per_scan": float(n_lines.mean()),
           "features_e2e": {"value": n_scans / feat_s, "unit": "scans/s", "mean_features_per_scan": float(feat["n_feats"].mean()),
                            "what": "ekf_hough_get_features: Hough lines + fitLineSegments + extractCorners + getStructCompass"},
           # The accumulator (288 KB per scan) never leaves shared memory: HBM traffic is ~5 KB per scan, so an HBM
           # fraction says nothing about this kernel. It is bound on chip - instruction issue (votes, compaction
           # sweeps, the order-dependent 200-peak selection the reference's semantics serialise per scan): the
           # figure reported is the issue-slot utilisation of the ncu capture, not a live measurement.
           "roofline": {"bound": "issue", "kernel": "hough_scan_kernel", "achieved": profile_facts().get("hough_issue_active_pct"),
                        "peak": 100.0, "unit": "% of issue slots (ncu smsp__issue_active, profiles/r01_hough_scan_metrics.csv)",
                        "frac": (profile_facts().get("hough_issue_active_pct") or 0.0) / 100.0 or None, "traffic": None,
                        "hbm_gbs_for_reference": alg / (ms / n * 1e-3) / 1e9,
                        "note": "on-chip bound; HBM carries ~5 KB per scan (%.1f GB/s here, %.2f %% of the copy peak)"
                                % (alg / (ms / n * 1e-3) / 1e9, 100.0 * alg / (ms / n * 1e-3) / 1e9 / hbm_peak)}}
    if with_cpu:
        chk = HoughRef() if HoughRef.available() else HoughOracle()
        cores = os.cpu_count() or 1
        per = 24
        sample = min(n_scans, per * cores)
        chunks = [(i, min(i + per, sample)) for i in range(0, sample, per)]
        chk.run_scans(X[:2], Y[:2], R[:2])
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            list(ex.map(lambda ab: chk.run_scans(X[ab[0]:ab[1]], Y[ab[0]:ab[1]], R[ab[0]:ab[1]]), chunks))
        cpu_s = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": sample / cpu_s, "unit": "scans/s", "cores": cores,
                               "kind": "reference" if HoughRef.available() else "port",
                               "sample": "%d scans, %d per task over %d host threads; features/houghtransform.cpp compiled "
                                         "unmodified (-O2)" % (sample, per, cores)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--filters", type=int, default=65536,
                    help="filters of the whole job, split evenly over the ranks (strong scaling; BASELINE configs[2])")
    ap.add_argument("--filters-per-gpu", type=int, default=0,
                    help="if > 0: this many filters on every GPU instead (weak scaling) and --filters is ignored")
    ap.add_argument("--weak-filters-per-gpu", type=int, default=4096,
                    help="filters per GPU of the weak-scaling side run reported under 'weak' (0 = skip; BASELINE configs[1])")
    ap.add_argument("--large-map", default="2000,10000", help="comma list of landmark counts for the regime-B leg ('' = skip)")
    ap.add_argument("--sharded-map", default="auto",
                    help="landmark count for the multi-GPU sharded single-map leg; 'auto' = 10000 over the job's GPUs "
                         "when --gpus > 1 (run by rank 0 after the batch measurement), skipped on one GPU; '' = skip")
    ap.add_argument("--shard-devices", default="", help="comma list of device ordinals for --sharded-map (default: all visible)")
    ap.add_argument("--hough", type=int, default=4096, help="scans in the Hough front-end leg (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compass-every", type=int, default=0,
                    help="structural-compass update (doUpdateCompass) every k-th step; 0 = none (the headline)")
    ap.add_argument("--meas", type=int, default=1,
                    help="measurements (doUpdate calls) per step; the headline is 1, SURVEY 8d also asks for 4")
    args = ap.parse_args()
    global MAX_MEAS, COMPASS_EVERY
    MAX_MEAS = args.meas
    COMPASS_EVERY = args.compass_every
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
        return

    ekf = load_product()
    rank, world, local, dist = dist_setup(args.gpus)
    strong = args.filters_per_gpu <= 0
    F = args.filters // world if strong else args.filters_per_gpu
    f_base = rank * F
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    # ---- inputs: this rank's filter range, generated once on the host into pinned memory ------------
    syn = ekf.Synth(N_LM, steps_per_lap=T_LAP, max_meas=MAX_MEAS, compass_every=COMPASS_EVERY)
    L = syn.record_len
    pinned = ekf.PinnedArray((F, T_LAP, L))
    syn.generate(F, T_LAP, f0=f_base, out=pinned.array)
    rec = pinned.array
    fb = ekf.FilterBatch(F, CAP_LM, device=local)
    outs = fb.alloc_outputs(T_LAP, MAX_MEAS, trace=True, pose_trace=False, pinned=True)

    # ---- device-resident throughput ("value") -------------------------------------------------------
    fb.upload_records(rec, MAX_MEAS)
    # Capacity = the world's 50 landmarks. In a large population a few filters see a spurious "New"
    # association (an outlier beyond Gamma_max) after their map is complete; the reference would grow
    # its state to 51 landmarks, here the association is dropped, flagged (EKF_ERR_CAPACITY) and
    # counted in config.dropped_new_associations.
    for _ in range(args.warmup):           # lap 0 builds the 50-landmark maps; later laps are full size
        fb.run_resident(trace=True)
    fb.sync(allow_capacity=True)
    fb.capacity_flags(clear=True)          # dropped New associations are counted over every timed lap below
    fb.kernel_time()
    sampler = ClockSampler(local)
    sampler.start()
    barrier(dist, local)
    l0 = fb.kernel_launches()
    fb.timer_start()
    for _ in range(args.steps):
        fb.run_resident(trace=True)
    ms = fb.timer_stop()
    barrier(dist, local)
    l1 = fb.kernel_launches()
    kms, kn = fb.kernel_time()
    ms_max = max_over_ranks(dist, local, ms)
    last = fb.download_outputs(trace=True, outputs=outs, allow_capacity=True)
    assert (last["final_nlm"] >= N_LM).all(), "maps must be complete in the timed laps"
    dec = last["decision"]
    n_grown = int((last["final_nlm"] > FAST_TILES_LM).sum())     # filters finished by the larger-tile instance
    n_old = int((dec == 1).sum())
    n_steps_lap = F * T_LAP
    n_meas = int((dec >= 0).sum())
    n = 3 + 2 * N_LM
    # F_min: one propagate per step, gating per measurement, downdate per Old update
    flops_per_launch = n_steps_lap * (8 * N_LM + 170) + n_meas * 150 * N_LM + n_old * (2 * n * n + 42 * n)
    if COMPASS_EVERY > 0:   # rank-1 symmetric downdate (half the matrix, one fma per pair) + gain / state, per compass step
        flops_per_launch += (n_steps_lap // COMPASS_EVERY) * (n * n + 6 * n)
    value = world * F * T_LAP * args.steps / (ms_max * 1e-3)

    # ---- end-to-end through the C ABI with host buffers ("e2e") ---------------------------------------
    for _ in range(2):
        fb.run(rec, MAX_MEAS, outputs=outs, allow_capacity=True)
    barrier(dist, local)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fb.run(rec, MAX_MEAS, outputs=outs, allow_capacity=True)
    e2e_s = time.perf_counter() - t0
    barrier(dist, local)
    e2e_s = max_over_ranks(dist, local, e2e_s)
    clocks = sampler.stop()
    e2e_val = world * F * T_LAP * args.steps / e2e_s
    # what the host side of this box can deliver: the same pinned records copied host -> device and nothing
    # else, on every rank at the same time (the end-to-end lap cannot be shorter than this copy)
    barrier(dist, local)
    h2d_only_ms = -1.0
    try:                                                # local work only inside the try: the collectives below always run
        import torch
        with torch.cuda.device(local):
            src = torch.from_numpy(rec.reshape(-1))
            dst = torch.empty(src.numel(), dtype=src.dtype, device="cuda:%d" % local)
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize(local)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize(local)
            h2d_only_ms = e0.elapsed_time(e1) / 3
            del dst
    except Exception:                                   # noqa: BLE001  an extra, never fatal
        h2d_only_ms = -1.0
    h2d_only_ms = max_over_ranks(dist, local, h2d_only_ms)
    if h2d_only_ms <= 0:
        h2d_only_ms = None
    h2d = int(rec.nbytes)
    d2h = fb.output_bytes(outs)
    # dropped New associations over ALL timed laps (device-resident and end-to-end) and ALL ranks
    n_dropped = int(sum_over_ranks(dist, local, fb.capacity_flags()))
    n_grown = int(sum_over_ranks(dist, local, n_grown))
    fb.close()
    assert n_dropped == 0, "%d filters dropped a New association at capacity %d" % (n_dropped, CAP_LM)

    # ---- weak-scaling side run (BASELINE configs[1] per GPU) ------------------------------------------
    weak = None
    if strong and args.weak_filters_per_gpu > 0:
        Fw = min(args.weak_filters_per_gpu, F)
        fw = ekf.FilterBatch(Fw, CAP_LM, device=local)
        fw.upload_records(rec[:Fw], MAX_MEAS)       # the first filters of this rank's range
        for _ in range(args.warmup):
            fw.run_resident(trace=True)
        fw.sync(allow_capacity=True)
        barrier(dist, local)
        fw.timer_start()
        for _ in range(args.steps):
            fw.run_resident(trace=True)
        wms = fw.timer_stop()
        barrier(dist, local)
        wms = max_over_ranks(dist, local, wms)
        fw.close()
        weak = {"filters_per_gpu": Fw, "value": world * Fw * T_LAP * args.steps / (wms * 1e-3), "unit": UNIT,
                "ms_per_step": wms / args.steps, "scaling": "weak"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused batch kernel, FP64 pipe) ------------------------------
    fp64_peak = ekf.measure_fp64_peak(local)
    achieved = flops_per_launch / (kms * 1e-3) if kn else None
    prof = profile_facts()
    roofline = {"bound": "fp64", "kernel": "ekf_batch_stile_kernel<13> (+ a continuation launch of <14> for maps beyond 50 landmarks)",
                "achieved": achieved / 1e12 if achieved else None, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if achieved else None,
                "traffic": (prof.get("batch_traffic_bytes_per_filter_lap") or 0) * F or None,
                "algorithmic_flops_per_launch": flops_per_launch, "avg_kernel_ms": kms, "launches_timed": kn,
                "peak_source": "measured live: DFMA-chain microbenchmark (ekf_measure_fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "note": "no dense contraction on this path -> tensor cores unused; graded flops are F_min "
                        "(SURVEY 8d): 150/landmark gated + 2n^2+42n per Old update + 8N+170 per propagate"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "filters_total": world * F, "filters_per_gpu": F, "landmarks": N_LM,
                       "landmark_capacity": CAP_LM, "fast_tile_capacity": FAST_TILES_LM,
                       "steps_per_lap": T_LAP, "measurements_per_step": MAX_MEAS, "compass_every": COMPASS_EVERY,
                       "old_fraction": n_old / max(n_meas, 1), "updates_per_step": n_meas / n_steps_lap,
                       "dropped_new_associations_all_ranks_all_timed_laps": n_dropped,
                       "filters_grown_beyond_50_landmarks_all_ranks": n_grown,
                       "l2": "inputs larger than L2: %.0f MB of step records + %.0f MB of covariance per pass"
                             % (rec.nbytes / 1e6, F * (3 + 2 * CAP_LM) * (4 + 2 * CAP_LM) * 8 / 1e6)},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "h2d_gbs_per_rank": h2d / (e2e_s / args.steps) / 1e9, "d2h_gbs_per_rank": d2h / (e2e_s / args.steps) / 1e9,
                    "h2d_only_ms": h2d_only_ms,
                    "h2d_only_gbs_per_rank": (h2d / (h2d_only_ms * 1e-3) / 1e9) if h2d_only_ms else None,
                    "h2d_only_note": "this rank's records copied host -> device with nothing else running, all ranks at "
                                     "once (max over ranks): the floor the host side of the box sets for an end-to-end lap",
                    "note": "per rank: pinned H2D of the step records and D2H of the decisions, pipelined over chunks of filters "
                            "with the kernel (three streams); the byte counts are this rank's"},
            "gpu_launches": int(l1 - l0), "clocks": clocks, "roofline": roofline}
    if weak is not None:
        line["weak"] = weak

    if world == 1 and args.large_map:
        legs = []
        for tok in args.large_map.split(","):
            n_lm = int(tok)
            steps = 10000 if n_lm <= 2000 else 100     # BASELINE configs[3]: 10,000 update steps
            legs.append(large_map_leg(ekf, n_lm, steps, hbm_peak, local, verify_steps=4 if n_lm <= 2000 else 2))
        line["large_map"] = legs
        line["roofline_hbm"] = dict(legs[-1]["roofline"], peak_source=hbm_src)
    if world == 1 and args.hough > 0:
        try:                                   # an auxiliary leg must never cost the headline line
            line["hough"] = hough_leg(ekf, args.hough, hbm_peak, local, with_cpu=not args.no_cpu_baseline)
        except Exception as e:                 # noqa: BLE001
            line["hough"] = {"error": "%s: %s" % (type(e).__name__, e)}
    sh = args.sharded_map
    if sh == "auto":
        sh = "10000" if world > 1 else ""
    if sh:
        # one process (rank 0) drives every GPU of the job: the covariance of ONE map column-sharded over
        # them, gain rows exchanged by peer stores over NVLink; the other ranks are done
        devs = [int(t) for t in args.shard_devices.split(",") if t.strip()] or list(range(min(max(world, 1), ekf.device_count()) if world > 1 else ekf.device_count()))
        steps = 400 if int(sh) <= 4000 else 60
        try:
            legs = [sharded_map_leg(ekf, int(sh), steps, hbm_peak, devs)]
            if len(devs) > 1:                  # the same map on one GPU, for the speed-up
                legs.append(sharded_map_leg(ekf, int(sh), steps, hbm_peak, devs[:1]))
                legs[0]["speedup_vs_one_gpu"] = legs[1]["ms_per_step"] / legs[0]["ms_per_step"]
            line["sharded_map"] = legs
        except Exception as e:                 # noqa: BLE001  an auxiliary leg must never cost the headline line
            line["sharded_map"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_baseline(rec)
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
