"""One sharded large-map measurement (bench.py's sharded_map_leg) without the rest of the bench:

    python profiles/shard_run.py 10000 0,1 [steps]     -> one JSON line
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

if __name__ == "__main__":
    n_lm = int(sys.argv[1])
    devs = [int(t) for t in sys.argv[2].split(",")]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else (400 if n_lm <= 4000 else 60)
    ekf = bench.load_product()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6560.0)) if isinstance(peaks, dict) else 6560.0
    print(json.dumps(bench.sharded_map_leg(ekf, n_lm, steps, hbm, devs)), flush=True)
