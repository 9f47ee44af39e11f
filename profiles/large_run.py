"""One single-GPU large-map measurement (bench.py's large_map_leg) without the rest of the bench:

    python profiles/large_run.py 2000 [steps]     -> one JSON line
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
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else (400 if n_lm <= 4000 else 30)
    ekf = bench.load_product()
    print(json.dumps(bench.large_map_leg(ekf, n_lm, steps, 6560.0, 0)), flush=True)
