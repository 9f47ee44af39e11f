"""One Hough front-end measurement (bench.py's hough_leg) without the rest of the bench:

    python profiles/hough_run.py [n_scans] [cpu]     -> one JSON line
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    ekf = bench.load_product()
    print(json.dumps(bench.hough_leg(ekf, n, 6560.0, 0, with_cpu=len(sys.argv) > 2)), flush=True)
