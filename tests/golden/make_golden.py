"""Generate tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libekf_ref.so: the
reference's odometry/*.cpp compiled unmodified over oracle/shim). Run in the build container,
where /root/reference exists:

    python tests/golden/make_golden.py

Each file stores the exact input step records and what the reference produced for them:
decisions, landmark indices, Mahalanobis distances, the pose after every step, and the final
state and covariance. The reference ships no golden vectors of its own, so these
are the pinned vectors: the oracle (tests/test_golden.py, CPU) and the CUDA core
(tests/test_gpu_golden.py, GPU) are both checked against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import load_product  # noqa: E402
from oracle_lib import Ref  # noqa: E402

CASES = {
    # BASELINE configs[0]: single filter, 20 landmarks, 1,000 propagate/update steps
    "config1_n20_t1000": dict(N=20, F=1, T=1000, laps=1, M=1, compass_every=0, cap=24),
    # second lap over a complete map + several measurements per step + structural compass
    "n12_m2_compass": dict(N=12, F=2, T=300, laps=2, M=2, compass_every=5, cap=16),
}


def main():
    ekf = load_product()
    ref = Ref()
    for name, c in CASES.items():
        syn = ekf.Synth(c["N"], steps_per_lap=c["T"], max_meas=c["M"], compass_every=c["compass_every"])
        lap = syn.generate(c["F"], c["T"])
        rec = np.ascontiguousarray(np.concatenate([lap] * c["laps"], axis=1))
        out = ref.run_batch(rec, c["M"], c["cap"], pose_trace=True, final_state=True)
        assert not out["bad"]
        nl = out["final_nlm"]
        n = 3 + 2 * int(nl.max())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), records=rec, max_meas=c["M"], cap=c["cap"],
                            decision=out["decision"], index=out["index"], mahal=out["mahal"],
                            pose_trace=out["pose_trace"], final_nlm=nl,
                            final_x=out["final_x"][:, :n], final_P=out["final_P"][:, :n, :n])
        d = out["decision"]
        print("%s: New %d Old %d Ignore %d, landmarks %s" % (name, (d == 0).sum(), (d == 1).sum(), (d == 2).sum(), nl))


if __name__ == "__main__":
    main()
