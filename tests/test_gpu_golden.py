"""GPU test: the CUDA core against the committed golden vectors from the reference build."""
import numpy as np
import pytest

from test_golden import GOLDEN, check_against_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", [1, 2, 3, 4], ids=["smem", "tile", "stile", "dtile"])
@pytest.mark.parametrize("path", GOLDEN, ids=[p.split("/")[-1][:-4] for p in GOLDEN])
def test_cuda_matches_golden(ekf, path, kernel):
    g = np.load(path)
    rec = np.ascontiguousarray(g["records"])
    fb = ekf.FilterBatch(rec.shape[0], int(g["cap"]), batch_kernel=kernel)
    got = fb.run(rec, int(g["max_meas"]), trace=True, pose_trace=True)
    check_against_golden(g, got, lambda f, n: fb.get_state(f), 1e-9)
    fb.close()
