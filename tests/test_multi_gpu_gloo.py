"""CPU test (not gpu) of the N>1 path's host logic with torch.distributed (gloo, world_size 2):
ranks own disjoint contiguous filter ranges, generate their own records, run independently (the
oracle stands in for the device here), and the only communication is the barrier / max-over-ranks
timing reduction plus one gather at the end - exactly what bench.py does with NCCL."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
WORLD, F_PER_RANK, N, T = 2, 3, 10, 150


def _worker(rank, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    from conftest import load_product
    from oracle_lib import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    ekf = load_product()
    syn = ekf.Synth(N, steps_per_lap=T)
    rec = syn.generate(F_PER_RANK, T, f0=rank * F_PER_RANK)          # this rank's filter range only
    dist.barrier()
    out = Oracle().run_batch(rec, 1, N + 2, pose_trace=True)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)       # stand-in for the rank's elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dec = torch.from_numpy(out["decision"].copy())
    pose = torch.from_numpy(out["final_pose"].copy())
    decs = [torch.empty_like(dec) for _ in range(WORLD)]
    poses = [torch.empty_like(pose) for _ in range(WORLD)]
    dist.all_gather(decs, dec)
    dist.all_gather(poses, pose)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), decision=torch.cat(decs).numpy(),
                 final_pose=torch.cat(poses).numpy(), tmax=t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_equals_single_process(ekf, oracle, tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.start_processes(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True, start_method="spawn")
    g = np.load(tmp_path / "gathered.npz")
    syn = ekf.Synth(N, steps_per_lap=T)
    rec = syn.generate(WORLD * F_PER_RANK, T)
    want = oracle.run_batch(rec, 1, N + 2, pose_trace=True)
    assert np.array_equal(g["decision"], want["decision"])
    assert np.array_equal(g["final_pose"], want["final_pose"])
    assert g["tmax"][0] == WORLD
