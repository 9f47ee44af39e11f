"""GPU parity tests of the column-sharded large map (`ekf_sharded_*`): same
oracle, same tolerances as the single-GPU regimes; results must not depend on the shard count.
Shards may share a device, so the whole exchange logic runs on a single-GPU box; the tests that
need real peers skip unless the box has >= 2 GPUs (run them with `gpurun --gpus 2`)."""
import numpy as np
import pytest

from parity import TOL, assert_state_close, assert_trace_equal, injected_state, rel_state

pytestmark = pytest.mark.gpu


def _device_lists(ekf):
    n = ekf.device_count()
    lists = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [[0, 1], [0, 1, 0, 1]]
    if n >= 4:
        lists.append([0, 1, 2, 3])
    if n >= 8:
        lists.append(list(range(8)))
    return lists


def _oracle_run(oracle, rec, M, cap):
    want = oracle.run_batch(rec, M, cap, pose_trace=True, final_state=True)
    assert not want["bad"]
    n = 3 + 2 * int(want["final_nlm"][0])
    return want, want["final_x"][0, :n].copy(), want["final_P"][0, :n, :n].T.copy()


def test_sharded_run_matches_oracle_for_every_shard_count(ekf, oracle):
    """Map built from empty (New / Old / Ignore, compass, two measurements per step), compared
    step by step; the final state is bit-identical across shard counts."""
    N, T, cap, M = 24, 400, 28, 2
    syn = ekf.Synth(N, steps_per_lap=T // 2, max_meas=M, compass_every=9)
    lap = syn.generate(1, T // 2)
    rec = np.ascontiguousarray(np.concatenate([lap, lap], axis=1))
    want, xr, Pr = _oracle_run(oracle, rec, M, cap)
    first = None
    for devs in _device_lists(ekf):
        sm = ekf.ShardedMap(devs, cap)
        got = sm.run(rec, M, trace=True, pose_trace=True)
        what = "shards on devices %s" % devs
        assert_trace_equal(got, want, what)
        assert np.array_equal(got["final_nlm"], want["final_nlm"]) and got["final_nlm"][0] == N
        assert rel_state(got["pose_trace"], want["pose_trace"]) <= TOL
        x, P = sm.get_state()
        assert_state_close(x, P, xr, Pr, what)
        assert np.array_equal(P, P.T), "covariance must stay bit-symmetric across shards"
        for s in range(len(devs)):                      # replicas agree bit for bit with the owners' data
            nl, xs, prr = sm.get_replica(s)
            assert nl == N and np.array_equal(xs, x) and np.array_equal(prr, P[:3, :3])
        if first is None:
            first = (x, P, got)
        else:
            assert np.array_equal(x, first[0]) and np.array_equal(P, first[1]), what + ": depends on shard count"
            assert np.array_equal(got["mahal"], first[2]["mahal"])
        sm.close()


def test_sharded_equals_single_gpu_large_regime_bitwise(ekf):
    """Same kernels' arithmetic as regime B on one GPU: identical bits."""
    N, T, cap, M = 20, 300, 22, 1
    syn = ekf.Synth(N, steps_per_lap=T, max_meas=M, compass_every=11)
    rec = syn.generate(1, T)
    fb = ekf.FilterBatch(1, cap, regime=2)
    a = fb.run(rec, M, trace=True, pose_trace=True)
    xa, Pa = fb.get_state(0)
    sm = ekf.ShardedMap([0, 0, 0], cap)
    b = sm.run(rec, M, trace=True, pose_trace=True)
    xb, Pb = sm.get_state()
    for k in ("decision", "index", "mahal", "pose_trace", "final_pose", "final_nlm"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(xa, xb) and np.array_equal(Pa, Pb)
    fb.close()
    sm.close()


def test_sharded_lookahead_run_is_bit_identical_to_the_exchange_per_gating_chain(ekf, monkeypatch):
    """ekf_sharded_run() overlaps gating / decision with the previous sweep through a replicated O(n) cache
    (no gating exchange, compass without any exchange, second stream per shard). Same bits as the chain with
    an exchange per gating pass (EKF_SHARD_LOOKAHEAD=0), on every shard layout, for runs cut into several
    calls (cache loaded and written back each time) and for per-call operations that follow a run."""
    N, T, cap, M = 26, 300, 30, 2
    syn = ekf.Synth(N, steps_per_lap=T, max_meas=M, compass_every=7)
    lap = syn.generate(1, T)
    rec = np.ascontiguousarray(np.concatenate([lap, lap], axis=1))
    ref = None
    for devs in _device_lists(ekf):
        for la in ("0", "1"):
            monkeypatch.setenv("EKF_SHARD_LOOKAHEAD", la)
            sm = ekf.ShardedMap(devs, cap)
            # flags need one GPU per shard: device lists with repeats run the event chain whatever the switch says
            assert ("look-ahead" in sm.run_mode()) == (la == "1" and len(set(devs)) == len(devs))
            parts, lo = [], 0
            for hi in (1, 10, 250, 2 * T):
                parts.append(sm.run(np.ascontiguousarray(rec[:, lo:hi]), M, trace=True, pose_trace=True))
                lo = hi
            got = {k: np.concatenate([q[k] for q in parts], axis=1) for k in ("decision", "index", "mahal", "pose_trace")}
            r = rec[0, 5]
            sm.propagate(r[0], r[1], r[2])
            sm.update_compass(r[3] + 0.01, 0.02)
            dec, idx, mah = sm.update(r[8:10].reshape(1, 2), r[10:14].reshape(1, 4))
            x, P = sm.get_state()
            reps = [sm.get_replica(s) for s in range(len(devs))]
            sm.close()
            assert parts[-1]["final_nlm"][0] == N
            for nl, xs, prr in reps:
                assert 3 + 2 * nl == len(x) and np.array_equal(xs, x) and np.array_equal(prr, P[:3, :3])
            assert np.array_equal(P, P.T)
            cur = (got, x, P, dec, idx, mah)
            if ref is None:
                ref = cur
                continue
            what = "devices %s, look-ahead %s" % (devs, la)
            for k in got:
                assert np.array_equal(got[k], ref[0][k]), what + ": " + k
            for u, v in zip(cur[1:], ref[1:]):
                assert np.array_equal(u, v), what


def _assert_trace_equal_but_dropped(got, want, what):
    """assert_trace_equal, except that the Mahalanobis distance of a dropped association is not compared: the
    reference has no such case (its map grows without bound) and the oracle harness reports 0 there."""
    keep = want["decision"] != 3
    assert np.array_equal(got["decision"], want["decision"]) and np.array_equal(got["index"], want["index"]), what
    m, mr = got["mahal"][keep], want["mahal"][keep]
    assert (np.abs(m - mr) <= TOL * np.maximum(1.0, np.abs(mr))).all(), what + ": Mahalanobis distances differ"


def test_sharded_run_reports_capacity_in_both_run_modes(ekf, oracle, monkeypatch):
    """A run whose map wants more landmarks than the handle holds: the New associations that do not fit are
    dropped (decision 3, index -1), the call returns EKF_ERR_CAPACITY, everything else goes on - the same
    trace and the same state from the look-ahead run, the event chain and the oracle."""
    N, T, cap, M = 14, 160, 11, 2
    syn = ekf.Synth(N, steps_per_lap=T, max_meas=M, compass_every=6)
    rec = syn.generate(1, T)
    want = oracle.run_batch(rec, M, cap, pose_trace=True, final_state=True)
    assert (want["decision"] == 3).any() and int(want["final_nlm"][0]) == cap
    n = 3 + 2 * cap
    res = []
    for devs in _device_lists(ekf):
        for la in ("1", "0"):
            monkeypatch.setenv("EKF_SHARD_LOOKAHEAD", la)
            sm = ekf.ShardedMap(devs, cap)
            got = sm.run(rec, M, trace=True, pose_trace=True, allow_capacity=True)
            x, P = sm.get_state()
            sm.close()
            what = "devices %s, look-ahead %s" % (devs, la)
            _assert_trace_equal_but_dropped(got, want, what)
            assert rel_state(got["pose_trace"], want["pose_trace"]) <= TOL
            assert_state_close(x, P, want["final_x"][0, :n], want["final_P"][0, :n, :n].T, what)
            res.append((got, x, P))
    for got, x, P in res[1:]:
        for k in ("decision", "index", "mahal", "pose_trace"):
            assert np.array_equal(got[k], res[0][0][k]), k
        assert np.array_equal(x, res[0][1]) and np.array_equal(P, res[0][2])


def test_sharded_percall_surface(ekf, oracle):
    """doPropagation / doUpdateCompass / doUpdate(n_z = 3, gating bound frozen at call entry) one
    call at a time on 3 shards."""
    N, T, cap = 10, 90, 12
    syn = ekf.Synth(N, steps_per_lap=T, max_meas=3, compass_every=5)
    rec = syn.generate(1, T)[0]
    sm = ekf.ShardedMap([0, 0, 0], cap)
    of = oracle.new_filter(cap)
    for t in range(T):
        r = rec[t]
        sm.propagate(r[0], r[1], r[2])
        of.propagate(r[0], r[1], r[2])
        if r[6] != 0:
            sm.update_compass(r[3], r[4])
            of.update_compass(r[3], r[4])
        nz = int(r[5])
        if nz:
            zr = r[8:8 + 6 * nz].reshape(nz, 6)
            dec, idx, mah = sm.update(zr[:, :2], zr[:, 2:])
            n_run = of.n
            for m, tr in enumerate(of.update_chunk(zr[:, :2], zr[:, 2:])):
                assert dec[m] == tr.decision
                assert idx[m] == (n_run if tr.decision == 0 else tr.opt_i)
                assert abs(mah[m] - tr.mahal) <= TOL * max(1.0, abs(tr.mahal))
                n_run += 2 * (tr.decision == 0)
        if t % 10 == 0 or t == T - 1:
            x, P = sm.get_state()
            xr, Pr = of.get_state()
            assert_state_close(x, P, xr, Pr, "step %d" % t)
    pose, nlm = sm.get_pose()
    assert nlm == of.num_landmarks == N
    sm.close()


def test_sharded_same_corner_twice_in_one_call(ekf, oracle):
    z, R = np.array([2.0, 1.0]), np.array([0.01, 0.0, 0.0, 0.02])
    sm = ekf.ShardedMap([0, 0], 6)
    dec, idx, _ = sm.update([z, z, z + 0.01], [R, R, R])
    assert list(dec) == [0, 0, 0] and list(idx) == [3, 5, 7]
    dec2, idx2, _ = sm.update([z, z, z + 0.01], [R, R, R])
    of = oracle.new_filter(6)
    of.update_chunk([z, z, z + 0.01], [R, R, R])
    trs = of.update_chunk([z, z, z + 0.01], [R, R, R])
    assert list(dec2) == [t.decision for t in trs] and list(idx2) == [t.opt_i for t in trs]
    x, P = sm.get_state()
    xr, Pr = of.get_state()
    assert_state_close(x, P, xr, Pr, "chunk")
    sm.close()


def test_sharded_capacity_is_reported(ekf):
    sm = ekf.ShardedMap([0, 0], 2)
    R = np.array([0.01, 0.0, 0.0, 0.02])
    sm.update([[2.0, 1.0]], [R])
    sm.update([[-3.0, 4.0]], [R])
    with pytest.raises(ekf.EkfError) as e:
        sm.update([[6.0, -5.0]], [R])
    assert e.value.code == ekf.ERR_CAPACITY
    assert sm.get_pose()[1] == 2
    sm.close()


@pytest.mark.parametrize("N,steps", [(300, 30), (2000, 4)])
def test_sharded_large_map_injected_state(ekf, oracle, N, steps):
    """BASELINE config 4 shape on every available shard layout: injected state, Old updates."""
    syn = ekf.Synth(N, steps_per_lap=20000, max_meas=1)
    rec = syn.generate(1, steps)
    x0, P0 = injected_state(syn.world(), seed=N)
    of = oracle.new_filter(N + 2).set_state(x0, P0)
    want_dec = []
    for t in range(steps):
        r = rec[0, t]
        of.propagate(r[0], r[1], r[2])
        want_dec.append(of.update(r[8:10], r[10:14]).decision)
    xr, Pr = of.get_state()
    assert sum(d == 1 for d in want_dec) >= steps // 2
    for devs in _device_lists(ekf)[1:]:
        sm = ekf.ShardedMap(devs, N + 2)
        sm.set_state(x0, P0, symmetric=True)
        got = sm.run(rec, 1, trace=True)
        assert list(got["decision"][0, :, 0]) == want_dec
        x, P = sm.get_state()
        assert_state_close(x, P, xr, Pr, "N=%d on %s" % (N, devs))
        assert np.array_equal(P, P.T)
        sm.close()


def test_sharded_needs_distinct_shards_and_valid_devices(ekf):
    with pytest.raises(ekf.EkfError):
        ekf.ShardedMap([0, 0, 0], 2)          # fewer landmarks than shards
    with pytest.raises(ekf.EkfError) as e:
        ekf.ShardedMap([0, 99], 10)
    assert e.value.code == ekf.ERR_NO_DEVICE
