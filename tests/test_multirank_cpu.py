"""N>1 host logic on CPU with world_size-2 gloo: contiguous walker sharding keyed by global walker ids,
allreduce of the packed accumulator vector, and the redundant (deterministic) finaliser + optimizer step
on every rank must reproduce the single-process result.  The per-rank "device work" is played by the
oracle here (tests may use it); on GPUs it is mole_sweep + mole_acc_allreduce (tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest

from common import SEED0, cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_moments(name, W, offset, steps, bs, quirk=0):
    import oracle as O
    import mole_b200 as m
    from test_abi import _acc_from_samples
    c = cases()[name]
    cfgs = np.array([O.init_uniform(SEED0, offset + w, c["ne"]) for w in range(W)])
    obs = O.OBS_ENERGY | O.OBS_WFVALUE | O.OBS_PGRAD
    r = O.ensemble_run(c["owf"], c["oham"], O.run_options(O.METROP_DIFFUSE, 0.25, obs, quirk_vector_div=quirk, nan_reject=1),
                       cfgs, SEED0, steps, bs, walker_offset=offset)
    o = r["pgrad"][:, :, :c["np"]] / r["wfvalue"][:, :, None]
    acc = _acc_from_samples(m, r["energy"], o, bs, n_accept=float(r["accept"].sum()), n_moves=float(r["accept"].size))
    return acc, r


def _worker(rank, world, port, name, W, steps, bs, out):
    import torch.distributed as dist
    import mole_b200 as m
    from mole_b200 import distributed as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        n_local, offset = D.shard(W, world, rank)
        acc, _ = _rank_moments(name, n_local, offset, steps, bs)
        tot = D.allreduce_acc(acc)
        e, err, accp, g = m.acc_finalize(tot)
        opt = m.StochasticReconfiguration(0.05, tot.n_params)
        dp = opt.compute_parameter_update(np.array([0.5]), tot)
        out[rank] = dict(n_local=n_local, offset=offset, arr=D.acc_to_array(tot), e=e, err=err, accp=accp, g=g.copy(), dp=dp.copy())
    finally:
        dist.destroy_process_group()


def test_shard_ranges():
    from mole_b200 import distributed as D
    for W, world in ((8, 2), (10, 4), (1 << 20, 8), (7, 3)):
        parts = [D.shard(W, world, r) for r in range(world)]
        assert sum(p[0] for p in parts) == W
        off = 0
        for n, o in parts:
            assert o == off
            off += n


def test_acc_pack_roundtrip():
    import mole_b200 as m
    from mole_b200 import distributed as D
    a = m.ffi.AccHost()
    a.n_params = 3
    a.n_samples, a.sum_e, a.sum_oo[5] = 10.0, -3.5, 7.25
    arr = D.acc_to_array(a)
    assert arr[0] == 10.0 and arr[1] == -3.5 and arr[10 + 8 + 8 + 5] == 7.25
    b = D.array_to_acc(arr, 3)
    assert b.n_samples == 10.0 and b.sum_oo[5] == 7.25 and b.n_params == 3


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_process():
    import torch.multiprocessing as mp
    import mole_b200 as m
    name, W, steps, bs = "h2", 16, 60, 10
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, name, W, steps, bs, out), nprocs=2, join=True)
    acc1, _ = _rank_moments(name, W, 0, steps, bs)
    e1, err1, accp1, g1 = m.acc_finalize(acc1)
    dp1 = m.StochasticReconfiguration(0.05, 1).compute_parameter_update(np.array([0.5]), acc1)
    r0, r1 = out[0], out[1]
    assert (r0["n_local"], r0["offset"], r1["n_local"], r1["offset"]) == (8, 0, 8, 8)
    assert np.array_equal(r0["arr"], r1["arr"])                      # every rank holds the same reduced vector
    assert np.array_equal(r0["dp"], r1["dp"])                        # -> identical redundant solve, no broadcast needed
    assert r0["arr"][0] == W * (steps - bs) and r0["arr"][5] == W * (steps - bs) // bs
    assert abs(r0["e"] - e1) < 1e-12 * abs(e1) and abs(r0["err"] - err1) < 1e-9 * err1
    assert r0["accp"] == accp1
    assert np.allclose(r0["g"], g1, rtol=1e-10, atol=1e-13) and np.allclose(r0["dp"], dp1, rtol=1e-9, atol=1e-13)


def test_rebalance_plan_shares_and_moves():
    """Host arithmetic of mole_rebalance (north_star (5)): shares proportional to the ranks' total weights, summing to
    the global walker count for every shared draw; surplus matched to free slots; unbiased over the draw."""
    import mole_b200 as m
    counts = np.array([1000, 1000, 1000, 1000])
    totals = np.array([1000.0, 3000.0, 500.0, 1500.0])
    acc = np.zeros(4)
    for u in np.linspace(0.0, 0.999, 200):
        shares, moves = m.rebalance_plan(totals, counts, float(u))
        assert shares.sum() == counts.sum() and np.all(shares >= 0)
        assert np.all(np.abs(shares - 4000 * totals / totals.sum()) < 1.0 + 1e-9)
        surplus = shares - counts
        assert np.array_equal(moves.sum(axis=1), np.maximum(surplus, 0))       # everything a rank has too much leaves it
        assert np.array_equal(moves.sum(axis=0), np.maximum(-surplus, 0))      # every free slot is filled
        assert np.all(np.diag(moves) == 0)
        acc += shares
    assert np.allclose(acc / 200, 4000 * totals / totals.sum(), atol=0.51)
    # equal islands: nothing moves; one dead rank: it is refilled entirely by the others
    shares, moves = m.rebalance_plan([2.0, 2.0], [7, 7], 0.3)
    assert list(shares) == [7, 7] and moves.sum() == 0
    shares, moves = m.rebalance_plan([0.0, 5.0, 5.0], [10, 10, 10], 0.5)
    assert shares[0] == 0 and shares.sum() == 30 and moves[:, 0].sum() == 10
    with pytest.raises(m.MoleError):
        m.rebalance_plan([0.0, 0.0], [4, 4], 0.1)
