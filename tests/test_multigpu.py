"""Multi-rank DEVICE paths: mole_acc_allreduce and mole_gram_allreduce over NCCL, the multi-rank
mole_dmc_block (population islands, one all-gather per block) and mole_rebalance on two GPUs, against single-rank runs of
the same global walker ids.
Needs >= 2 visible GPUs (gpurun --gpus 2); skipped otherwise.  One process per GPU, like the bench."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = bytes([5] * 32)


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def _vmc_setup(m, ctx):
    wf = m.SlaterJastrow(2, 2, (3.68, 0.96, 0.96), (0.5, 1.0, 0.2, 0.1), 1.0, ctx=ctx)
    op = m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [4], ctx=ctx)
    return wf, op, m.MetropolisDiffuse.from_rng(0.05, SEED)


def _dmc_setup(m, ctx):
    wf = m.STO(0.9, ctx=ctx)
    op = m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1], ctx=ctx)
    return wf, op, m.MetropolisDiffuse.from_rng(0.025, SEED)


def _lsj_setup(m, ctx):
    h4 = [[-2.1, 0, 0], [-0.7, 0, 0], [0.7, 0, 0], [2.1, 0, 0]]
    wf = m.LcaoSlaterJastrow(2, 2, h4, [1.0, 1 / 1.1, 1 / 1.1, 1.0], [[1, 1, 1, 1], [1, 0.5, -0.5, -1]], [0.5, 1.0, 0.1, -0.05], ctx=ctx)
    return wf, m.ElectronicHamiltonian.from_ions(h4, [1, 1, 1, 1], ctx=ctx)


def _rank(rank, world, uid, W_local, out):
    sys.path.insert(0, ROOT)
    import torch
    import mole_b200 as m
    from mole_b200 import distributed as D
    torch.cuda.set_device(rank)
    ctx = m.Context(rank)
    ctx.comm_init(world, rank, uid)
    obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
    # ---- VMC: sharded sweep + NCCL allreduce of the packed moments
    wf, op, met = _vmc_setup(m, ctx)
    ens = m.Ensemble(W_local, 4, SEED, walker_offset=rank * W_local, ctx=ctx)
    ens.init_uniform(-1.0, 1.0)
    ens.sweep(wf, met, op, n_sweeps=40, n_discard=10, block_size=10, observables=obs)
    local = D.acc_to_array(ens.acc_get())
    ens.acc_allreduce()
    tot = ens.acc_get()
    res = dict(vmc_local=local, vmc_total=D.acc_to_array(tot), vmc_cfgs=ens.get_configs(), vmc_health=ens.health())
    # ---- DMC: the multi-rank block (islands + one gather) ...
    dwf, dop, dmet = _dmc_setup(m, ctx)
    a = m.Ensemble(W_local, 1, SEED, walker_offset=rank * W_local, ctx=ctx)
    a.init_normal(1.0)
    res["dmc_block_energies"] = a.dmc_block(dwf, dmet, dop, m.ffi.BRANCH_SR, 0.025, -0.5, 8)
    res["dmc_block_cfgs"] = a.get_configs()
    res["dmc_block_weights"] = a.get_weights()
    # ... against the step-by-step entry points on a communicator-free context of the same GPU
    ctx1 = m.Context(rank)
    swf, sop, smet = _dmc_setup(m, ctx1)
    b = m.Ensemble(W_local, 1, SEED, walker_offset=rank * W_local, ctx=ctx1)
    b.init_normal(1.0)
    rows = []
    for _ in range(8):
        rows.append(b.dmc_step(swf, smet, sop, 0.025, -0.5))
        b.branch(m.ffi.BRANCH_SR)
    res["dmc_rows"] = np.array(rows)
    res["dmc_step_cfgs"] = b.get_configs()
    # ---- population rebalancing: rank 0 carries weight 1 per walker, rank 1 weight 3; walkers are tagged by x of electron 0
    r = m.Ensemble(W_local, 1, SEED, walker_offset=rank * W_local, ctx=ctx)
    tag = np.zeros((W_local, 1, 3))
    tag[:, 0, 0] = rank * W_local + np.arange(W_local)
    r.set_configs(tag)
    r.set_weights(np.full(W_local, 1.0 + 2.0 * rank))
    r.rebalance()
    res["reb_cfgs"] = r.get_configs()
    res["reb_weights"] = r.get_weights()
    del r
    # ---- island imbalance after a block, and the rebalancing mole_dmc_diffuse triggers on it
    i1 = a.island_imbalance()
    a.set_weights(np.full(W_local, 1.0 + 2.0 * rank))
    a.dmc_block(dwf, dmet, dop, m.ffi.BRANCH_SR, 0.025, -0.5, 1)
    i2 = a.island_imbalance()
    a.rebalance()
    a.dmc_block(dwf, dmet, dop, m.ffi.BRANCH_SR, 0.025, -0.5, 2)
    res["imbalance"] = (i1, i2, a.island_imbalance())
    # the driver itself: DmcRunner::diffuse over both ranks, starting from islands that differ 3 : 1 in weight per walker
    run = m.DmcRunner.new(dwf, W_local, -0.5, dop, dmet, m.SRBrancher.new(), identical_start=False, walker_offset=rank * W_local)
    run.ensemble.set_weights(np.full(W_local, 1.0 + 2.0 * rank))
    en, er = run.diffuse(0.025, 60, 10, 1)
    res["diffuse"] = (en, er, run.ensemble.island_imbalance(), run.ensemble.get_weights())
    # ---- large-P path: sample rows on each shard, Gram matrices summed over NCCL
    lwf, lop = _lsj_setup(m, ctx)
    g = m.Ensemble(W_local // 2, 4, SEED, walker_offset=rank * (W_local // 2), ctx=ctx)
    g.init_normal(1.5)
    g.sweep(lwf, m.MetropolisDiffuse.from_rng(0.05, SEED), lop, n_sweeps=16, n_discard=4, block_size=4, observables=obs)
    res["gram_local"] = g.gram_get()
    g.gram_allreduce()
    res["gram_total"] = g.gram_get()
    del g
    out[rank] = res
    del ens, a, b
    ctx.close()


@pytest.mark.timeout(600)
def test_two_gpu_allreduce_and_dmc_block():
    if not _two_gpus():
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import mole_b200 as m
    from mole_b200 import distributed as D
    W_local = 1000
    uid = m.comm_unique_id()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank, args=(2, uid, W_local, out), nprocs=2, join=True)
    r0, r1 = out[0], out[1]
    # VMC: every rank holds the same reduced vector = the sum of the local ones; health counters ride along
    assert np.array_equal(r0["vmc_total"], r1["vmc_total"])
    assert np.allclose(r0["vmc_total"], r0["vmc_local"] + r1["vmc_local"], rtol=1e-14, atol=0)
    assert r0["vmc_health"] == (0, 0)
    # ... and equals the single-rank run over the same global walker ids (configurations bit for bit)
    ctx = m.default_context()
    wf, op, met = _vmc_setup(m, ctx)
    ens = m.Ensemble(2 * W_local, 4, SEED)
    ens.init_uniform(-1.0, 1.0)
    obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
    ens.sweep(wf, met, op, n_sweeps=40, n_discard=10, block_size=10, observables=obs)
    assert np.array_equal(ens.get_configs(), np.concatenate([r0["vmc_cfgs"], r1["vmc_cfgs"]]))
    one = D.acc_to_array(ens.acc_get())
    scale = np.maximum(np.abs(one), 1e-300)
    assert np.max(np.abs(one - r0["vmc_total"]) / scale) < 1e-11
    # DMC: each rank's island follows the step-by-step entry points bit for bit ...
    for r in (r0, r1):
        assert np.array_equal(r["dmc_block_cfgs"], r["dmc_step_cfgs"])
        assert np.all(r["dmc_block_weights"] == r["dmc_block_weights"][0])
    # ... and every rank forms the same ensemble energies: sum over ranks of {sum w E, sum w}, rank order
    assert np.array_equal(r0["dmc_block_energies"], r1["dmc_block_energies"])
    swe = r0["dmc_rows"][:, 0] + r1["dmc_rows"][:, 0]
    sw = r0["dmc_rows"][:, 1] + r1["dmc_rows"][:, 1]
    assert np.array_equal(r0["dmc_block_energies"], swe / sw)
    # island weights per walker: equal within 2 % after a block from equal weights, 3 : 1 after the hand-set weights,
    # equal again after mole_rebalance; both ranks see the same numbers
    assert r0["imbalance"] == r1["imbalance"]
    i1, i2, i3 = r0["imbalance"]
    assert 1.0 <= i1 < 1.03 and 2.8 < i2 < 3.2 and 1.0 <= i3 < 1.03, r0["imbalance"]
    # DmcRunner::diffuse: the same energies on both ranks, near -0.5 for the STO guide; the 3 : 1 islands were merged
    # after the first block (equal weights per walker on both ranks at the end)
    assert np.array_equal(r0["diffuse"][0], r1["diffuse"][0]) and np.array_equal(r0["diffuse"][1], r1["diffuse"][1])
    assert np.all(np.abs(r0["diffuse"][0] + 0.5) < 0.05) and r0["diffuse"][2] < 1.05
    assert abs(r0["diffuse"][3][0] / r1["diffuse"][3][0] - 1.0) < 0.05
    # large-P path: every rank holds the same summed Gram matrix = the single-rank contraction over the same global walkers
    assert np.array_equal(r0["gram_total"], r1["gram_total"])
    assert np.allclose(r0["gram_total"], r0["gram_local"] + r1["gram_local"], rtol=1e-14, atol=0)
    lwf, lop = _lsj_setup(m, ctx)
    g = m.Ensemble(W_local, 4, SEED)
    g.init_normal(1.5)
    g.sweep(lwf, m.MetropolisDiffuse.from_rng(0.05, SEED), lop, n_sweeps=16, n_discard=4, block_size=4, observables=obs)
    one = g.gram_get()
    assert one.shape == (14, 14) and one[0, 0] == W_local * 12
    assert np.max(np.abs(np.triu(one - r0["gram_total"]))) < 1e-11 * np.max(np.abs(one))
    # rebalancing: total weight conserved, equal weights everywhere, rank 1's walkers fill 3/4 of all slots
    for r in (r0, r1):
        assert np.all(r["reb_weights"] == 2.0) and r["reb_cfgs"].shape[0] == W_local
    tags = np.concatenate([r0["reb_cfgs"][:, 0, 0], r1["reb_cfgs"][:, 0, 0]]).astype(np.int64)
    from0, from1 = (tags < W_local).sum(), (tags >= W_local).sum()
    assert abs(from0 - W_local // 2) <= 1 and from0 + from1 == 2 * W_local
    mult = np.bincount(tags, minlength=2 * W_local)
    assert mult[:W_local].max() <= 1 and 1 <= mult[W_local:].min() and mult[W_local:].max() <= 2   # systematic: copies within 1 of the expectation
    assert np.all(r1["reb_cfgs"][:, 0, 0] >= W_local)                  # the heavy rank keeps its own walkers ...
    assert (r0["reb_cfgs"][:, 0, 0] >= W_local).sum() == W_local - from0  # ... and its surplus fills rank 0's free slots
