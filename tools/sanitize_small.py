# small end-to-end pass over every kernel family for compute-sanitizer
import sys; sys.path.insert(0, '.'); 
import numpy as np, mole_b200 as m
seed = bytes(range(32))
obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
# thread-per-walker kinds
wf = m.HydrogenMoleculeWaveFunction(1.4, [0.5]); op = m.ElectronicHamiltonian.from_ions([[-0.7,0,0],[0.7,0,0]],[1,1])
ens = m.Ensemble(300, 2, seed); ens.init_uniform()
ens.sweep(wf, m.MetropolisDiffuse(0.25, seed), op, n_sweeps=12, block_size=5, observables=obs, traces=("energy","accept"), keep_series=True)
ens.sweep(wf, m.MetropolisBox(1.0, seed), op, n_sweeps=30, block_size=5, observables=m.ffi.OBS_ENERGY, append_series=True)
print("series", ens.series_analyze(block_sizes=[1,2,3], per_walker=True)["tcorr"])
# LCAO determinant kinds (sweeps with SR moments, DMC step with the nodal triplet)
bs = m.Hydrogen1sBasis([[-0.7, 0, 0], [0.7, 0, 0]], [0.85])
tri = m.SingleDeterminant([m.Orbital([[1.0], [1.0]], bs), m.Orbital([[1.0], [-1.0]], bs)])
el = m.Ensemble(333, 2, seed); el.init_uniform()
el.sweep(tri, m.MetropolisDiffuse(0.1, seed), op, n_sweeps=12, block_size=5, observables=obs, traces=("pgrad",))
el.dmc_step(tri, m.MetropolisDiffuse(0.02, seed), op, 0.02, -0.5); el.branch(m.ffi.BRANCH_SR)
sg = m.SingleDeterminant([m.Orbital([[1.0], [1.0]], m.Hydrogen1sBasis([[-1.25, 0, 0], [1.25, 0, 0]], [1.0]))])
e1 = m.Ensemble(77, 1, seed); e1.init_uniform()
e1.sweep(sg, m.MetropolisBox(1.0, seed), op, n_sweeps=12, block_size=5, observables=obs)
print("lcao acc", el.acc_get().sum_e, e1.acc_get().sum_e)
# Slater-Jastrow
sj = m.SlaterJastrow(5, 5, (9.64, 2.88, 2.88), (0.5, 1.0, 0.1, -0.05), 1.0); opn = m.ElectronicHamiltonian.from_ions([[0,0,0]],[10])
e2 = m.Ensemble(50, 10, seed); e2.init_normal(0.5)
e2.sweep(sj, m.MetropolisBox(0.4, seed), opn, n_sweeps=10, observables=0)
e2.sweep(sj, m.MetropolisDiffuse(0.02, seed), opn, n_sweeps=20, block_size=5, observables=obs)
print("sj acc", e2.acc_get().sum_e)
sjl = m.SlaterJastrow(2, 1, (2.69, 0.64, 0.64), (0.4, 0.8, 0.0, 0.0), 1.5); opl = m.ElectronicHamiltonian.from_ions([[0,0,0]],[3])
e3 = m.Ensemble(31, 3, seed); e3.init_normal(0.8)
e3.sweep(sjl, m.MetropolisDiffuse(0.02, seed), opl, n_sweeps=20, block_size=5, observables=obs)
# DMC: step, both branchers, block loop, SJ DMC
st = m.STO(0.9); oph = m.ElectronicHamiltonian.from_ions([[0,0,0]],[1]); met = m.MetropolisDiffuse.from_rng(0.025, seed)
d = m.DmcRunner.new(st, 3000, -0.45, oph, met, m.SRBrancher.new(), identical_start=False).ensemble
print("block", d.dmc_block(st, met, oph, m.ffi.BRANCH_SR, 0.025, -0.47, 5)[-1])
d.dmc_step(st, met, oph, 0.025, -0.47); d.branch(m.ffi.BRANCH_SIMPLE)
sjb = m.SlaterJastrow(2, 2, (3.68, 0.96, 0.96), (0.5, 1.0, 0.2, 0.1), 1.0); opb = m.ElectronicHamiltonian.from_ions([[0,0,0]],[4])
db = m.DmcRunner.new(sjb, 100, -14.6, opb, m.MetropolisDiffuse.from_rng(0.01, seed), m.SRBrancher.new(), identical_start=False).ensemble
print("sj dmc", db.dmc_block(sjb, m.MetropolisDiffuse.from_rng(0.01, seed), opb, m.ffi.BRANCH_SR, 0.01, -14.6, 3)[-1])
# second round: general LCAO Slater-Jastrow kind (rows + Gram on both pipes, DMC), persistent DMC block kernel against the
# per-step launches, single-rank rebalancing, the probes
h4 = [[-2.1, 0, 0], [-0.7, 0, 0], [0.7, 0, 0], [2.1, 0, 0]]
lw = m.LcaoSlaterJastrow(2, 2, h4, [1.0, 1 / 1.1, 1 / 1.1, 1.0], [[1, 1, 1, 1], [1, 0.5, -0.5, -1]], [0.5, 1.0, 0.1, -0.05])
lop = m.ElectronicHamiltonian.from_ions(h4, [1, 1, 1, 1])
for impl in (0, 1):
    le = m.Ensemble(130, 4, seed); le.init_normal(1.5); le.gram_select(impl)
    le.sweep(lw, m.MetropolisDiffuse(0.05, seed), lop, n_sweeps=12, n_discard=4, block_size=4, observables=obs)
    g = le.gram_get()
    print("lsj gram", impl, g.shape, float(g[0][0]), float(g[0][1]))
le.dmc_step(lw, m.MetropolisDiffuse(0.02, seed), lop, 0.02, -2.0); le.branch(m.ffi.BRANCH_SR)
for W in (300, 5000):
    out = []
    for impl in (0, 1):
        de = m.Ensemble(W, 1, seed); de.init_normal(1.0); de.dmc_block_select(impl)
        out.append(de.dmc_block(st, met, oph, m.ffi.BRANCH_SR, 0.025, -0.47, 7))
    assert (out[0] == out[1]).all()
    print("dmc block fused == per-step", W, out[0][-1])
de.rebalance()
print("probes", m.default_context().bench_gram(4096, 20, 38, 0, 1)[0] > 0, m.default_context().dmma_peak_tflops(2, 4) > 0)
m.default_context().synchronize()
print("done")
