# one equilibrated SJ(Ne) sweep launch of 16384 walkers x 50 sweeps for ncu (-k regex:sj_sweep -s 2 -c 1)
import sys; sys.path.insert(0, '.')
import mole_b200 as m
ctx = m.default_context()
SEED = bytes(32)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 14
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 50
wf = m.SlaterJastrow(5, 5, (9.64, 2.88, 2.88), (0.5, 1.0, 0.0, 0.0), 1.0)
op = m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [10])
ens = m.Ensemble(W, 10, SEED); ens.init_normal(0.5)
ens.sweep(wf, m.MetropolisBox.from_rng(0.5, SEED), op, n_sweeps=100, observables=0)
met = m.MetropolisDiffuse.from_rng(0.02, SEED)
obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
ens.sweep(wf, met, op, n_sweeps=20, block_size=10, observables=0)
ens.sweep(wf, met, op, n_sweeps=NS, block_size=10, observables=obs)
ctx.synchronize()
