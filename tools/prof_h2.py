# one H2 (Heitler-London STO) diffuse sweep launch with SR moments for ncu (-k regex:sweep_kernel -s 1 -c 1)
import sys; sys.path.insert(0, '.')
import mole_b200 as m
ctx = m.default_context()
seed = bytes(32)
W = 1 << 20
wf = m.HydrogenMoleculeWaveFunction(1.4, [0.5]); op = m.ElectronicHamiltonian.from_ions([[-0.7, 0, 0], [0.7, 0, 0]], [1, 1])
ens = m.Ensemble(W, 2, seed); ens.init_uniform()
met = m.MetropolisDiffuse(0.25, seed)
obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
ens.sweep(wf, met, op, n_sweeps=20, block_size=10, observables=obs)
ens.sweep(wf, met, op, n_sweeps=200, block_size=10, observables=obs)
ctx.synchronize()
