# one diffuse sweep launch with SR moments of a thread-per-walker kind for ncu (-k regex:sweep_kernel -s 1 -c 1):
#   python tools/prof_h2.py            H2 Heitler-London STO (the default)
#   python tools/prof_h2.py lcao_he    any case name of tests/common.py (he, lcao_h2p, lcao_h2_singlet, ...)
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import mole_b200 as m
from common import cases
name = sys.argv[1] if len(sys.argv) > 1 else 'h2'
c = cases()[name]
ctx = m.default_context()
seed = bytes(32)
W = 1 << 20
wf, op = c['make'](m)
ens = m.Ensemble(W, c['ne'], seed); ens.init_uniform()
met = m.MetropolisDiffuse(0.25 if name == 'h2' else 0.1, seed)
obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
ens.sweep(wf, met, op, n_sweeps=20, block_size=10, observables=obs)
ens.sweep(wf, met, op, n_sweeps=200, block_size=10, observables=obs)
ctx.synchronize()
