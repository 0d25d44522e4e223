# DMC block loop (mole_dmc_block, SRBrancher) of W walkers: wall-clock per time step, or under
#   ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv python tools/prof_dmc.py 32768 30
# the per-kernel durations of the three launches of a step.
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mole_b200 as m
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
seed = bytes([1] * 32)
st = m.STO(0.9); op = m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1]); met = m.MetropolisDiffuse.from_rng(0.025, seed)
ens = m.DmcRunner.new(st, W, -0.45, op, met, m.SRBrancher.new(), identical_start=False).ensemble
ens.dmc_block(st, met, op, m.ffi.BRANCH_SR, 0.025, -0.47, 20)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    e = ens.dmc_block(st, met, op, m.ffi.BRANCH_SR, 0.025, -0.47, steps)
    ts.append(1e6 * (time.perf_counter() - t0) / steps)
print("dmc_block W=%d steps=%d: us/step %s  last E %.6f" % (W, steps, " ".join("%.2f" % t for t in ts), e[-1]), flush=True)
