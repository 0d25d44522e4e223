"""The large-P path's two kernels measured alone (B200): the Gram contraction (DMMA vs FP64 vector pipe, HBM roofline)
and the general LCAO Slater-Jastrow sweep (H8, P = 36).   python tools/prof_gram.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mole_b200 as m  # noqa: E402

ctx = m.default_context()
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    pass
hbm = peaks.get("hbm_gbs", 6650.0)
print("HBM peak used: %.0f GB/s (%s)" % (hbm, "MEASURED_PEAKS.json" if peaks else "fallback"))
for ch in (1, 2, 4, 8):
    print("dmma m8n8k4 probe, %2d chains/warp: " % ch + "  ".join("%dw %.1f" % (w, ctx.dmma_peak_tflops(ch, w)) for w in (4, 8, 16, 32)) + "  TFLOP/s", flush=True)
for W, ns, cols in ((1 << 16, 190, 38), (1 << 17, 40, 38), (1 << 16, 190, 14), (1 << 14, 190, 46)):
    out = []
    for impl in (0, 1):
        ms, gbs, cs = ctx.bench_gram(W, ns, cols, impl, 5)
        flops = (cols * (cols + 1)) * W * ns          # upper triangle, 2 flops per fma
        out.append((ms, gbs, cs))
        print("gram %-5s W=%-7d samples=%-4d cols=%-3d  %8.3f ms  %7.1f GB/s (%.2f of HBM)  %6.2f TFLOP/s" % (
            "dmma" if impl == 0 else "fma", W, ns, cols, ms, gbs, gbs / hbm, flops / (ms * 1e-3) / 1e12), flush=True)
    assert abs(out[0][2] - out[1][2]) <= 1e-9 * abs(out[1][2]), out

from common import cases  # noqa: E402
c = cases()["lsj_h8"]
wf, op = c["make"](m)
seed = bytes(32)
for W in (1 << 14, 1 << 16):
    ens = m.Ensemble(W, 8, seed)
    ens.init_uniform(-4.0, 4.0)
    met = m.MetropolisDiffuse.from_rng(0.05, seed)
    ens.sweep(wf, m.MetropolisBox.from_rng(1.0, seed), op, n_sweeps=20, observables=0)
    obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
    for label, o in (("moves only", 0), ("VMC + SR rows + Gram", obs)):
        ens.acc_reset(); ctx.synchronize()
        t0 = time.perf_counter()
        ens.sweep(wf, met, op, n_sweeps=50, n_discard=10, block_size=10, observables=o)
        ctx.synchronize()
        dt = time.perf_counter() - t0
        print("lsj_h8 (P=36) W=%-6d 50 sweeps %-22s %8.2f ms  %.3e walker-steps/s" % (W, label, 1e3 * dt, W * 50 / dt), flush=True)
