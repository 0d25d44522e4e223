#!/usr/bin/env python
"""Generates tests/golden/statfor_golden.json by importing the reference's own scripts/statfor.py
(matplotlib is absent in the image and only used for plotting, so it is stubbed) and running its
correlation() - the same algorithm as scripts/statfor.rs:31-54 - on seeded series.
Run in the build container only (/root/reference does not exist on the GPU box):
    python tests/golden/make_statfor_golden.py
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/scripts"
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)
import statfor  # noqa: E402  (the reference module)


def ar1(n, phi, seed):
    rng = np.random.default_rng(seed)
    x = np.empty(n)
    x[0] = rng.normal()
    for i in range(1, n):
        x[i] = phi * x[i - 1] + rng.normal()
    return x - 0.5


cases = []
for name, series in [("white_400", np.random.default_rng(7).normal(size=400) + 1.0),
                     ("ar1_0.8_600", ar1(600, 0.8, 11)),
                     ("ar1_0.95_1000", ar1(1000, 0.95, 12)),
                     ("short_50", ar1(50, 0.5, 13)),
                     ("anticorrelated_300", ar1(300, -0.6, 14))]:
    mean, var = series.mean(), series.var(ddof=1)           # statfor.py:84-85
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            tcorr, neff, sigma = statfor.correlation(series, mean, var)
            corr = [float(l.split()[1]) for l in open("corr.out")]
        finally:
            os.chdir(cwd)
    cases.append(dict(name=name, series=[float(v) for v in series], mean=float(mean), variance=float(var),
                      tcorr=float(tcorr), n_eff=float(neff), sigma=float(sigma), corr=corr))

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "statfor_golden.json")
json.dump(dict(source="scripts/statfor.py correlation() of the reference, corr.out printed with 10 digits", cases=cases),
          open(out, "w"))
print("wrote", out, [(c["name"], round(c["tcorr"], 4)) for c in cases])
