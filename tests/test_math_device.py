"""Accuracy of the kernels' branch-free fp64 elementary functions (mole_b200/csrc/mole_math.cuh),
evaluated on the device through mole_math_probe and compared with numpy in units of last place."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ulp_err(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_exp(ctx):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-2, 2, 200000), -rng.exponential(5.0, 200000),
                        np.linspace(-30, 0, 10001)])
    got = ctx.math_probe(0, x)
    assert ulp_err(got, np.exp(x)).max() <= 2.0
    # tails and special values: underflow to 0 (through denormals), overflow to inf, NaN propagates
    edge = np.array([-1e300, -800.0, -746.0, -745.2, -740.0, -720.0, -708.5, 0.0, 709.7, 709.9, 720.0, 1e300, np.nan, -0.0])
    g = ctx.math_probe(0, edge)
    with np.errstate(over="ignore", under="ignore"):
        r = np.exp(edge)
    assert g[0] == 0 and g[1] == 0 and g[2] == 0 and np.isinf(g[9]) and np.isinf(g[10]) and np.isinf(g[11]) and np.isnan(g[12])
    assert g[7] == 1.0 and g[13] == 1.0
    den = np.abs(g[3:6] - r[3:6]) / 4.94e-324          # denormal results: within a few denormal steps
    assert np.all(den <= 2**12 * np.maximum(r[3:6] / 2.2e-308, 1e-16) + 4)
    assert ulp_err(g[6:7], r[6:7]).max() <= 2 and ulp_err(g[8:9], r[8:9]).max() <= 2


def test_rcp_rsqrt_sqrt(ctx):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(1e-3, 1e3, 300000), 10.0 ** rng.uniform(-200, 200, 300000)])
    assert ulp_err(ctx.math_probe(1, x), 1.0 / x).max() <= 2.0
    assert ulp_err(ctx.math_probe(1, -x), -1.0 / x).max() <= 2.0
    assert ulp_err(ctx.math_probe(2, x), 1.0 / np.sqrt(x)).max() <= 2.0
    assert ulp_err(ctx.math_probe(3, x), np.sqrt(x)).max() <= 1.0
