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
    assert ulp_err(ctx.math_probe(3, x), np.sqrt(x)).max() <= 2.0      # x * rsqrt(x), no Heron correction


def test_log_sincos_turn(ctx):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(0.0, 1.0, 300000) + 2.0 ** -53, 2.0 ** -rng.uniform(0, 53, 100000), 10.0 ** rng.uniform(-300, 300, 100000)])
    got = ctx.math_probe(4, x)
    ref = np.log(x)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)) < 4.5e-16
    assert ctx.math_probe(4, np.array([1.0]))[0] == 0.0
    a = np.concatenate([rng.uniform(0.0, 1.0, 400000), np.arange(0, 2 ** 12) / 2.0 ** 12, [0.0, 0.25, 0.5, 0.75, 1.0 - 2.0 ** -32]])
    s, c = ctx.math_probe(5, a), ctx.math_probe(6, a)
    # numpy's sin(2*pi*a) carries the rounding of 2*pi*a (~9e-16 in the angle); mpmath on a subset is the tight check
    assert np.max(np.abs(s - np.sin(2 * np.pi * a))) < 2e-15 and np.max(np.abs(c - np.cos(2 * np.pi * a))) < 2e-15
    assert np.max(np.abs(s * s + c * c - 1.0)) < 6e-16
    import mpmath as mp
    mp.mp.dps = 30
    for i in range(0, 3000, 3):
        t = 2 * mp.pi * mp.mpf(float(a[i]))
        assert abs(mp.mpf(float(s[i])) - mp.sin(t)) < 3e-16 and abs(mp.mpf(float(c[i])) - mp.cos(t)) < 3e-16
    # exact quadrant values (numpy's sin(2 pi a) is not exact there)
    assert list(ctx.math_probe(5, np.array([0.0, 0.25, 0.5, 0.75]))) == [0.0, 1.0, 0.0, -1.0]
    assert list(ctx.math_probe(6, np.array([0.0, 0.25, 0.5, 0.75]))) == [1.0, 0.0, -1.0, 0.0]
