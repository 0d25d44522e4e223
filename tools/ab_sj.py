"""A/B timing of variant builds of libmole_b200.so on the bench workload's dominant kernel:
   python tools/ab_sj.py lib1.so lib2.so ...      (each library is timed in its own process)
One Ne Slater-Jastrow VMC+SR launch of W walkers x NS sweeps (bench.py's shape), wall-clock around a
stream synchronize (the launch is ~70 ms, launch overhead is noise).  Prints ms per launch and the reduced
energy so that variants can be checked against each other.  Variants are built with
MOLE_OUT=... MOLE_OBJ=... MOLE_NVCC_EXTRA="-D..." mole_b200/csrc/build.sh."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one_small(kind, reps):
    """the thread-per-walker kinds at 2^20 walkers x 200 sweeps (H2 Heitler-London + SR, He, Gaussian / H atom)"""
    sys.path.insert(0, ROOT)
    import mole_b200 as m
    ctx = m.default_context()
    SEED = bytes(32)
    W, NS = 1 << 20, 200
    if kind == "h2":
        wf, op, ne, tau = m.HydrogenMoleculeWaveFunction(1.4, [0.5]), m.ElectronicHamiltonian.from_ions([[-0.7, 0, 0], [0.7, 0, 0]], [1, 1]), 2, 0.25
    elif kind == "he":
        wf, op, ne, tau = m.HeliumAtomWaveFunction(1.69), m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [2]), 2, 0.1
    elif kind.startswith("lcao"):                                  # the LCAO determinant kinds (tests/common.py cases)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from common import cases
        c = cases()[kind]
        (wf, op), ne, tau = c["make"](m), c["ne"], 0.1
    else:
        wf, op, ne, tau = m.GaussianWaveFunction(1.0), m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1]), 1, 0.025
    ens = m.Ensemble(W, ne, SEED); ens.init_uniform()
    met = m.MetropolisDiffuse(tau, SEED)
    obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
    ens.sweep(wf, met, op, n_sweeps=20, block_size=10, observables=obs)
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        ens.acc_reset(); ctx.synchronize()
        t0 = time.perf_counter()
        ens.sweep(wf, met, op, n_sweeps=NS, n_discard=10, block_size=10, observables=obs)
        ctx.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    e, de, acc, g = m.acc_finalize(ens.acc_get())
    print("%-28s %-15s ms/launch %s  best %.2f  = %.3e walker-steps/s  E %.10f acc %.6f" % (
        os.path.basename(os.environ.get("MOLE_B200_LIB", "default")), kind, " ".join("%.2f" % t for t in ts), min(ts),
        W * NS / (min(ts) * 1e-3), e, acc), flush=True)


def one(W, NS, reps):
    sys.path.insert(0, ROOT)
    import mole_b200 as m
    ctx = m.default_context()
    SEED = bytes(32)
    wf = m.SlaterJastrow(5, 5, (9.64, 2.88, 2.88), (0.5, 1.0, 0.0, 0.0), 1.0)
    op = m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [10])
    ens = m.Ensemble(W, 10, SEED); ens.init_normal(0.5)
    ens.sweep(wf, m.MetropolisBox.from_rng(0.5, SEED), op, n_sweeps=100, observables=0)
    met = m.MetropolisDiffuse.from_rng(0.02, SEED)
    obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
    ens.sweep(wf, met, op, n_sweeps=20, block_size=10, observables=0)
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        ens.acc_reset()
        ctx.synchronize()
        t0 = time.perf_counter()
        ens.sweep(wf, met, op, n_sweeps=NS, n_discard=10, block_size=10, observables=obs)
        ctx.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    e, de, acc, g = m.acc_finalize(ens.acc_get())
    t0 = time.perf_counter()
    ens.sweep(wf, met, op, n_sweeps=NS, n_discard=10, block_size=10, observables=0)     # moves only
    ctx.synchronize()
    t_moves = 1e3 * (time.perf_counter() - t0)
    print("%-40s ms/launch %s  best %.2f  moves-only %.2f  E %.10f +/- %.2e acc %.6f g0 %.8e" % (
        os.path.basename(os.environ.get("MOLE_B200_LIB", "default")), " ".join("%.2f" % t for t in ts), min(ts), t_moves, e, de, acc, g[0]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    elif len(sys.argv) > 1 and sys.argv[1] == "--one-small":
        one_small(sys.argv[2], int(sys.argv[3]))
    elif len(sys.argv) > 1 and sys.argv[1] == "--small":          # python tools/ab_sj.py --small lib1.so lib2.so
        kinds = ("h2", "he", "gauss", "lcao_h2p", "lcao_he", "lcao_h2_singlet", "lcao_h2_triplet")
        if "--kinds" in sys.argv:                                  # --kinds lcao_h2p,lcao_he
            kinds = sys.argv[sys.argv.index("--kinds") + 1].split(",")
        for kind in kinds:
            for lib in [a for a in sys.argv[2:] if a.endswith(".so")]:
                env = dict(os.environ, MOLE_B200_LIB=os.path.abspath(lib))
                subprocess.run([sys.executable, os.path.abspath(__file__), "--one-small", kind, "3"], env=env, timeout=300)
    else:
        libs = [a for a in sys.argv[1:] if a.endswith(".so")]
        nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
        W, NS, reps = (nums + [1 << 17, 200, 3][len(nums):])[:3]
        for lib in libs:
            env = dict(os.environ, MOLE_B200_LIB=os.path.abspath(lib), MOLE_B200_AB_OLD_LIB="1")
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(W), str(NS), str(reps)], env=env, timeout=300)
