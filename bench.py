#!/usr/bin/env python
"""bench.py — VMC walker-steps/s incl. local energy (BASELINE.json metric) on the synthetic
10-electron Slater-Jastrow VMC + stochastic-reconfiguration configuration (configs[4]).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N>1 under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)
  python bench.py --workload dmc ...                       # configs[3]: DMC with SRBrancher, 2^15 walkers per GPU

Workload "vmc" (default): a "step" is one SR optimisation iteration of VmcRunner::run_optimization
(vmc.rs:55-99) on this rank's walker shard: reseed, `sweeps` Metropolis sweeps with the first block discarded,
local energy + O_k + SR moments accumulated on the device, allreduce of the accumulator vector, host finaliser,
SR solve, parameter update.  value = all ranks' walker-sweeps / max-over-ranks device time.
The optimisation is REAL (parameters move every step) and must stay healthy: S is regularised with an absolute
diagonal shift (mole_opt_set_sr_regularization) because the Jastrow pair (b1, b2) is nearly redundant and the
reference's diag x 1.01 alone lets the loop diverge after ~20 iterations; the run aborts with a
one-line reason if a sample was non-finite, a parameter left its sane range or the energy left the physical window.

Workload "dmc": a "step" is one block of `--dmc-steps` DMC time steps (dmc.rs:84-141: E_L, N_e drift-diffusion
moves, E_L, weight update, ensemble energy, SRBrancher) enqueued without host reads, ranks as population islands
with ONE all-gather per block; the reference energy is updated between blocks as DmcRunner::update_energies does.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ZETA = (9.64, 2.88, 2.88)
JB = (0.5, 1.0, 0.0, 0.0)
KAPPA = 1.0
TAU = 0.02
BLOCK = 10
SR_STEP = 0.02          # Delta p = SR_STEP * S^-1 (-g/2)
SR_DIAG = (1.01, 1e-2)  # S_kk <- 1.01 S_kk + 1e-2 (reference: 1.01, 0)
E_WINDOW = (-129.2, -126.5)   # Ne: exact -128.94; the equilibrated start is near -127.2, 30 SR steps reach -128.5
SEED = bytes(32)
FLOPS = json.load(open(os.path.join(ROOT, "bench_data", "flops.json")))
# DMC (BASELINE.json configs[3], examples/dmc.rs:189-210): H atom, Gaussian guide at its VMC optimum 1/a^2 = 8/(9 pi)
DMC_A = float(np.sqrt(9.0 * np.pi / 8.0))   # psi = exp(-(r/a)^2), examples/dmc.rs:44-57: optimum 1/a^2 = 8/(9 pi)
DMC_TAU = 0.025
REBALANCE_RATIO = 1.05      # include/mole_b200.h MOLE_REBALANCE_RATIO
# 1s STO guide with alpha > 1: E_L = -alpha^2/2 + (alpha - 1)/r is bounded BELOW, so walkers next to the nucleus die instead of
# multiplying.  With alpha < 1 (or the example's cusp-less Gaussian) E_L -> -inf at the nucleus and the reference's cut-off-free
# DMC collapses a large population onto it sooner or later (observed at 2 x 2^15 walkers after ~4400 steps with alpha = 0.9:
# island energy -55, weights 1e192, then overflow; upstream behaviour)
DMC_STO_ALPHA = 1.1
DMC_EREF = {"gaussian": -0.4244, "sto": -0.495}   # VMC energies of the guides: -4/(3 pi); alpha^2/2 - alpha at alpha = 1.1
DMC_SEED = bytes([1] * 32)


def fail(reason):
    """One line, then a non-zero exit: never a bare rc=1 (VERDICT r1 next#1a)."""
    print("bench.py: ABORT: " + reason, file=sys.stderr, flush=True)
    sys.exit(3)


def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return
    while not stop.is_set():
        line = p.stdout.readline()
        if not line:
            break
        out.append([t.strip() for t in line.split(",")])
    p.terminate()


def summarize_clocks(rows):
    if not rows:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
    mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in rows)]
    load = sorted(sm)[len(sm) // 2:] if sm else []          # upper half = samples taken under load
    return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------------------------------------ workload descriptions
def vmc_config(args, world, walkers=None, sweeps=None, bounded=False):
    walkers = args.walkers if walkers is None else walkers
    sweeps = args.sweeps if sweeps is None else sweeps
    cfg = {"workload": "ne_slater_jastrow_vmc_sr (BASELINE.json configs[4]): Ne, 10 electrons (5 up, 5 dn), STO 1s/2s/2p Slater "
                       "determinants x Pade-polynomial e-e Jastrow, P=7, MetropolisDiffuse tau=%.3g, SR step %g, S diag x %g + %g"
                       % (TAU, args.sr_step, SR_DIAG[0], args.sr_shift),
           "walkers_per_gpu": walkers, "global_walkers": walkers * world, "sweeps_per_step": sweeps,
           "block_size": BLOCK, "sampled_sweeps_per_step": sweeps - BLOCK, "parallelism": "walkers sharded, dp%d" % world,
           "cache": "L2 flushed between timed steps (256 MiB write); walker state is read once per %d-sweep launch" % sweeps}
    if bounded:
        # the CPU arm times a BOUNDED SAMPLE of the GPU arm's workload (same physics, same sampler, fewer walkers and
        # sweeps per step): the driver's ratio is a throughput ratio, not a same-size comparison
        cfg["bounded_sample"] = True
        cfg["gpu_arm_walkers_per_gpu"] = args.walkers
        cfg["gpu_arm_sweeps_per_step"] = args.sweeps
        cfg["parallelism"] = "OpenMP over walkers on rank 0's host cores"
        cfg["cache"] = "n/a (CPU)"
    return cfg


def dmc_config(args, world, walkers=None, steps=None, bounded=False):
    walkers = args.walkers if walkers is None else walkers
    steps = args.dmc_steps if steps is None else steps
    cfg = {"workload": "h_atom_dmc_sr_brancher (BASELINE.json configs[3], examples/dmc.rs:189-210): H atom, %s guide, "
                       "MetropolisDiffuse tau=%.3g, SRBrancher, E_ref updated between blocks" % (
                           "Gaussian 1/a^2=8/(9 pi)" if args.dmc_guide == "gaussian" else "1s STO alpha=%.2f" % DMC_STO_ALPHA, DMC_TAU),
           "walkers_per_gpu": walkers, "global_walkers": walkers * world, "time_steps_per_step": steps,
           "parallelism": "walkers sharded (population islands, one all-gather per block), dp%d" % world,
           "cache": "L2 flushed between timed steps (256 MiB write)"}
    if bounded:
        cfg["bounded_sample"] = True
        cfg["gpu_arm_walkers_per_gpu"] = args.walkers
        cfg["gpu_arm_time_steps_per_step"] = args.dmc_steps
        cfg["parallelism"] = "oracle port, one host thread (DmcRunner::diffuse is serial upstream)"
        cfg["cache"] = "n/a (CPU)"
    return cfg


def dmc_pair(m_or_o, guide, ctx=None, oracle=False):
    if oracle:
        O = m_or_o
        wf = O.wf_desc(O.WF_GAUSSIAN, [DMC_A]) if guide == "gaussian" else O.wf_desc(O.WF_STO_1S, [DMC_STO_ALPHA])
        return wf, O.ham_desc(O.HAM_ELECTRONIC, [[0, 0, 0]], [1])
    m = m_or_o
    wf = m.GaussianWaveFunction(DMC_A, ctx=ctx) if guide == "gaussian" else m.STO(DMC_STO_ALPHA, ctx=ctx)
    return wf, m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [1], ctx=ctx)


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def run_reference(args, rank, world):
    """The reference's own CPU algorithm for this path (oracle port: the Rust reference cannot be
    built in this image), all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses all host threads (libgomp reads this at load time)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import oracle as O
    O.build()
    times = []
    threads = 1
    if args.workload == "vmc":
        W, sweeps = args.ref_walkers, args.ref_sweeps
        wf = O.wf_desc(O.WF_SLATER_JASTROW, list(ZETA) + list(JB), [KAPPA, 5, 5])
        ham = O.ham_desc(O.HAM_ELECTRONIC, [[0, 0, 0]], [10])
        opts = O.run_options(O.METROP_DIFFUSE, TAU, O.OBS_ENERGY | O.OBS_PGRAD | O.OBS_WFVALUE)
        cfgs = np.array([O.init_normal(SEED, w, 10, 0.5) for w in range(W)])
        for it in range(args.warmup + args.steps):
            secs, esum, threads = O.bench_vmc(wf, ham, opts, cfgs, sweeps, BLOCK, O.derive_seed(SEED, it))
            if it >= args.warmup:
                times.append(secs)
        units = W * sweeps
        config = vmc_config(args, world, W, sweeps, bounded=True)
        metric = "vmc_walker_steps_per_sec_incl_local_energy"
        sample = "%d walkers x %d sweeps per step (reference-faithful: 3 value+gradient evaluations per diffusion move, " \
                 "from-scratch determinants), OpenMP over walkers" % (W, sweeps)
    else:
        W, nst = args.ref_dmc_walkers, args.ref_dmc_steps
        wf, ham = dmc_pair(O, args.dmc_guide, oracle=True)
        cfgs = np.array([O.init_normal(DMC_SEED, w, 1, 1.0) for w in range(W)])
        wts, eref = np.ones(W), DMC_EREF[args.dmc_guide]
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            r = O.dmc_diffuse(wf, ham, wts, cfgs, DMC_TAU, eref, 0, O.derive_seed(DMC_SEED, it), DMC_TAU, nst, nst, 0)
            secs = time.perf_counter() - t0
            wts, cfgs = r["weights"], r["cfgs"]
            if it >= args.warmup:
                times.append(secs)
        units = W * nst
        config = dmc_config(args, world, W, nst, bounded=True)
        metric = "dmc_walker_steps_per_sec"
        sample = "%d walkers x %d time steps per step, DmcRunner::diffuse restated (serial, as upstream)" % (W, nst)
    total = sum(times)
    value = units * args.steps / total
    line = {"impl": "reference", "metric": metric, "value": value, "unit": "walker-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "walker-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ timing scaffold
class Rig:
    """Device context, torch.distributed plumbing, barrier, clocks sampler and the timed loop shared by the workloads."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        # NCCL_DEBUG=VERSION makes NCCL print its version banner on stdout, ahead of the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist
        import mole_b200 as m
        self.torch, self.dist, self.m = torch, dist, m
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.ctx = m.Context(self.local_rank)
        if self.world > 1:
            idt = torch.zeros(m.ffi.NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                idt.copy_(torch.tensor(list(m.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            self.ctx.comm_init(self.world, self.rank, bytes(idt.cpu().tolist()))
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=torch.device("cuda", self.local_rank))
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        self.ctx.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, reset):
        """W untimed + K timed calls of step(it, timed) between barriers; CUDA events on the context's stream."""
        torch, args = self.torch, self.args
        reset()
        # the nvidia-smi sampler is started BEFORE the warm-up: its start-up takes the driver lock for
        # ~100 ms and must not land inside the timed region; only rows taken inside it are summarised
        rows, stop = [], threading.Event()
        th = threading.Thread(target=clocks_sampler, args=(stop, rows, self.local_rank), daemon=True)
        th.start()
        t_wait = time.time()
        while not rows and time.time() - t_wait < 5.0:
            time.sleep(0.05)
        for it in range(args.warmup):
            with torch.cuda.stream(self.stream):
                self.flush.zero_()                                  # also loads torch's fill kernel outside the timed region
            step(it, False)
        self.barrier()
        n0 = len(rows)
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        l0 = self.ctx.launch_count()
        t0.record(self.stream)
        for it in range(args.steps):
            with torch.cuda.stream(self.stream):
                self.flush.zero_()                                  # L2 flush between timed iterations
            step(args.warmup + it, True)
        t1.record(self.stream)
        self.barrier()
        ms = t0.elapsed_time(t1)
        n1 = len(rows)
        stop.set()
        th.join(timeout=2)
        rows = rows[n0:max(n1, n0 + 1)]
        launches = self.ctx.launch_count() - l0
        return self.max_over_ranks(ms), launches, summarize_clocks(rows)

    def emit(self, line):
        if self.rank == 0:
            print(json.dumps(line))
        if self.world > 1:
            self.dist.barrier()

    def close(self):
        if self.world > 1:
            self.ctx.close()                                        # after the ensembles (the library also defers the free)
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ VMC + SR (configs[4])
def run_vmc(args):
    rig = Rig(args)
    m, torch, ctx, rank, world = rig.m, rig.torch, rig.ctx, rig.rank, rig.world
    import ctypes as C
    lib = m.ffi.lib()
    W = args.walkers
    wf = m.SlaterJastrow(5, 5, ZETA, JB, KAPPA, ctx=ctx)
    op = m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [10], ctx=ctx)
    met = m.MetropolisDiffuse.from_rng(TAU, SEED)
    opt = m.StochasticReconfiguration(args.sr_step, 7).set_regularization(SR_DIAG[0], args.sr_shift)
    ens = m.Ensemble(W, 10, SEED, walker_offset=rank * W, ctx=ctx)
    ens.init_normal(0.5)
    # untimed equilibration: N(0,0.5) is far from |psi|^2, and the reference's drift-diffusion move has
    # no drift limiting, so walkers that START next to a node stay there for a very long time.  Box moves
    # (metrop.rs:60-96) carry no drift and relax the ensemble first; the production sampler then takes over.
    ens.sweep(wf, m.MetropolisBox.from_rng(0.5, SEED), op, n_sweeps=args.equil_box_sweeps, observables=0)
    ens.sweep(wf, met, op, n_sweeps=args.equil_diffuse_sweeps, observables=0)
    obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
    params0 = wf.parameters().copy()
    host_cfgs = torch.empty((W, 10, 3), dtype=torch.float64, pin_memory=True)
    host_cfgs.numpy()[...] = ens.get_configs()
    kernel_ms, hist = [], []
    state = {"e2e": False, "bad": 0}

    def step(it, timed):
        if state["e2e"]:
            ens.set_configs(host_cfgs.numpy())                      # H2D of this step's walkers from pinned memory
        ens.reseed(m.derive_seed(SEED, it))                         # vmc.rs:59-61
        ens.acc_reset()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(rig.stream)
        m.ffi.check(lib.mole_runner_run(ens.handle, wf.handle, met.handle, op.handle, C.c_uint32(obs), C.c_uint32(0),
                                        C.c_int32(args.sweeps), C.c_int32(BLOCK), None, None, None, None, None), ctx.handle)
        e1.record(rig.stream)
        if world > 1:
            ens.acc_allreduce()                                     # concatenate_worker_data, vmc.rs:108-130
        acc = ens.acc_get()                                         # D2H of the reduced moments
        e, err, accp, g = m.acc_finalize(acc)
        try:
            dp = opt.compute_parameter_update(wf.parameters(), acc)  # SR: S^-1 (-g/2), optimizers.rs:237-252
        except m.MoleError as ex:
            fail("step %d: SR update refused: %s (health %s)" % (it, ex, ens.health()))
        wf.update_parameters(dp)
        hist.append((e, err, accp))
        if args.verbose and rank == 0:
            print("  step %d (%s): E = %.5f +/- %.5f  p = %s" % (it, "e2e" if state["e2e"] else "resident", e, err,
                                                                np.array2string(wf.parameters(), precision=4)), file=sys.stderr)
        if timed and not state["e2e"]:
            kernel_ms.append((e0, e1))

    def reset():
        wf.set_parameters(params0)
        del hist[:]

    def check(leg):
        """the run must mean something at every N: finite, healthy, inside the physical window, parameters sane"""
        bad, _ = ens.health()
        state["bad"] += bad
        p = wf.parameters()
        es = np.array([h[0] for h in hist])
        if not np.all(np.isfinite(es)) or not np.all(np.isfinite(p)):
            fail("%s leg: non-finite energy or parameter (E = %s, p = %s)" % (leg, es[-3:], p))
        if es.min() < E_WINDOW[0] or es.max() > E_WINDOW[1]:
            fail("%s leg: energy left the physical window %s: min %.4f max %.4f (SR diverged?)" % (leg, E_WINDOW, es.min(), es.max()))
        if np.any(np.abs(p[:3] - params0[:3]) > 0.5 * params0[:3]) or np.any(np.abs(p[3:] - params0[3:]) > 3.0):
            fail("%s leg: parameters drifted out of range: %s (start %s)" % (leg, p, params0))
        if bad > 1e-6 * W * world * args.sweeps:
            fail("%s leg: %d non-finite samples in the last step" % (leg, bad))

    ms, launches, clocks = rig.timed(step, reset)
    check("resident")
    first, last, p_end = hist[0], hist[-1], wf.parameters().copy()
    kms = rig.max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in kernel_ms])))
    state["e2e"] = True
    ms_e2e, _, _ = rig.timed(step, reset)
    check("e2e")
    fp64_peak = ctx.fp64_peak_tflops()

    wsteps = W * world * args.sweeps * args.steps
    value = wsteps / (ms * 1e-3)
    value_e2e = wsteps / (ms_e2e * 1e-3)
    fl = FLOPS["ne_slater_jastrow_vmc_sr"]
    f_alg = fl["f_alg_per_walker_step"]
    kernel_rate = W * args.sweeps / (kms * 1e-3)                    # per GPU, dominant kernel only
    achieved = kernel_rate * f_alg / 1e12
    nominal = FLOPS["fp64_nominal_tflops"]
    line = {
        "metric": "vmc_walker_steps_per_sec_incl_local_energy", "value": value, "unit": "walker-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": vmc_config(args, world),
        "sr_opt_step_ms": ms / args.steps,
        # energy of the LAST timed step (parameters after warmup+steps SR updates) and of the FIRST step (the starting
        # parameters: the number to compare between runs at different N)
        "energy": {"value": last[0], "blocking_error": last[1], "acceptance": last[2]},
        "energy_first_step": {"value": first[0], "blocking_error": first[1]},
        "parameters_end": [float(x) for x in p_end],
        "health": {"nonfinite_samples": int(state["bad"])},
        "roofline": {"bound": "fp64", "kernel": "sj_sweep_kernel<DIFFUSE,OPT>", "achieved": achieved, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": fl.get("dram_bytes_per_launch"),
                     "peak_source": "DFMA chain measured live by mole_bench_fp64_peak (MEASURED_PEAKS.json has no fp64 entry); "
                                    "nominal 148 SM x 64 lanes x 2 x 1.965 GHz = %.1f" % nominal,
                     "frac_of_nominal": achieved / nominal, "f_alg_per_walker_step": f_alg,
                     "kernel_ms_per_launch": kms, "kernel_walker_steps_per_s_per_gpu": kernel_rate,
                     "frac_from_scratch_count": kernel_rate * fl["f_from_scratch_per_walker_step"] / 1e12 / fp64_peak,
                     # NOT measured in this run: ncu counters of the same launch (profiles/), for context
                     "ncu_reference": fl.get("ncu_executed_fp64")},
        "e2e": {"value": value_e2e, "unit": "walker-steps/s", "h2d_bytes_per_step": W * 30 * 8 * world,
                "d2h_bytes_per_step": 62 * 8 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_extras:
        line["sr_opt_step_ms_h2"] = h2_sr_step(rig)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O                                          # cpu_baseline leg: the oracle as the timed CPU port
        O.build()
        Wc, Sc = args.cpu_baseline_walkers, args.cpu_baseline_sweeps
        owf = O.wf_desc(O.WF_SLATER_JASTROW, list(ZETA) + list(JB), [KAPPA, 5, 5])
        oham = O.ham_desc(O.HAM_ELECTRONIC, [[0, 0, 0]], [10])
        oopts = O.run_options(O.METROP_DIFFUSE, TAU, O.OBS_ENERGY | O.OBS_PGRAD | O.OBS_WFVALUE)
        ocfg = np.array([O.init_normal(SEED, w, 10, 0.5) for w in range(Wc)])
        secs, _, threads = O.bench_vmc(owf, oham, oopts, ocfg, Sc, BLOCK, SEED)
        line["cpu_baseline"] = {"value": Wc * Sc / secs, "unit": "walker-steps/s", "cores": threads, "kind": "port",
                                "sample": "%d walkers x %d sweeps of the same workload, reference-faithful oracle "
                                          "(oracle/, OpenMP over walkers), %.1f s" % (Wc, Sc, secs)}
    rig.emit(line)
    del ens
    rig.close()


def h2_sr_step(rig):
    """BASELINE.json metric, second half: SR optimisation-step time of configs[0] at the GPU size of SURVEY 8(d) #1 -
    H2 Heitler-London STO, 2^16 walkers, one iteration = 2500 sweeps (block 10, 250 blocks), P = 1, Diffuse tau = 0.25
    (examples/hydrogen_molecule.rs:191-199,252).  Extra key, N = 1 only; device-timed like the main loop."""
    m, torch, ctx = rig.m, rig.torch, rig.ctx
    W, steps, iters = 1 << 16, 2500, 5
    wf = m.HydrogenMoleculeWaveFunction(1.4, [0.5], ctx=ctx)
    op = m.ElectronicHamiltonian.from_ions([[-0.7, 0, 0], [0.7, 0, 0]], [1, 1], ctx=ctx)
    met = m.MetropolisDiffuse.from_rng(0.25, SEED)
    opt = m.StochasticReconfiguration(0.05, 1)
    ens = m.Ensemble(W, 2, SEED, ctx=ctx)
    ens.init_uniform(-1.0, 1.0)
    obs = m.ffi.OBS_ENERGY | m.ffi.OBS_PGRAD | m.ffi.OBS_WFVALUE
    out = []

    def one(it):
        ens.reseed(m.derive_seed(SEED, it))
        ens.acc_reset()
        ens.sweep(wf, met, op, n_sweeps=steps, n_discard=BLOCK, block_size=BLOCK, observables=obs)
        acc = ens.acc_get()
        e, err, _, _ = m.acc_finalize(acc)
        wf.update_parameters(opt.compute_parameter_update(wf.parameters(), acc))
        return e, err

    one(0)
    rig.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(rig.stream)
    for it in range(iters):
        out.append(one(1 + it))
    t1.record(rig.stream)
    rig.barrier()
    ms = t0.elapsed_time(t1) / iters
    return {"value": ms, "unit": "ms per SR iteration", "walkers": W, "sweeps_per_iteration": steps, "iterations_timed": iters,
            "walker_steps_per_s": W * steps / (ms * 1e-3), "energy": out[-1][0], "blocking_error": out[-1][1],
            "alpha_end": float(wf.parameters()[0]),
            "workload": "hydrogen_molecule VMC + SR (BASELINE.json configs[0] at 2^16 walkers), intended O_k, SR step 0.05"}


# ------------------------------------------------------------------------------------------------ DMC (configs[3])
def run_dmc(args):
    rig = Rig(args)
    m, torch, ctx, rank, world = rig.m, rig.torch, rig.ctx, rig.rank, rig.world
    W, S = args.walkers, args.dmc_steps
    wf, op = dmc_pair(m, args.dmc_guide, ctx=ctx)
    met = m.MetropolisDiffuse.from_rng(DMC_TAU, DMC_SEED).fix_nodes()       # examples/dmc.rs:196
    ens = m.Ensemble(W, 1, DMC_SEED, walker_offset=rank * W, ctx=ctx)
    ens.init_normal(1.0)                                                    # independent starts (dmc.rs:49-58 clones one)
    ens.sweep(wf, met, op, n_sweeps=200, observables=0)                     # untimed: sample |psi|^2 first
    ens.dmc_block_select(args.dmc_block_impl)
    host_cfgs = torch.empty((W, 1, 3), dtype=torch.float64, pin_memory=True)
    host_cfgs.numpy()[...] = ens.get_configs()
    host_w = torch.ones(W, dtype=torch.float64, pin_memory=True)
    state = {"e2e": False, "eref": DMC_EREF[args.dmc_guide], "bad": 0}
    if world > 1:
        ens.rebalance()      # untimed: opens the NCCL send/recv channels once (~0.3 s); the islands start with equal weights anyway
    hist = []

    def step(it, timed):
        if state["e2e"]:
            ens.set_configs(host_cfgs.numpy())                              # H2D of this step's walkers
            ens.set_weights(host_w.numpy())
        se = ens.dmc_block(wf, met, op, m.ffi.BRANCH_SR, DMC_TAU, state["eref"], S)   # D2H: the block's step energies
        eb = float(se.mean())
        if not np.isfinite(eb):
            fail("DMC block %d: non-finite ensemble energy (health %s)" % (it, ens.health()))
        if args.verbose:
            print("  rank %d block %d (%s): E = %.5f  [%.4f, %.4f]  E_ref %.5f" % (rank, it, "e2e" if state["e2e"] else "resident", eb,
                                                                             se.min(), se.max(), state["eref"]), file=sys.stderr, flush=True)
        state["eref"] = 0.5 * (state["eref"] + eb)                          # dmc.rs:163-177
        hist.append(eb)
        if world > 1:                                                       # as mole_dmc_diffuse does between blocks
            ratio = ens.island_imbalance()
            if args.verbose:
                print("  rank %d block %d: island weight ratio %.6f" % (rank, it, ratio), file=sys.stderr, flush=True)
            if ratio > REBALANCE_RATIO:
                t_r = time.perf_counter()
                ens.rebalance()
                state["rebalanced"] = state.get("rebalanced", 0) + 1
                if args.verbose:
                    ctx.synchronize()
                    print("  rank %d block %d: rebalanced in %.3f ms" % (rank, it, 1e3 * (time.perf_counter() - t_r)), file=sys.stderr, flush=True)

    def reset():
        state["eref"] = DMC_EREF[args.dmc_guide]
        del hist[:]
        ens.acc_reset()

    ms, launches, clocks = rig.timed(step, reset)
    state["bad"] += ens.health()[1]
    e_mean = float(np.mean(hist[args.warmup:]))
    e_err = float(np.std(hist[args.warmup:], ddof=1) / np.sqrt(max(len(hist) - args.warmup, 1))) if len(hist) - args.warmup > 1 else None
    if not (-0.56 < e_mean < -0.44) or not (-0.7 < state["eref"] < -0.3):
        fail("DMC energy %.5f (reference energy %.4f) left the physical window around -0.5" % (e_mean, state["eref"]))
    state["e2e"] = True
    ms_e2e, _, _ = rig.timed(step, reset)
    state["bad"] += ens.health()[1]
    fp64_peak = ctx.fp64_peak_tflops()
    units = W * world * S * args.steps
    value = units / (ms * 1e-3)
    f_alg = FLOPS["h_atom_dmc"]["f_alg_per_walker_step"][args.dmc_guide]
    achieved = (W * S * args.steps / (ms * 1e-3)) * f_alg / 1e12
    line = {
        "metric": "dmc_walker_steps_per_sec", "value": value, "unit": "walker-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": dmc_config(args, world),
        "us_per_time_step": 1e3 * ms / (args.steps * S),
        "island_rebalances": state.get("rebalanced", 0),
        "energy": {"value": e_mean, "error_over_blocks": e_err, "exact": -0.5, "reference_energy_end": state["eref"]},
        "health": {"nonfinite_dmc_walker_steps": int(state["bad"])},
        "roofline": {"bound": "latency (one cooperative launch per block: ~7 dependent L2 round trips and 2 grid barriers per time step; fp64 figure for context)", "kernel": "dmc_block_kernel<KIND>",
                     "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                     "f_alg_per_walker_step": f_alg, "traffic": None},
        "e2e": {"value": units / (ms_e2e * 1e-3), "unit": "walker-steps/s", "h2d_bytes_per_step": W * 4 * 8 * world,
                "d2h_bytes_per_step": 2 * S * 8 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O
        O.build()
        Wc, Sc = args.ref_dmc_walkers, 20 * args.ref_dmc_steps
        owf, oham = dmc_pair(O, args.dmc_guide, oracle=True)
        ocfg = np.array([O.init_normal(DMC_SEED, w, 1, 1.0) for w in range(Wc)])
        t0 = time.perf_counter()
        O.dmc_diffuse(owf, oham, np.ones(Wc), ocfg, DMC_TAU, DMC_EREF[args.dmc_guide], 0, DMC_SEED, DMC_TAU, Sc, Sc, 0)
        secs = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": Wc * Sc / secs, "unit": "walker-steps/s", "cores": 1, "kind": "port",
                                "sample": "%d walkers x %d time steps, DmcRunner::diffuse restated (serial, as upstream), %.1f s" % (Wc, Sc, secs)}
    rig.emit(line)
    del ens
    rig.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mole_b200", choices=["mole_b200", "reference"])
    ap.add_argument("--workload", default="vmc", choices=["vmc", "dmc"])
    ap.add_argument("--walkers", type=int, default=None, help="walkers per GPU (vmc: 2^17, i.e. 2^20 over 8 GPUs; dmc: 2^15, i.e. 2^18 over 8)")
    ap.add_argument("--sweeps", type=int, default=200, help="vmc: sweeps per optimisation iteration (first block discarded)")
    ap.add_argument("--sr-step", type=float, default=SR_STEP)
    ap.add_argument("--sr-shift", type=float, default=SR_DIAG[1])
    ap.add_argument("--dmc-steps", type=int, default=400, help="dmc: time steps per block (examples/dmc.rs:192)")
    ap.add_argument("--dmc-block-impl", type=int, default=0, choices=[0, 1, 2],
                    help="dmc: 0 = one persistent launch per block where eligible (default), 1 = per-step launches (A/B)")
    ap.add_argument("--dmc-guide", default="sto", choices=["gaussian", "sto"],
                    help="sto: 1s STO alpha=0.9, the guide examples/dmc.rs:157 keeps as its commented alternative (stable); gaussian: "
                         "the example's cusp-less guide, whose E_L -> -inf at the nucleus makes large populations collapse (upstream behaviour)")
    # CPU legs: bounded samples of the same workload (~4e4 walker-steps/s on 16 cores):
    # reference arm ~1 s per step, cpu_baseline ~10 s in total
    ap.add_argument("--ref-walkers", type=int, default=1024)
    ap.add_argument("--ref-sweeps", type=int, default=40)
    ap.add_argument("--ref-dmc-walkers", type=int, default=32768)
    ap.add_argument("--ref-dmc-steps", type=int, default=100)
    ap.add_argument("--cpu-baseline-walkers", type=int, default=1024)
    ap.add_argument("--cpu-baseline-sweeps", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the H2 SR-step timing (extra key at N=1)")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--equil-box-sweeps", type=int, default=200)
    ap.add_argument("--equil-diffuse-sweeps", type=int, default=50)
    args = ap.parse_args()
    if args.walkers is None:
        args.walkers = (1 << 17) if args.workload == "vmc" else (1 << 15)
    args.warmup = max(args.warmup, 3) if args.impl == "mole_b200" else args.warmup

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
    elif args.workload == "vmc":
        run_vmc(args)
    else:
        run_dmc(args)


if __name__ == "__main__":
    main()
