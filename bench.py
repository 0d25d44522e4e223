#!/usr/bin/env python
"""bench.py — VMC walker-steps/s incl. local energy (BASELINE.json metric) on the synthetic
10-electron Slater-Jastrow VMC + stochastic-reconfiguration configuration (configs[4]).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N>1 under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one SR optimisation iteration of VmcRunner::run_optimization (vmc.rs:55-99) on this
rank's walker shard: reseed, `sweeps` Metropolis sweeps with the first block discarded, local energy
+ O_k + SR moments accumulated on the device, allreduce of the 62-double accumulator, host finaliser,
SR solve, parameter update.  value = all ranks' walker-sweeps / max-over-ranks device time.
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
SR_STEP = 0.005
SEED = bytes(32)
FLOPS = json.load(open(os.path.join(ROOT, "bench_data", "flops.json")))


def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
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


def run_reference(args, rank, world):
    """The reference's own CPU algorithm for this path (oracle port: the Rust reference cannot be
    built in this image), all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses all host threads (libgomp reads this at load time)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import oracle as O
    O.build()
    W, sweeps = args.ref_walkers, args.ref_sweeps
    wf = O.wf_desc(O.WF_SLATER_JASTROW, list(ZETA) + list(JB), [KAPPA, 5, 5])
    ham = O.ham_desc(O.HAM_ELECTRONIC, [[0, 0, 0]], [10])
    opts = O.run_options(O.METROP_DIFFUSE, TAU, O.OBS_ENERGY | O.OBS_PGRAD | O.OBS_WFVALUE)
    cfgs = np.array([O.init_normal(SEED, w, 10, 0.5) for w in range(W)])
    times = []
    threads = 1
    for it in range(args.warmup + args.steps):
        secs, esum, threads = O.bench_vmc(wf, ham, opts, cfgs, sweeps, BLOCK, O.derive_seed(SEED, it))
        if it >= args.warmup:
            times.append(secs)
    total = sum(times)
    value = W * sweeps * args.steps / total
    line = {"impl": "reference", "metric": "vmc_walker_steps_per_sec_incl_local_energy", "value": value, "unit": "walker-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": "walker-steps/s", "cores": threads, "kind": "port",
                             "sample": "%d walkers x %d sweeps per step (reference-faithful: 3 value+gradient evaluations per "
                                       "diffusion move, from-scratch determinants), OpenMP over walkers" % (W, sweeps)},
            "e2e": {"value": value, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "ne_slater_jastrow_vmc_sr (BASELINE.json configs[4]): Ne, 10 electrons (5 up, 5 dn), STO 1s/2s/2p Slater "
                        "determinants x Pade-polynomial e-e Jastrow, P=7, MetropolisDiffuse tau=%.3g, SR" % TAU,
            "walkers_per_gpu": args.walkers, "global_walkers": args.walkers * world, "sweeps_per_step": args.sweeps,
            "block_size": BLOCK, "sampled_sweeps_per_step": args.sweeps - BLOCK, "parallelism": "walkers sharded, dp%d" % world,
            "cache": "L2 flushed between timed steps (256 MiB write); walker state is read once per %d-sweep launch" % args.sweeps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mole_b200", choices=["mole_b200", "reference"])
    ap.add_argument("--walkers", type=int, default=1 << 17, help="walkers per GPU (2^20 over 8 GPUs)")
    ap.add_argument("--sweeps", type=int, default=200, help="sweeps per optimisation iteration (first block discarded)")
    # CPU legs: bounded samples of the same workload (~4e4 walker-steps/s on 16 cores):
    # reference arm ~1 s per step, cpu_baseline ~10 s in total
    ap.add_argument("--ref-walkers", type=int, default=1024)
    ap.add_argument("--ref-sweeps", type=int, default=40)
    ap.add_argument("--cpu-baseline-walkers", type=int, default=1024)
    ap.add_argument("--cpu-baseline-sweeps", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--equil-box-sweeps", type=int, default=200)
    ap.add_argument("--equil-diffuse-sweeps", type=int, default=50)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "mole_b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # NCCL_DEBUG=VERSION makes NCCL print its version banner on stdout, ahead of the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    import mole_b200 as m

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = m.Context(local_rank)
    if world > 1:
        idt = torch.zeros(m.ffi.NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.tensor(list(m.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(world, rank, bytes(idt.cpu().tolist()))
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    W = args.walkers
    wf = m.SlaterJastrow(5, 5, ZETA, JB, KAPPA, ctx=ctx)
    op = m.ElectronicHamiltonian.from_ions([[0, 0, 0]], [10], ctx=ctx)
    met = m.MetropolisDiffuse.from_rng(TAU, SEED)
    opt = m.StochasticReconfiguration(SR_STEP, 7)
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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    import ctypes as C
    lib = m.ffi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    kernel_ms = []

    def step(it, e2e, timed):
        if e2e:
            ens.set_configs(host_cfgs.numpy())                      # H2D of this step's walkers from pinned memory
        ens.reseed(m.derive_seed(SEED, it))                         # vmc.rs:59-61
        ens.acc_reset()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        m.ffi.check(lib.mole_runner_run(ens.handle, wf.handle, met.handle, op.handle, C.c_uint32(obs), C.c_uint32(0),
                                        C.c_int32(args.sweeps), C.c_int32(BLOCK), None, None, None, None, None), ctx.handle)
        e1.record(stream)
        if world > 1:
            ens.acc_allreduce()                                     # concatenate_worker_data, vmc.rs:108-130
        acc = ens.acc_get()                                         # D2H of the reduced moments
        e, err, accp, g = m.acc_finalize(acc)
        dp = opt.compute_parameter_update(wf.parameters(), acc)     # SR: S^-1 (-g/2), optimizers.rs:237-252
        wf.update_parameters(dp)
        if timed:
            kernel_ms.append((e0, e1))
        return e, err, accp

    def run(e2e):
        wf.set_parameters(params0)
        # the nvidia-smi sampler is started BEFORE the warm-up: its start-up takes the driver lock for
        # ~100 ms and must not land inside the timed region; only rows taken inside it are summarised
        rows, stop = [], threading.Event()
        th = threading.Thread(target=clocks_sampler, args=(stop, rows, local_rank), daemon=True)
        th.start()
        t_wait = time.time()
        while not rows and time.time() - t_wait < 5.0:
            time.sleep(0.05)
        for it in range(args.warmup):
            with torch.cuda.stream(stream):
                flush.zero_()                                       # also loads torch's fill kernel outside the timed region
            step(it, e2e, False)
        barrier()
        n0 = len(rows)
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        t0.record(stream)
        last = None
        for it in range(args.steps):
            with torch.cuda.stream(stream):
                flush.zero_()                                       # L2 flush between timed iterations
            tw = time.perf_counter()
            last = step(args.warmup + it, e2e, not e2e)
            if args.verbose and rank == 0:
                print("  step %d (%s): %.2f ms wall, E = %.5f +/- %.5f" % (it, "e2e" if e2e else "resident",
                                                                           1e3 * (time.perf_counter() - tw), last[0], last[1]), file=sys.stderr)
        t1.record(stream)
        barrier()
        ms = t0.elapsed_time(t1)
        n1 = len(rows)
        stop.set()
        th.join(timeout=2)
        rows = rows[n0:max(n1, n0 + 1)]
        launches = ctx.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, last, summarize_clocks(rows)

    ms, launches, last, clocks = run(False)
    kms = float(np.mean([a.elapsed_time(b) for a, b in kernel_ms]))
    if world > 1:
        t = torch.tensor([kms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kms = float(t.item())
    ms_e2e, _, _, _ = run(True)
    fp64_peak = ctx.fp64_peak_tflops()

    wsteps = W * world * args.sweeps * args.steps
    value = wsteps / (ms * 1e-3)
    value_e2e = wsteps / (ms_e2e * 1e-3)
    f_alg = FLOPS["ne_slater_jastrow_vmc_sr"]["f_alg_per_walker_step"]
    kernel_rate = W * args.sweeps / (kms * 1e-3)                    # per GPU, dominant kernel only
    achieved = kernel_rate * f_alg / 1e12
    nominal = FLOPS["fp64_nominal_tflops"]
    line = {
        "metric": "vmc_walker_steps_per_sec_incl_local_energy", "value": value, "unit": "walker-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "sr_opt_step_ms": ms / args.steps,
        "energy": {"value": last[0], "blocking_error": last[1], "acceptance": last[2]},
        "roofline": {"bound": "fp64", "kernel": "sj_sweep_kernel<DIFFUSE,OPT>", "achieved": achieved, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": FLOPS["ne_slater_jastrow_vmc_sr"].get("dram_bytes_per_launch"),
                     "peak_source": "DFMA chain measured live by mole_bench_fp64_peak (MEASURED_PEAKS.json has no fp64 entry); "
                                    "nominal 148 SM x 64 lanes x 2 x 1.965 GHz = %.1f" % nominal,
                     "frac_of_nominal": achieved / nominal, "f_alg_per_walker_step": f_alg,
                     "kernel_ms_per_launch": kms, "kernel_walker_steps_per_s_per_gpu": kernel_rate,
                     "frac_from_scratch_count": kernel_rate * FLOPS["ne_slater_jastrow_vmc_sr"]["f_from_scratch_per_walker_step"] / 1e12 / fp64_peak,
                     # NOT measured in this run: ncu counters of the same launch (profiles/r01b_sj_sweep_v8_*), for context
                     "ncu_reference": FLOPS["ne_slater_jastrow_vmc_sr"].get("ncu_executed_fp64")},
        "e2e": {"value": value_e2e, "unit": "walker-steps/s", "h2d_bytes_per_step": W * 30 * 8 * world,
                "d2h_bytes_per_step": 62 * 8 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
    }
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
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        del ens                                                     # before its context (the library also defers the free)
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
