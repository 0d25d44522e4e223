"""The cooperative Slater-Jastrow kernels (mole_b200/csrc/mole_sj.cuh, mole_sj_move.cuh) compiled UNCHANGED by g++
and run on the host through tests/native/cuda_emu.h (one std::thread per CUDA thread, barriers for __syncwarp /
__syncthreads, shuffles through a scratch row) against the oracle: accept/reject bit-exact, psi / grad / lap / E_L /
O_k within 1e-10, accumulators against the traces.  CPU only - it checks the device LOGIC (mailbox protocol, lane
ownership, synchronisation placement, reductions, ragged last chunk) before GPU minutes are spent; the parity tests
proper are tests/test_gpu_parity.py.  The emulated MUFU seeds differ from the device's in the last bits, nothing else."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from common import cases, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACC_DEV_LEN = 64


@pytest.fixture(scope="module")
def sj_emu(tmp_path_factory):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("emu") / "sj_emu")
    subprocess.check_call([cxx, "-O1", "-std=c++17", "-pthread", "-mfma", "-ffp-contract=off", "-DMOLE_SJ_WARPS=1",
                           "-I", os.path.join(ROOT, "mole_b200", "csrc"), "-I", os.path.join(ROOT, "tests", "native"),
                           os.path.join(ROOT, "tests", "native", "sj_emu.cpp"), "-o", exe])
    return exe


def run_emu(exe, tmp, c, cfgs, metrop, param, steps, ndisc, bs, seed, compat=0, offset=0):
    W, ne = cfgs.shape[0], cfgs.shape[1]
    owf = c["owf"]
    nup, ndn = int(owf.geom[1]), int(owf.geom[2])
    par = list(owf.params[:7]) + [owf.geom[0], float(c["oham"].ion_charge[0]), param]
    inp, out = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<8q", W, nup, ndn, metrop, steps, ndisc, bs, offset))
        f.write(struct.pack("<10d", *par))
        f.write(bytes(seed))
        f.write(struct.pack("<2I", compat, 0))
        f.write(np.ascontiguousarray(cfgs, dtype=np.float64).tobytes())
    r = subprocess.run([exe, inp, out], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    buf = open(out, "rb").read()
    ns, pos = steps - ndisc, 0

    def take(n, dt=np.float64):
        nonlocal pos
        a = np.frombuffer(buf, dtype=dt, count=n, offset=pos)
        pos += a.nbytes
        return a
    res = dict(cfgs=take(W * ne * 3).reshape(W, ne, 3), acc=take(ACC_DEV_LEN),
               energy=take(ns * W).reshape(ns, W).T, wfvalue=take(ns * W).reshape(ns, W).T,
               pgrad=np.moveaxis(take(ns * 7 * W).reshape(ns, 7, W), -1, 0),
               accept=np.moveaxis(take(steps * ne * W, np.uint8).reshape(steps, ne, W), -1, 0),
               psi=take(W), grad=take(W * ne * 3).reshape(W, ne, 3), lap=take(W), hpsi=take(W), pg=take(W * 7).reshape(W, 7))
    assert pos == len(buf)
    return res


def close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b)))
    return np.max(np.abs(a - b) / scale) < tol


@pytest.mark.parametrize("name,metrop", [("sj_ne", "diffuse"), ("sj_ne", "box"), ("sj_be", "diffuse"), ("sj_li", "diffuse"), ("sj_li", "box")])
def test_emulated_sj_kernels_match_oracle(orc, sj_emu, tmp_path, name, metrop):
    c = cases()[name]
    ne = c["ne"]
    W, steps, bs = 14, 20, 5                                     # 14 walkers: chunks of 6, 6 and a ragged 2
    seed = bytes([7] * 32)
    cfgs = np.array([orc.init_uniform(seed, 100 + w, ne) for w in range(W)])
    kind = orc.METROP_BOX if metrop == "box" else orc.METROP_DIFFUSE
    param = 0.4 if metrop == "box" else 0.02
    got = run_emu(sj_emu, str(tmp_path), c, cfgs, kind, param, steps, bs, bs, seed, offset=100)
    # batched evaluation
    ref = orc.eval_batch(c["owf"], c["oham"], cfgs)
    assert rel_err(got["psi"], ref["psi"]) < 1e-10
    assert close(got["grad"], ref["grad"], 1e-10) and close(got["lap"], ref["lap"], 1e-10)
    assert close(got["hpsi"] / got["psi"], ref["hpsi"] / ref["psi"], 1e-10) and close(got["pg"], ref["pgrad"], 1e-10)
    # fused sweep with a shared Philox stream
    obs = orc.OBS_ENERGY | orc.OBS_WFVALUE | orc.OBS_PGRAD
    r = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(kind, param, obs, nan_reject=1), cfgs, seed, steps, bs, walker_offset=100)
    assert np.array_equal(got["accept"], r["accept"]), "accept/reject decisions differ"
    assert close(got["cfgs"], r["cfgs"], 1e-10)
    assert close(got["energy"], r["energy"], 1e-9) and close(got["wfvalue"], r["wfvalue"], 1e-10)
    assert close(got["pgrad"], r["pgrad"][:, :, :7], 1e-9)
    acc, en = got["acc"], r["energy"]
    assert acc[0] == en.size and acc[7] == W * steps * ne and acc[6] == r["accept"].sum() and acc[62] == 0
    assert abs(acc[1] - en.sum()) < 1e-9 * np.abs(en).sum()
    bm = en.reshape(W, -1, bs).mean(axis=2)
    assert acc[5] == bm.size and abs(acc[4] - (bm ** 2).sum()) < 1e-9 * (bm ** 2).sum()
    o = r["pgrad"][:, :, :7] / r["wfvalue"][:, :, None]
    for k in range(7):
        assert abs(acc[10 + k] - o[:, :, k].sum()) < 1e-9 * np.abs(o[:, :, k]).sum() + 1e-300
        assert abs(acc[18 + k] - (o[:, :, k] * en).sum()) < 1e-9 * np.abs(o[:, :, k] * en).sum() + 1e-300
    q = 0
    for k in range(7):
        for l in range(k, 7):
            s = (o[:, :, k] * o[:, :, l]).sum()
            assert abs(acc[26 + q] - s) < 1e-9 * np.abs(o[:, :, k] * o[:, :, l]).sum() + 1e-300
            q += 1


@pytest.mark.parametrize("metrop", ["box", "diffuse"])
def test_emulated_sj_kernel_with_nonfinite_walkers(orc, sj_emu, tmp_path, metrop):
    """Walkers whose psi / E_L are NaN (an electron ON the nucleus) next to healthy ones in the same warp: every
    warp-wide vote and mailbox reduction must still be reached by all lanes (a short-circuited __ballot_sync hung the
    GPU in round 2's first run - the emulator's watchdog reports the diverging synchronisation sites), the NaN samples
    are counted and skipped, the healthy walkers' sums are untouched."""
    c = cases()["sj_be"]
    W, steps, bs = 14, 20, 10
    cfgs = np.random.default_rng(2).normal(0.0, 0.8, size=(W, 4, 3))
    cfgs[:3, 0, :] = 0.0
    os.environ["MOLE_EMU_WATCHDOG"] = "60"
    kind, param = (orc.METROP_BOX, 1e-200) if metrop == "box" else (orc.METROP_DIFFUSE, 0.02)
    got = run_emu(sj_emu, str(tmp_path), c, cfgs, kind, param, steps, 0, bs, bytes(32))
    en, acc = got["energy"], got["acc"]
    assert not np.isfinite(en[:3]).any() and np.isfinite(en[3:]).all()
    assert acc[0] == (W - 3) * steps and acc[62] == 3 * steps and acc[5] == (W - 3) * steps // bs
    assert abs(acc[1] - en[3:].sum()) < 1e-9 * np.abs(en[3:]).sum()
    assert np.isfinite(acc[:62]).all()
