"""The general LCAO Slater-Jastrow kernels (mole_b200/csrc/mole_lsj.cuh) compiled unchanged by g++ and run on the host
through tests/native/cuda_emu.h against the oracle (see tests/test_emu_sj.py): psi, grad, lap, E_L, O_k, accept/reject
bits, the per-sample rows that feed the Gram matrix.  CPU only."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from common import cases, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lsj_emu(tmp_path_factory):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("emu") / "lsj_emu")
    subprocess.check_call([cxx, "-O1", "-std=c++17", "-pthread", "-mfma", "-ffp-contract=off",
                           "-I", os.path.join(ROOT, "mole_b200", "csrc"), "-I", os.path.join(ROOT, "tests", "native"),
                           os.path.join(ROOT, "tests", "native", "lsj_emu.cpp"), "-o", exe])
    return exe


def run_lsj(exe, tmp, c, cfgs, metrop, param, steps, ndisc, bs, seed, offset=0):
    W, ne = cfgs.shape[0], cfgs.shape[1]
    owf, oham = c["owf"], c["oham"]
    P = owf.n_params
    pos = np.array(oham.ion_pos)[:24]
    z = np.array([float(v) for v in oham.ion_charge])[:8]
    rep = sum(z[i] * z[j] / np.linalg.norm(pos[3 * i:3 * i + 3] - pos[3 * j:3 * j + 3])
              for i in range(oham.n_ions) for j in range(i + 1, oham.n_ions))
    inp, out = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<8q", W, P, metrop, steps, ndisc, bs, offset, oham.n_ions))
        f.write(struct.pack("<48d", *list(owf.params)))
        f.write(struct.pack("<40d", *list(owf.geom)))
        f.write(struct.pack("<24d", *pos))
        f.write(struct.pack("<8d", *z))
        f.write(struct.pack("<2d", rep, param))
        f.write(bytes(seed))
        f.write(np.ascontiguousarray(cfgs, dtype=np.float64).tobytes())
    r = subprocess.run([exe, inp, out], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    buf = open(out, "rb").read()
    ns, pos_ = steps - ndisc, 0

    def take(n, dt=np.float64):
        nonlocal pos_
        a = np.frombuffer(buf, dtype=dt, count=n, offset=pos_)
        pos_ += a.nbytes
        return a
    res = dict(cfgs=take(W * ne * 3).reshape(W, ne, 3), acc=take(64),
               energy=take(ns * W).reshape(ns, W).T, wfvalue=take(ns * W).reshape(ns, W).T,
               pgrad=np.moveaxis(take(ns * P * W).reshape(ns, P, W), -1, 0),
               accept=np.moveaxis(take(steps * ne * W, np.uint8).reshape(steps, ne, W), -1, 0),
               rows=np.moveaxis(take(ns * (P + 2) * W).reshape(ns, P + 2, W), -1, 0),
               psi=take(W), grad=take(W * ne * 3).reshape(W, ne, 3), lap=take(W), hpsi=take(W), pg=take(W * P).reshape(W, P))
    assert pos_ == len(buf)
    return res


def close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-3 * np.max(np.abs(b)))
    return np.max(np.abs(a - b) / scale) < tol


@pytest.mark.parametrize("name,metrop", [("lsj_h4", "diffuse"), ("lsj_h4", "box"), ("lsj_h3", "diffuse"), ("lsj_h8", "diffuse")])
def test_emulated_lsj_kernels_match_oracle(orc, lsj_emu, tmp_path, name, metrop):
    c = cases()[name]
    ne, P = c["ne"], c["np"]
    W, steps, bs = 70, 20, 5                                     # 70 walkers: one full CTA of 64 and a ragged one
    seed = bytes([7] * 32)
    cfgs = np.array([orc.init_uniform(seed, 100 + w, ne, -1.5, 1.5) for w in range(W)])
    kind = orc.METROP_BOX if metrop == "box" else orc.METROP_DIFFUSE
    param = 0.6 if metrop == "box" else 0.05
    got = run_lsj(lsj_emu, str(tmp_path), c, cfgs, kind, param, steps, bs, bs, seed, offset=100)
    ref = orc.eval_batch(c["owf"], c["oham"], cfgs)
    assert rel_err(got["psi"], ref["psi"]) < 1e-10
    assert close(got["grad"], ref["grad"], 1e-10) and close(got["lap"], ref["lap"], 1e-10)
    assert close(got["hpsi"] / got["psi"], ref["hpsi"] / ref["psi"], 1e-10) and close(got["pg"], ref["pgrad"], 1e-10)
    obs = orc.OBS_ENERGY | orc.OBS_WFVALUE | orc.OBS_PGRAD
    r = orc.ensemble_run(c["owf"], c["oham"], orc.run_options(kind, param, obs, nan_reject=1), cfgs, seed, steps, bs, walker_offset=100)
    assert np.array_equal(got["accept"], r["accept"]), "accept/reject decisions differ"
    assert close(got["cfgs"], r["cfgs"], 1e-10)
    assert close(got["energy"], r["energy"], 1e-9) and close(got["wfvalue"], r["wfvalue"], 1e-10)
    assert close(got["pgrad"], r["pgrad"][:, :, :P], 1e-9)
    o = r["pgrad"][:, :, :P] / r["wfvalue"][:, :, None]
    assert np.all(got["rows"][:, :, 0] == 1.0) and close(got["rows"][:, :, 1], r["energy"], 1e-9) and close(got["rows"][:, :, 2:], o, 1e-9)
    acc = got["acc"]
    assert acc[0] == r["energy"].size and acc[6] == r["accept"].sum() and acc[62] == 0
