// Runs the Slater-Jastrow sweep / evaluation kernels of mole_b200/csrc/mole_sj.cuh UNCHANGED on the host through
// tests/native/cuda_emu.h (one std::thread per CUDA thread).  Test infrastructure: checks the cooperative kernels'
// logic against the oracle without a GPU (tests/test_emu_sj.py builds and drives it).
//
//   sj_emu <in.bin> <out.bin>
// in : int64 hdr[8] = {W, n_up, n_dn, metrop (0 box, 1 diffuse), n_sweeps, n_discard, block_size, walker_offset},
//      double par[10] = {zeta1..3, b1..4, kappa, Z, metrop_param}, uint8 seed[32], uint32 compat, uint32 pad,
//      double cfg[W][ne][3]
// out: double cfg[W][ne][3], double acc[ACC_DEV_LEN], double energy[ns][W], double wfvalue[ns][W],
//      double pgrad[ns][7][W], uint8 accept[n_sweeps][ne][W], then the sj_eval_kernel outputs of the INPUT configs:
//      double psi[W], grad[W][ne][3], lap[W], hpsi[W], pgrad[W][7]
#define MOLE_EMU 1
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>

#include "mole_kernels.cuh"
#include "mole_sj.cuh"

RngKey mole_key_from_seed(const uint8_t seed[32]) {   // mole_host.cpp (the Philox key of a 32-byte seed)
  uint32_t s[8];
  for (int i = 0; i < 8; ++i)
    s[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
  return RngKey{s[0] ^ s[2] ^ s[4] ^ s[6], s[1] ^ s[3] ^ s[5] ^ s[7]};
}

template <class T>
static void rd(FILE* f, T* p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int64_t hdr[8];
  double par[10];
  uint8_t seed[32];
  uint32_t cw[2];
  rd(f, hdr, 8); rd(f, par, 10); rd(f, seed, 32); rd(f, cw, 2);
  const int64_t W = hdr[0];
  const int nup = (int)hdr[1], ndn = (int)hdr[2], metrop = (int)hdr[3], ns_all = (int)hdr[4], ndisc = (int)hdr[5], bs = (int)hdr[6];
  const int ne = nup + ndn, n = 3 * ne;
  const int64_t nsamp = ns_all - ndisc;
  std::vector<double> aos((size_t)W * n), soa((size_t)W * n);
  rd(f, aos.data(), aos.size());
  fclose(f);
  for (int64_t w = 0; w < W; ++w)
    for (int c = 0; c < n; ++c) soa[(size_t)c * W + w] = aos[(size_t)w * n + c];
  const std::vector<double> soa_in = soa;

  SweepParams sp;
  memset(&sp, 0, sizeof(sp));
  const int blocks = (int)((W + SJ_WPB - 1) / SJ_WPB);
  std::vector<double> blk((size_t)W, 0.0), acc(ACC_DEV_LEN, 0.0), partials((size_t)blocks * ACC_LEN, 0.0);
  std::vector<double> tr_e((size_t)std::max<int64_t>(nsamp, 1) * W), tr_psi(tr_e.size()), tr_pg(tr_e.size() * SJ_NP);
  std::vector<uint8_t> tr_acc((size_t)ns_all * ne * W);
  unsigned int ticket = 0;
  sp.x = soa.data(); sp.blk = blk.data(); sp.acc = acc.data(); sp.partials = partials.data(); sp.ticket = &ticket;
  sp.W = W; sp.walker_offset = (uint64_t)hdr[7]; sp.key = mole_key_from_seed(seed); sp.step0 = 0;
  sp.n_sweeps = ns_all; sp.n_discard = ndisc; sp.block_size = bs; sp.blk_fill = 0;
  sp.observables = MOLE_OBS_ENERGY | MOLE_OBS_PGRAD | MOLE_OBS_WFVALUE; sp.compat = cw[0]; sp.metrop_param = par[9];
  sp.tr_energy = tr_e.data(); sp.tr_wfvalue = tr_psi.data(); sp.tr_pgrad = tr_pg.data(); sp.tr_accept = tr_acc.data();
  sp.wf.kind = MOLE_WF_SLATER_JASTROW; sp.wf.ne = ne; sp.wf.np = 7;
  for (int i = 0; i < 7; ++i) sp.wf.p[i] = par[i];
  sp.wf.geom[0] = par[7]; sp.wf.geom[1] = nup; sp.wf.geom[2] = ndn;
  sp.ham.kind = MOLE_OP_ELECTRONIC; sp.ham.n_ions = 1; sp.ham.ion_z[0] = par[8]; sp.ham.ionic_repulsion = 0.0;

  // batched evaluation of the input configurations first
  std::vector<double> psi((size_t)W), grad((size_t)W * n), lap((size_t)W), hpsi((size_t)W), pg((size_t)W * SJ_NP);
  emu::launch((unsigned)blocks, SJ_THREADS, SJ_SMEM_BYTES, [&] {
    sj_eval_kernel(soa_in.data(), W, sp.wf, sp.ham, 1, psi.data(), grad.data(), lap.data(), hpsi.data(), pg.data());
  });
  emu::launch((unsigned)blocks, SJ_THREADS, SJ_SMEM_BYTES, [&] {
    if (metrop == MOLE_METROP_BOX) sj_sweep_kernel<MOLE_METROP_BOX, true>(sp);
    else sj_sweep_kernel<MOLE_METROP_DIFFUSE, true>(sp);
  });

  for (int64_t w = 0; w < W; ++w)
    for (int c = 0; c < n; ++c) aos[(size_t)w * n + c] = soa[(size_t)c * W + w];
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 2;
  fwrite(aos.data(), 8, aos.size(), o);
  fwrite(acc.data(), 8, acc.size(), o);
  fwrite(tr_e.data(), 8, (size_t)nsamp * W, o);
  fwrite(tr_psi.data(), 8, (size_t)nsamp * W, o);
  fwrite(tr_pg.data(), 8, (size_t)nsamp * W * SJ_NP, o);
  fwrite(tr_acc.data(), 1, tr_acc.size(), o);
  fwrite(psi.data(), 8, psi.size(), o);
  fwrite(grad.data(), 8, grad.size(), o);
  fwrite(lap.data(), 8, lap.size(), o);
  fwrite(hpsi.data(), 8, hpsi.size(), o);
  fwrite(pg.data(), 8, pg.size(), o);
  fclose(o);
  printf("ok %lld walkers %d sweeps %d blocks\n", (long long)W, ns_all, blocks);
  return 0;
}
