// Host emulation harness for the general LCAO Slater-Jastrow kernels (mole_b200/csrc/mole_lsj.cuh), see sj_emu.cpp.
//   lsj_emu <in.bin> <out.bin>
// in : int64 hdr[8] = {W, n_params, metrop, n_sweeps, n_discard, block_size, walker_offset, n_ions},
//      double params[48], double geom[40], double ion_pos[24], double ion_z[8], double ionic_repulsion, double metrop_param,
//      uint8 seed[32], double cfg[W][ne][3]
// out: cfg[W][ne][3], acc[ACC_DEV_LEN], energy[ns][W], wfvalue[ns][W], pgrad[ns][P][W], accept[n_sweeps][ne][W] (uint8),
//      rows[ns][P+2][W], then eval of the INPUT configs: psi[W], grad[W][ne][3], lap[W], hpsi[W], pgrad[W][P]
#define MOLE_EMU 1
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>

#include "mole_kernels.cuh"
#include "mole_lsj.cuh"

RngKey mole_key_from_seed(const uint8_t seed[32]) {
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
  SweepParams sp;
  memset(&sp, 0, sizeof(sp));
  uint8_t seed[32];
  rd(f, hdr, 8);
  rd(f, sp.wf.p, 48); rd(f, sp.wf.geom, 40); rd(f, sp.ham.ion_pos, 24); rd(f, sp.ham.ion_z, 8);
  rd(f, &sp.ham.ionic_repulsion, 1); rd(f, &sp.metrop_param, 1); rd(f, seed, 32);
  const int64_t W = hdr[0];
  const int P = (int)hdr[1], metrop = (int)hdr[2], ns_all = (int)hdr[3], ndisc = (int)hdr[4], bs = (int)hdr[5];
  const int ne = (int)sp.wf.geom[1] + (int)sp.wf.geom[2], n = 3 * ne, cols = P + 2;
  const int64_t nsamp = ns_all - ndisc;
  std::vector<double> aos((size_t)W * n), soa((size_t)W * n);
  rd(f, aos.data(), aos.size());
  fclose(f);
  for (int64_t w = 0; w < W; ++w)
    for (int c = 0; c < n; ++c) soa[(size_t)c * W + w] = aos[(size_t)w * n + c];
  const std::vector<double> soa_in = soa;
  sp.wf.kind = MOLE_WF_LCAO_SJ; sp.wf.ne = ne; sp.wf.np = P;
  sp.ham.kind = MOLE_OP_ELECTRONIC; sp.ham.n_ions = (int)hdr[7];
  const int blocks = (int)((W + LSJ_THREADS - 1) / LSJ_THREADS);
  std::vector<double> blk((size_t)W, 0.0), acc(ACC_DEV_LEN, 0.0), partials((size_t)blocks * ACC_LEN, 0.0);
  const size_t nsw = (size_t)std::max<int64_t>(nsamp, 1) * W;
  std::vector<double> tr_e(nsw), tr_psi(nsw), tr_pg(nsw * P), rows(nsw * cols);
  std::vector<uint8_t> tr_acc((size_t)ns_all * ne * W);
  unsigned int ticket = 0;
  sp.x = soa.data(); sp.blk = blk.data(); sp.acc = acc.data(); sp.partials = partials.data(); sp.ticket = &ticket;
  sp.W = W; sp.walker_offset = (uint64_t)hdr[6]; sp.key = mole_key_from_seed(seed); sp.step0 = 0;
  sp.n_sweeps = ns_all; sp.n_discard = ndisc; sp.block_size = bs; sp.blk_fill = 0;
  sp.observables = MOLE_OBS_ENERGY | MOLE_OBS_PGRAD | MOLE_OBS_WFVALUE;
  sp.tr_energy = tr_e.data(); sp.tr_wfvalue = tr_psi.data(); sp.tr_pgrad = tr_pg.data(); sp.tr_accept = tr_acc.data();
  sp.osamp = rows.data();
  std::vector<double> psi((size_t)W), grad((size_t)W * n), lap((size_t)W), hpsi((size_t)W), pg((size_t)W * P);
  emu::launch((unsigned)blocks, LSJ_THREADS, 0, [&] {
    lsj_eval_kernel(soa_in.data(), W, sp.wf, sp.ham, 1, psi.data(), grad.data(), lap.data(), hpsi.data(), pg.data());
  });
  emu::launch((unsigned)blocks, LSJ_THREADS, 0, [&] {
    if (metrop == MOLE_METROP_BOX) lsj_sweep_kernel<MOLE_METROP_BOX, true>(sp);
    else lsj_sweep_kernel<MOLE_METROP_DIFFUSE, true>(sp);
  });
  for (int64_t w = 0; w < W; ++w)
    for (int c = 0; c < n; ++c) aos[(size_t)w * n + c] = soa[(size_t)c * W + w];
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 2;
  fwrite(aos.data(), 8, aos.size(), o);
  fwrite(acc.data(), 8, acc.size(), o);
  fwrite(tr_e.data(), 8, (size_t)nsamp * W, o);
  fwrite(tr_psi.data(), 8, (size_t)nsamp * W, o);
  fwrite(tr_pg.data(), 8, (size_t)nsamp * W * P, o);
  fwrite(tr_acc.data(), 1, tr_acc.size(), o);
  fwrite(rows.data(), 8, (size_t)nsamp * W * cols, o);
  fwrite(psi.data(), 8, psi.size(), o);
  fwrite(grad.data(), 8, grad.size(), o);
  fwrite(lap.data(), 8, lap.size(), o);
  fwrite(hpsi.data(), 8, hpsi.size(), o);
  fwrite(pg.data(), 8, pg.size(), o);
  fclose(o);
  printf("ok %lld walkers %d sweeps\n", (long long)W, ns_all);
  return 0;
}
