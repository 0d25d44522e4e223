// ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path (mole_b200/).
// C ABI (ctypes) over the templates in oracle_{rng,wf,mc}.hpp.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load this.
#include <cstring>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_mc.hpp"
#include "oracle_stat.hpp"

using namespace orc;

#define ORC_TRY try {
#define ORC_CATCH(rv) } catch (const std::exception&) { return rv; }

extern "C" {

// ---- RNG
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = r.w[i];
}
void orc_key_from_seed(const uint8_t seed[32], uint32_t out[2]) {
  const Key k = key_from_seed(seed);
  out[0] = k.k0; out[1] = k.k1;
}
void orc_derive_seed(const uint8_t master[32], uint32_t n, uint8_t out[32]) { derive_seed(master, n, out); }
// kind 0: four uniforms; kind 1: three normals + one uniform
void orc_draw_move(int kind, const uint8_t seed[32], uint64_t walker, uint32_t step, uint32_t domain,
                   uint32_t elec, double out[4]) {
  const Key k = key_from_seed(seed);
  const MoveDraw d = kind == 0 ? draw_uniform4(k, walker, step, (Domain)domain, elec)
                               : draw_normal3_uniform1(k, walker, step, (Domain)domain, elec);
  out[0] = d.a; out[1] = d.b; out[2] = d.c; out[3] = d.u;
}
void orc_init_uniform(const uint8_t seed[32], uint64_t walker, int ne, double lo, double hi, double* cfg) {
  init_uniform(key_from_seed(seed), walker, ne, lo, hi, cfg);
}
void orc_init_normal(const uint8_t seed[32], uint64_t walker, int ne, double sigma, double* cfg) {
  init_normal(key_from_seed(seed), walker, ne, sigma, cfg);
}

// ---- pointwise wavefunction / operator evaluation (Function, Differentiate, Optimize, LocalOperator)
int orc_wf_value(const WfDesc* d, const double* cfg, double* out) {
  ORC_TRY *out = wf_value(Wf<double>(*d), cfg); return 0; ORC_CATCH(3)
}
int orc_wf_gradient(const WfDesc* d, const double* cfg, double* out) {
  ORC_TRY wf_gradient(Wf<double>(*d), cfg, out); return 0; ORC_CATCH(3)
}
int orc_wf_laplacian(const WfDesc* d, const double* cfg, double* out) {
  ORC_TRY *out = wf_laplacian(Wf<double>(*d), cfg); return 0; ORC_CATCH(3)
}
int orc_wf_parameter_gradient(const WfDesc* d, const double* cfg, double* out) {
  ORC_TRY wf_parameter_gradient(Wf<double>(*d), cfg, out); return 0; ORC_CATCH(3)
}
int orc_ham_act_on(const HamDesc* h, const WfDesc* d, const double* cfg, double* out) {
  ORC_TRY *out = ham_act_on(Ham<double>(*h), Wf<double>(*d), cfg); return 0; ORC_CATCH(3)
}
double orc_ionic_potential(const HamDesc* h, const double* cfg, int ne) {
  return ionic_potential(Ham<double>(*h), cfg, ne);
}
double orc_electronic_potential(const double* cfg, int ne) { return electronic_potential(cfg, ne); }

// batched pointwise evaluation: psi[W], grad[W*n], lap[W], hpsi[W], pgrad[W*P]  (any pointer may be NULL)
int orc_eval_batch(const WfDesc* d, const HamDesc* h, const double* cfgs, int64_t W, double* psi,
                   double* grad, double* lap, double* hpsi, double* pgrad) {
  ORC_TRY
  const Wf<double> wf(*d);
  const int n = 3 * wf.ne;
  int bad = 0;
#pragma omp parallel for schedule(static)
  for (int64_t w = 0; w < W; ++w) {
    try {
      const double* c = cfgs + w * n;
      if (psi) psi[w] = wf_value(wf, c);
      if (grad) wf_gradient(wf, c, grad + w * n);
      if (lap) lap[w] = wf_laplacian(wf, c);
      if (hpsi && h) hpsi[w] = ham_act_on(Ham<double>(*h), wf, c);
      if (pgrad) wf_parameter_gradient(wf, c, pgrad + w * wf.np);
    } catch (...) { bad = 1; }
  }
  return bad ? 3 : 0;
  ORC_CATCH(3)
}

// ---- one Metropolis::move_state; cfg is updated in place when accepted.  Returns 1/0, <0 on error.
int orc_move_state(const WfDesc* d, int metrop_kind, double param, double* cfg, int idx,
                   const uint8_t seed[32], uint64_t walker, uint32_t step, double* ratio_out) {
  ORC_TRY
  const Wf<double> wf(*d);
  const Key k = key_from_seed(seed);
  if (metrop_kind == METROP_BOX) return box_move_state(wf, cfg, idx, param, k, walker, step) ? 1 : 0;
  return diffuse_move_state(wf, cfg, idx, param, k, walker, step, ratio_out) ? 1 : 0;
  ORC_CATCH(-3)
}

// ---- Runner::run over an ensemble of W independent chains (walker ids walker_offset..+W).
// cfgs: W*n in/out.  Per-sample outputs (nullable): energy[W*ns], wfvalue[W*ns], kinetic[W*ns],
// pgrad[W*ns*P]; accept[W*steps*ne]; acceptance[W].  ns = (steps/block_size - 1)*block_size.
int orc_ensemble_run(const WfDesc* d, const HamDesc* h, const RunOptions* o, double* cfgs,
                     const uint8_t seed[32], uint64_t walker_offset, int64_t W, int steps, int block_size,
                     double* energy, double* wfvalue, double* kinetic, double* pgrad, uint8_t* accept,
                     double* acceptance) {
  ORC_TRY
  if (!(steps >= 2 * block_size)) return 10;
  const Wf<double> wf(*d);
  const Ham<double> ham(*h);
  const Key key = key_from_seed(seed);
  const int n = 3 * wf.ne, np = wf.np;
  const int64_t ns = (int64_t)(steps / block_size - 1) * block_size;
  const int64_t nsteps_eff = (int64_t)(steps / block_size) * block_size;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t w = 0; w < W; ++w) {
    try {
      RunResult r = runner_run(wf, ham, *o, cfgs + w * n, key, walker_offset + (uint64_t)w, steps,
                               block_size, accept != nullptr);
      std::copy(r.cfg.begin(), r.cfg.end(), cfgs + w * n);
      if (energy && !r.energy.empty()) std::copy(r.energy.begin(), r.energy.end(), energy + w * ns);
      if (wfvalue && !r.wfvalue.empty()) std::copy(r.wfvalue.begin(), r.wfvalue.end(), wfvalue + w * ns);
      if (kinetic && !r.kinetic.empty()) std::copy(r.kinetic.begin(), r.kinetic.end(), kinetic + w * ns);
      if (pgrad && !r.pgrad.empty()) std::copy(r.pgrad.begin(), r.pgrad.end(), pgrad + w * ns * np);
      if (accept) std::copy(r.accept.begin(), r.accept.end(), accept + w * nsteps_eff * wf.ne);
      if (acceptance) acceptance[w] = r.acceptance;
    } catch (...) { bad = 1; }
  }
  return bad ? 3 : 0;
  ORC_CATCH(3)
}

// ---- statistics
double orc_mean_fold(const double* v, int64_t n) { return mean_fold(v, (size_t)n); }
double orc_blocking_error(const double* v, int64_t n, int64_t block_size, double mean) {
  return blocking_error(v, (size_t)n, (size_t)block_size, mean);
}

// ---- optimizers
void* orc_opt_create(int kind, int np, double step, double momentum, int history, int quirk_sr_subtract) {
  return new Optimizer(kind, np, step, momentum, history, quirk_sr_subtract);
}
void orc_opt_destroy(void* p) { delete (Optimizer*)p; }
// raw sample series exactly as the reference optimizers receive them
int orc_opt_step(void* p, const double* pars, double energy_avg, const double* wfv, const double* pg,
                 const double* en, int64_t ns, double* deltap) {
  ORC_TRY
  Optimizer* o = (Optimizer*)p;
  std::vector<double> vp(pars, pars + o->np), vw(wfv, wfv + ns), vg(pg, pg + ns * o->np), ve(en, en + ns), dp;
  if (!o->compute_parameter_update(vp, energy_avg, vw, vg, ve, dp)) return 1;
  std::copy(dp.begin(), dp.end(), deltap);
  return 0;
  ORC_CATCH(3)
}
int orc_energy_gradient(const double* wfv, const double* pg, const double* en, int64_t ns, int np,
                        double energy, double* g) {
  std::vector<double> vw(wfv, wfv + ns), vg(pg, pg + ns * np), ve(en, en + ns);
  const std::vector<double> r = energy_gradient(vw, vg, ve, energy, np);
  std::copy(r.begin(), r.end(), g);
  return 0;
}
int orc_sr_matrix(void* p, const double* wfv, const double* pg, int64_t ns, double* S) {
  Optimizer* o = (Optimizer*)p;
  std::vector<double> vw(wfv, wfv + ns), vg(pg, pg + ns * o->np);
  const std::vector<double> r = o->sr_matrix(vg, vw);
  std::copy(r.begin(), r.end(), S);
  return 0;
}

// ---- VmcRunner::run_optimization
int orc_vmc_run_optimization(const WfDesc* d, const HamDesc* h, const RunOptions* o, void* opt,
                             const uint8_t master_seed[32], const double* cfg0, int iters,
                             int total_samples, int block_size, int nworkers, double* energies,
                             double* errors, double* acceptance, double* param_history,
                             double* final_params) {
  ORC_TRY
  VmcResult r = vmc_run_optimization(*d, *h, *o, *(Optimizer*)opt, master_seed, cfg0, iters, total_samples,
                                     block_size, nworkers);
  std::copy(r.energies.begin(), r.energies.end(), energies);
  std::copy(r.errors.begin(), r.errors.end(), errors);
  if (acceptance) std::copy(r.acceptance.begin(), r.acceptance.end(), acceptance);
  if (param_history) std::copy(r.param_history.begin(), r.param_history.end(), param_history);
  if (final_params) std::copy(r.params.begin(), r.params.end(), final_params);
  return 0;
  ORC_CATCH(3)
}

// ---- DMC
// one time step (no branching): weights[N], cfgs[N*n] in/out; returns ensemble energy
int orc_dmc_step(const WfDesc* d, const HamDesc* h, double* weights, double* cfgs, int64_t N,
                 double metrop_tau, double time_step, double e_ref, const uint8_t seed[32], uint32_t step,
                 double* ens_energy, double* total_weight) {
  ORC_TRY
  const Wf<double> wf(*d);
  Walkers wk;
  wk.ne = wf.ne;
  wk.w.assign(weights, weights + N);
  wk.cfg.assign(cfgs, cfgs + N * 3 * wf.ne);
  *ens_energy = dmc_step(wf, Ham<double>(*h), wk, metrop_tau, time_step, e_ref, key_from_seed(seed), step,
                         total_weight);
  std::copy(wk.w.begin(), wk.w.end(), weights);
  std::copy(wk.cfg.begin(), wk.cfg.end(), cfgs);
  return 0;
  ORC_CATCH(3)
}
// branching: in (weights[N], cfgs[N*n]) -> out buffers sized N (both branchers conserve N). src_out nullable.
int orc_branch(int kind, int ne, const double* weights, const double* cfgs, int64_t N,
               const uint8_t seed[32], uint32_t step, double* weights_out, double* cfgs_out) {
  ORC_TRY
  Walkers wk;
  wk.ne = ne;
  wk.w.assign(weights, weights + N);
  wk.cfg.assign(cfgs, cfgs + N * 3 * ne);
  const Walkers o = kind == BRANCH_SR ? branch_sr(wk, key_from_seed(seed), step)
                                      : branch_simple(wk, key_from_seed(seed), step);
  if ((int64_t)o.size() != N) return 2;
  std::copy(o.w.begin(), o.w.end(), weights_out);
  std::copy(o.cfg.begin(), o.cfg.end(), cfgs_out);
  return 0;
  ORC_CATCH(3)
}
// DmcRunner::diffuse. energies/errors sized >= num_iterations/block_size; returns count via n_out.
int orc_dmc_diffuse(const WfDesc* d, const HamDesc* h, double* weights, double* cfgs, int64_t N,
                    double metrop_tau, double e_ref, int branch_kind, const uint8_t seed[32],
                    double time_step, int num_iterations, int block_size, int num_eq_blocks,
                    double* energies, double* errors, int* n_out, double* step_energies,
                    double* e_ref_out) {
  ORC_TRY
  Walkers wk;
  wk.ne = d->n_elec;
  wk.w.assign(weights, weights + N);
  wk.cfg.assign(cfgs, cfgs + N * 3 * d->n_elec);
  DmcResult r = dmc_diffuse(*d, *h, wk, metrop_tau, e_ref, branch_kind, key_from_seed(seed), time_step,
                            num_iterations, block_size, num_eq_blocks);
  std::copy(r.energies.begin(), r.energies.end(), energies);
  std::copy(r.errors.begin(), r.errors.end(), errors);
  *n_out = (int)r.energies.size();
  if (step_energies) std::copy(r.step_energies.begin(), r.step_energies.end(), step_energies);
  if (e_ref_out) *e_ref_out = r.reference_energy;
  std::copy(r.walkers.w.begin(), r.walkers.w.end(), weights);
  std::copy(r.walkers.cfg.begin(), r.walkers.cfg.end(), cfgs);
  return 0;
  ORC_CATCH(3)
}

// ---- op counts of the reference-faithful algorithm (one sweep + one sample), for DESIGN.md
int orc_count_ops_faithful(const WfDesc* d, const HamDesc* h, const RunOptions* o, const double* cfg,
                           uint64_t out[7]) {
  ORC_TRY
  const Wf<Counted> wf(*d);
  const Ham<Counted> ham(*h);
  const int n = 3 * wf.ne;
  std::vector<Counted> c(n);
  for (int i = 0; i < n; ++i) c[i] = Counted(cfg[i]);
  const uint8_t seed[32] = {0};
  const Key key = key_from_seed(seed);
  op_counts() = OpCounts();
  for (int e = 0; e < wf.ne; ++e) {
    if (o->metrop_kind == METROP_BOX) box_move_state(wf, c.data(), e, o->metrop_param, key, 0, 0);
    else diffuse_move_state(wf, c.data(), e, o->metrop_param, key, 0, 0);
  }
  if (o->observables & OBS_ENERGY) { Counted e = ham_act_on(ham, wf, c.data()) / wf_value(wf, c.data()); (void)e; }
  if (o->observables & OBS_PGRAD) {
    Counted pg[WF_MAX_PARAMS];
    wf_parameter_gradient(wf, c.data(), pg);
    Counted v = wf_value(wf, c.data());
    for (int k = 0; k < wf.np; ++k) { Counted t = (v * pg[k]) / v; (void)t; }
  }
  if (o->observables & OBS_WFVALUE) { Counted v = wf_value(wf, c.data()); Counted t = (v * v) / v; (void)t; }
  const OpCounts& k = op_counts();
  out[0] = k.add; out[1] = k.mul; out[2] = k.div; out[3] = k.sqrt; out[4] = k.exp; out[5] = k.log; out[6] = k.cmp;
  return 0;
  ORC_CATCH(3)
}

// ---- CPU baseline leg: W chains x steps sweeps with energy sampling every sweep after block 0,
// OpenMP over walkers (the analogue of the rayon fan-out, vmc.rs:63-76).  Returns seconds.
double orc_bench_vmc(const WfDesc* d, const HamDesc* h, const RunOptions* o, const double* cfgs, int64_t W,
                     int steps, int block_size, const uint8_t seed[32], double* energy_sum, int* threads) {
  const Wf<double> wf(*d);
  const Ham<double> ham(*h);
  const Key key = key_from_seed(seed);
  const int n = 3 * wf.ne;
#ifdef _OPENMP
  *threads = omp_get_max_threads();
#else
  *threads = 1;
#endif
  double total = 0.0;
  const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total)
  for (int64_t w = 0; w < W; ++w) {
    RunResult r = runner_run(wf, ham, *o, cfgs + w * n, key, (uint64_t)w, steps, block_size, false);
    double s = 0.0;
    for (double e : r.energy) s += e;
    total += s;
  }
  const auto t1 = std::chrono::steady_clock::now();
  *energy_sum = total;
  return std::chrono::duration<double>(t1 - t0).count();
}

// scripts/statfor.rs on one series: stats = {average, variance(ddof 1), tcorr, n_eff, sigma};
// corr[min(200, n-1)] and errs[n_sizes] nullable.  Returns the number of lags.
int orc_statfor(const double* data, int64_t n, double* stats, double* corr, int n_sizes, const int* sizes, double* errs) {
  const double average = stat_mean(data, n);
  const double var = stat_variance(data, n, 1);
  double tcorr, neff, sigma;
  const int lags = stat_correlation(data, n, average, var, corr, &tcorr, &neff, &sigma);
  stats[0] = average; stats[1] = var; stats[2] = tcorr; stats[3] = neff; stats[4] = sigma;
  for (int k = 0; k < n_sizes; ++k) errs[k] = stat_block_error(data, n, sizes[k]);
  return lags;
}
int orc_statfor_block_sizes(int64_t n, int* sizes) {
  const std::vector<int> v = stat_block_sizes(n);
  if (sizes) std::copy(v.begin(), v.end(), sizes);
  return (int)v.size();
}

}  // extern "C"
