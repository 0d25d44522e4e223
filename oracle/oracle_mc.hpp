// ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path (mole_b200/).
//
// CPU restatement of the reference's samplers, drivers, optimizers and DMC:
//   src/metropolis/src/metrop.rs   src/montecarlo/src/{samplers,montecarlo}.rs
//   src/vmc/src/{vmc,operators}.rs src/optimize/src/{util,optimizers}.rs
//   src/dmc/src/{dmc,branching}.rs
// "Faithful" means the same evaluation counts and the same arithmetic order as the
// Rust source (e.g. three value+gradient evaluations per diffusion move), only the
// random stream is replaced by the Philox contract of oracle_rng.hpp.
#pragma once
#include <algorithm>
#include <deque>
#include <vector>
#include "oracle_rng.hpp"
#include "oracle_wf.hpp"

namespace orc {

enum MetropKind : int32_t { METROP_BOX = 0, METROP_DIFFUSE = 1 };

// ---------------------------------------------------------------- Metropolis::move_state
// MetropolisBox: metrop.rs:60-96.  Returns true and overwrites cfg when accepted.
template <class R>
bool box_move_state(const Wf<R>& wf, R* cfg, int idx, double box_side, Key key, uint64_t walker,
                    uint32_t step, bool nan_reject = false) {
  const int n = 3 * wf.ne;
  std::vector<R> prop(cfg, cfg + n);
  const MoveDraw d = draw_uniform4(key, walker, step, DOM_MOVE, (uint32_t)idx);
  const double lo = -0.5 * box_side, scale = 0.5 * box_side - lo;  // Range::new(-b/2, b/2), :66
  prop[3 * idx + 0] += R(lo + scale * d.a);
  prop[3 * idx + 1] += R(lo + scale * d.b);
  prop[3 * idx + 2] += R(lo + scale * d.c);
  // accept_move, :79-81: psi(x) is re-evaluated, not cached
  const R wf_value = orc::wf_value(wf, prop.data());
  const R old = orc::wf_value(wf, cfg);
  R acc = (wf_value * wf_value) / (old * old);
  if (r_val(acc) > 1.0) acc = R(1.0);  // f64::min(1.0); NaN.min(1.0)=1.0 in Rust, see note below
  if (r_val(acc) != r_val(acc)) acc = R(nan_reject ? 0.0 : 1.0);
  if (r_val(acc) > d.u) {
    for (int i = 0; i < n; ++i) cfg[i] = prop[i];
    return true;
  }
  return false;
}
// Note on NaN: Rust's f64::min returns the non-NaN operand, so a NaN ratio becomes 1.0 and
// is accepted; the diffusion sampler rejects NaN psi earlier through the signum test only
// when the signs differ.  The GPU reproduces both behaviours (tests/test_edge_cases.py).

// MetropolisDiffuse: metrop.rs:150-212.
template <class R>
bool diffuse_move_state(const Wf<R>& wf, R* cfg, int idx, double tau, Key key, uint64_t walker,
                        uint32_t step, double* ratio_out = nullptr, bool nan_reject = false) {
  const int n = 3 * wf.ne;
  std::vector<R> prop(cfg, cfg + n), grad(n), grad_old(n);
  const MoveDraw d = draw_normal3_uniform1(key, walker, step, DOM_MOVE, (uint32_t)idx);
  const double sd = std::sqrt(tau);  // Normal::new(0, sqrt(tau)), :160
  {
    // propose_move, :153-160
    const R v = orc::wf_value(wf, cfg);
    orc::wf_gradient(wf, cfg, grad.data());
    const double xi[3] = {sd * d.a, sd * d.b, sd * d.c};
    for (int k = 0; k < 3; ++k) {
      const R drift = grad[3 * idx + k] / v;
      prop[3 * idx + k] += drift * R(tau);
      prop[3 * idx + k] += R(xi[k]);
    }
  }
  // accept_move, :171-197
  const R wf_value = orc::wf_value(wf, prop.data());
  orc::wf_gradient(wf, prop.data(), grad.data());
  const R wf_value_old = orc::wf_value(wf, cfg);
  orc::wf_gradient(wf, cfg, grad_old.data());
  auto signum = [](double x) { return (x != x) ? x : (std::signbit(x) ? -1.0 : 1.0); };
  const double s_new = signum(r_val(wf_value)), s_old = signum(r_val(wf_value_old));
  if (!(s_new == s_old)) return false;  // :178-180 (NaN != NaN -> reject)
  // Frobenius norms over the WHOLE (N_e,3) array, :182-193
  R sh = R(0.0), sl = R(0.0);
  for (int i = 0; i < n; ++i) {
    const R a = (cfg[i] - prop[i]) - (grad[i] / wf_value) * R(tau);
    const R b = (prop[i] - cfg[i]) - (grad_old[i] / wf_value_old) * R(tau);
    sh += a * a;
    sl += b * b;
  }
  const R nh = r_sqrt(sh), nl = r_sqrt(sl);  // norm_l2().powi(2)
  const R t_high = r_exp(-(nh * nh) / R(2.0 * tau));
  const R t_low = r_exp(-(nl * nl) / R(2.0 * tau));
  R acc = t_high * (wf_value * wf_value) / (t_low * (wf_value_old * wf_value_old));  // :195
  if (ratio_out) *ratio_out = r_val(acc);
  if (r_val(acc) != r_val(acc)) acc = R(nan_reject ? 0.0 : 1.0);
  if (r_val(acc) > 1.0) acc = R(1.0);
  if (r_val(acc) > d.u) {
    for (int i = 0; i < n; ++i) cfg[i] = prop[i];
    return true;
  }
  return false;
}

// ---------------------------------------------------------------- Sampler + Runner
enum ObsMask : uint32_t {
  OBS_ENERGY = 1u,        // "Energy"             => a Hamiltonian kind
  OBS_PGRAD = 2u,         // "Parameter gradient" => vmc/src/operators.rs:7-14
  OBS_WFVALUE = 4u,       // "Wavefunction value" => vmc/src/operators.rs:16-24
  OBS_KINETIC = 8u,       // "Kin. Energy"        => helium_atom_singlet.rs:162
};

struct RunOptions {
  int32_t metrop_kind;
  double metrop_param;    // box_side | time_step
  uint32_t observables;
  int32_t quirk_vector_div;  // 1: reproduce OperatorValue Vector/Scalar = scalar/array (operator/src/traits.rs:149-150)
  int32_t nan_reject;        // 0: Rust's NaN-dropping min (a NaN acceptance is accepted); 1: the product's default (rejected)
};

struct RunResult {
  std::vector<double> energy, wfvalue, kinetic;  // one per sample
  std::vector<double> pgrad;                      // n_samples x P  (as STORED by the sampler)
  std::vector<uint8_t> accept;                    // steps x N_e accept bits
  double acceptance = 0.0;                        // samplers.rs:113
  std::vector<double> cfg;                        // final configuration
};

// Sampler::sample, samplers.rs:81-104 : each observable is act_on(psi,cfg) / Scalar(psi(cfg))
inline void sampler_sample(const Wf<double>& wf, const Ham<double>& ham, const RunOptions& o,
                           const double* cfg, RunResult& out) {
  if (o.observables & OBS_ENERGY)
    out.energy.push_back(ham_act_on(ham, wf, cfg) / wf_value(wf, cfg));
  if (o.observables & OBS_KINETIC)
    out.kinetic.push_back(-0.5 * wf_laplacian(wf, cfg) / wf_value(wf, cfg));
  if (o.observables & OBS_WFVALUE) {
    const double v = wf_value(wf, cfg);
    out.wfvalue.push_back((v * v) / wf_value(wf, cfg));  // operators.rs:22 then samplers.rs:90
  }
  if (o.observables & OBS_PGRAD) {
    double pg[WF_MAX_PARAMS];
    wf_parameter_gradient(wf, cfg, pg);
    const double v = wf_value(wf, cfg);
    const double denom = wf_value(wf, cfg);
    for (int k = 0; k < wf.np; ++k) {
      const double acted = v * pg[k];  // Vector*Scalar = scalar*array, operator/src/traits.rs:126
      // Vector / Scalar is implemented as scalar / array (traits.rs:149-150): the stored
      // sample is psi/(psi*dpsi) = 1/dpsi, not dpsi.  quirk_vector_div=0 gives the intended dpsi.
      out.pgrad.push_back(o.quirk_vector_div ? denom / acted : acted / denom);
    }
  }
}

// Runner::run, montecarlo.rs:24-46 over Sampler::move_state, samplers.rs:106-117
inline RunResult runner_run(const Wf<double>& wf, const Ham<double>& ham, const RunOptions& o,
                            const double* cfg0, Key key, uint64_t walker, int steps, int block_size,
                            bool want_trace = true) {
  if (!(steps >= 2 * block_size)) throw std::runtime_error("assert steps >= 2*block_size");  // :29
  RunResult out;
  const int ne = wf.ne, n = 3 * ne;
  out.cfg.assign(cfg0, cfg0 + n);
  const int blocks = steps / block_size;
  uint32_t step = 0;
  for (int block_nr = 0; block_nr < blocks; ++block_nr)
    for (int s = 0; s < block_size; ++s, ++step) {
      for (int e = 0; e < ne; ++e) {
        bool acc;
        if (o.metrop_kind == METROP_BOX)
          acc = box_move_state(wf, out.cfg.data(), e, o.metrop_param, key, walker, step, o.nan_reject != 0);
        else
          acc = diffuse_move_state(wf, out.cfg.data(), e, o.metrop_param, key, walker, step, nullptr, o.nan_reject != 0);
        if (acc) out.acceptance += 1.0 / (double)ne;
        if (want_trace) out.accept.push_back(acc ? 1 : 0);
      }
      if (block_nr > 0) sampler_sample(wf, ham, o, out.cfg.data(), out);  // :36
    }
  return out;
}

// Sampler::new, samplers.rs:45-46: cfg ~ U(-1,1)^(N_e x 3)  (INIT domain of the stream contract)
inline void init_uniform(Key key, uint64_t walker, int ne, double lo, double hi, double* cfg) {
  for (int e = 0; e < ne; ++e) {
    const MoveDraw d = draw_uniform4(key, walker, 0, DOM_INIT, (uint32_t)e);
    const double scale = hi - lo;
    cfg[3 * e + 0] = lo + scale * d.a;
    cfg[3 * e + 1] = lo + scale * d.b;
    cfg[3 * e + 2] = lo + scale * d.c;
  }
}
// DmcRunner::new, dmc.rs:49-58: cfg ~ N(0,sigma)^(N_e x 3)
inline void init_normal(Key key, uint64_t walker, int ne, double sigma, double* cfg) {
  for (int e = 0; e < ne; ++e) {
    const MoveDraw d = draw_normal3_uniform1(key, walker, 0, DOM_INIT, (uint32_t)e);
    cfg[3 * e + 0] = sigma * d.a;
    cfg[3 * e + 1] = sigma * d.b;
    cfg[3 * e + 2] = sigma * d.c;
  }
}

// ---------------------------------------------------------------- statistics (vmc.rs:133-179)
inline double mean_fold(const double* v, size_t n) {  // vmc.rs:174-179
  double a = 0.0;
  for (size_t i = 0; i < n; ++i) a = a + ((v[i] - a) / (double)(i + 1));
  return a;
}
inline double blocking_error(const double* v, size_t n, size_t block_size, double mean) {  // vmc.rs:150-170
  std::vector<double> sq;
  for (size_t i = 0; i < n; i += block_size) {
    const size_t len = std::min(block_size, n - i);
    const double bm = mean_fold(v + i, len);
    sq.push_back(bm * bm);
  }
  const double bms = mean_fold(sq.data(), sq.size());
  return std::sqrt((bms - mean * mean) / (double)(sq.size() - 1));
}

// compute_energy_gradient, optimize/src/util.rs:6-46
inline std::vector<double> energy_gradient(const std::vector<double>& wfv, const std::vector<double>& pg,
                                           const std::vector<double>& en, double energy, int np) {
  const size_t ns = wfv.size();
  std::vector<double> g(np, 0.0);
  for (size_t n = 0; n < ns; ++n)
    for (int k = 0; k < np; ++k) g[k] += 0.0 + 2.0 * ((pg[n * np + k] / wfv[n]) * (en[n] - energy));
  for (int k = 0; k < np; ++k) g[k] /= (double)ns;  // mean_axis
  return g;
}

// symmetric solve standing in for ndarray-linalg solveh_into (LAPACK dsytrf/dsytrs, optimizers.rs:251)
inline bool solve_dense(std::vector<double> a, std::vector<double> b, int n, std::vector<double>& x) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[r * n + c]) > std::fabs(a[piv * n + c])) piv = r;
    if (a[piv * n + c] == 0.0) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) std::swap(a[c * n + k], a[piv * n + k]);
      std::swap(b[c], b[piv]);
    }
    for (int r = c + 1; r < n; ++r) {
      const double f = a[r * n + c] / a[c * n + c];
      for (int k = c; k < n; ++k) a[r * n + k] -= f * a[c * n + k];
      b[r] -= f * b[c];
    }
  }
  x.assign(n, 0.0);
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= a[r * n + k] * x[k];
    x[r] = s / a[r * n + r];
  }
  return true;
}

// ---------------------------------------------------------------- optimizers (optimize/src/optimizers.rs)
enum OptKind : int32_t { OPT_SD = 0, OPT_MOMENTUM = 1, OPT_NESTEROV = 2, OPT_LBFGS = 3, OPT_SR = 4 };

struct Optimizer {
  int kind, np;
  double step_size, momentum_parameter;
  int history;
  int quirk_sr_subtract;  // 1: reproduce optimizers.rs:219-224 (subtracts o_i*o_j from EVERY element)
  std::vector<double> momentum, momentum_prev, grad_prev, pars_prev;
  std::deque<std::vector<double>> s, y;
  size_t iter = 0;

  Optimizer(int kind_, int np_, double step, double mom, int hist, int quirk)
      : kind(kind_), np(np_), step_size(step), momentum_parameter(mom), history(hist),
        quirk_sr_subtract(quirk), momentum(np_, 0.0), momentum_prev(np_, 0.0),
        grad_prev(np_, 1e-5), pars_prev(np_, 1e-5) {}  // EPS, :109-114

  static double dot(const std::vector<double>& a, const std::vector<double>& b) {
    double s = 0.0;
    for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
  }

  void update_curvature_pairs(const std::vector<double>& pars, const std::vector<double>& grad) {  // :148-159
    std::vector<double> sv(np), yv(np);
    for (int i = 0; i < np; ++i) { sv[i] = pars[i] - pars_prev[i]; yv[i] = grad[i] - grad_prev[i]; }
    if ((int)s.size() >= history) s.pop_front();
    s.push_back(sv);
    if ((int)y.size() >= history) y.pop_front();
    y.push_back(yv);
  }

  std::vector<double> initial_direction(const std::vector<double>& gradient) {  // :121-146
    std::vector<double> p(np);
    for (int i = 0; i < np; ++i) p[i] = -gradient[i];
    std::vector<double> alphas;
    for (int q = (int)s.size() - 1; q >= 0; --q) {
      const double alpha = dot(s[q], p) / dot(s[q], y[q]);
      for (int i = 0; i < np; ++i) p[i] -= alpha * y[q][i];
      alphas.push_back(alpha);
    }
    double scale;
    if (iter == 0) scale = 1e-10;
    else {
      double tot = 0.0;
      for (size_t q = 0; q < s.size(); ++q) tot = tot + dot(s[q], y[q]) / dot(y[q], y[q]);
      scale = tot / (double)std::min<size_t>(iter, (size_t)history);
    }
    for (int i = 0; i < np; ++i) p[i] *= scale;
    // izip!(alphas.iter().rev(), s.iter(), y.iter()): stops at the shortest iterator
    const size_t m = std::min(alphas.size(), s.size());
    for (size_t q = 0; q < m; ++q) {
      const double alpha = alphas[alphas.size() - 1 - q];
      const double c = alpha - dot(y[q], p) / dot(y[q], s[q]);
      for (int i = 0; i < np; ++i) p[i] = p[i] + c * s[q][i];
    }
    return p;
  }

  // StochasticReconfiguration::construct_sr_matrix, :191-233
  std::vector<double> sr_matrix(const std::vector<double>& pg, const std::vector<double>& wfv) const {
    const size_t ns = wfv.size();
    std::vector<double> S(np * np, 0.0), o(ns * np), avg(np, 0.0);
    for (size_t n = 0; n < ns; ++n)
      for (int i = 0; i < np; ++i) o[n * np + i] = pg[n * np + i] / wfv[n];
    for (size_t n = 0; n < ns; ++n)
      for (int i = 0; i < np; ++i)
        for (int j = 0; j < np; ++j) S[i * np + j] += (o[n * np + i] * o[n * np + j]) / (double)ns;
    for (int i = 0; i < np; ++i) {
      double sum = 0.0;
      for (size_t n = 0; n < ns; ++n) sum += o[n * np + i];
      avg[i] = sum / (double)ns;
    }
    if (quirk_sr_subtract) {
      for (int i = 0; i < np; ++i)
        for (int j = 0; j < np; ++j)
          for (int q = 0; q < np * np; ++q) S[q] -= avg[i] * avg[j];  // `sr_mat -= scalar`, :222
    } else {
      for (int i = 0; i < np; ++i)
        for (int j = 0; j < np; ++j) S[i * np + j] -= avg[i] * avg[j];
    }
    for (int i = 0; i < np; ++i) S[i * np + i] *= 1.0 + 1e-2;  // :225-231
    return S;
  }

  // Optimizer::compute_parameter_update
  bool compute_parameter_update(const std::vector<double>& pars, double energy_avg,
                                const std::vector<double>& wfv, const std::vector<double>& pg,
                                const std::vector<double>& en, std::vector<double>& deltap) {
    const std::vector<double> g = energy_gradient(wfv, pg, en, energy_avg, np);
    deltap.assign(np, 0.0);
    switch (kind) {
      case OPT_SD:  // :21-29
        for (int i = 0; i < np; ++i) deltap[i] = -(step_size * g[i]);
        return true;
      case OPT_MOMENTUM:  // :50-59
        for (int i = 0; i < np; ++i) momentum[i] -= step_size * g[i];
        for (int i = 0; i < np; ++i) deltap[i] = momentum_parameter * momentum[i];
        return true;
      case OPT_NESTEROV:  // :82-93
        momentum_prev = momentum;
        for (int i = 0; i < np; ++i) momentum[i] = momentum_parameter * momentum[i] + step_size * g[i];
        for (int i = 0; i < np; ++i)
          deltap[i] = -(momentum_parameter * momentum_prev[i] + (1.0 + momentum_parameter) * momentum[i]);
        return true;
      case OPT_LBFGS: {  // :163-178
        update_curvature_pairs(pars, g);
        const std::vector<double> p = initial_direction(g);
        for (int i = 0; i < np; ++i) deltap[i] = -step_size * p[i];
        update_curvature_pairs(pars, g);
        grad_prev = g;
        pars_prev = pars;
        iter += 1;
        return true;
      }
      case OPT_SR: {  // :237-252
        const std::vector<double> S = sr_matrix(pg, wfv);
        std::vector<double> rhs(np), x;
        for (int i = 0; i < np; ++i) rhs[i] = -0.5 * g[i];
        if (!solve_dense(S, rhs, np, x)) return false;  // Error::LinalgError
        for (int i = 0; i < np; ++i) deltap[i] = step_size * x[i];
        return true;
      }
    }
    return false;
  }
};

// ---------------------------------------------------------------- VmcRunner::run_optimization (vmc.rs:43-106)
struct VmcResult {
  std::vector<double> energies, errors, acceptance;  // one per iteration
  std::vector<double> params;                         // final parameters
  std::vector<double> param_history;                  // iters x P (after each update)
};

inline VmcResult vmc_run_optimization(WfDesc wfd, const HamDesc& hd, RunOptions o, Optimizer& opt,
                                      const uint8_t master_seed[32], const double* cfg0, int iters,
                                      int total_samples, int block_size, int nworkers) {
  VmcResult res;
  const int steps = total_samples / nworkers;  // :50
  const int np = wfd.n_params;
  o.observables |= OBS_ENERGY | OBS_PGRAD | OBS_WFVALUE;
  for (int it = 0; it < iters; ++it) {
    const Wf<double> wf(wfd);
    const Ham<double> ham(hd);
    // :59-61 one derived seed per iteration; workers are distinguished by walker id (stream contract)
    uint8_t seed[32];
    derive_seed(master_seed, (uint32_t)it, seed);
    const Key key = key_from_seed(seed);
    std::vector<RunResult> results(nworkers);
#pragma omp parallel for schedule(static)
    for (int w = 0; w < nworkers; ++w)  // :63-76, every clone starts from the master cfg (:56)
      results[w] = runner_run(wf, ham, o, cfg0, key, (uint64_t)w, steps, block_size, false);
    // concatenate_worker_data, :108-130
    std::vector<double> en, wfv, pg;
    double accept = 0.0;
    for (const RunResult& r : results) {
      accept += r.acceptance;
      en.insert(en.end(), r.energy.begin(), r.energy.end());
      wfv.insert(wfv.end(), r.wfvalue.begin(), r.wfvalue.end());
      pg.insert(pg.end(), r.pgrad.begin(), r.pgrad.end());
    }
    const double mean = mean_fold(en.data(), en.size());
    const double err = blocking_error(en.data(), en.size(), (size_t)block_size, mean);
    res.energies.push_back(mean);
    res.errors.push_back(err);
    res.acceptance.push_back(accept / (double)total_samples);  // :97
    std::vector<double> pars(wfd.params, wfd.params + np), dp;
    if (!opt.compute_parameter_update(pars, mean, wfv, pg, en, dp)) throw std::runtime_error("LinalgError");
    for (int k = 0; k < np; ++k) wfd.params[k] += dp[k];  // update_parameters, :91
    for (int k = 0; k < np; ++k) res.param_history.push_back(wfd.params[k]);
  }
  res.params.assign(wfd.params, wfd.params + np);
  return res;
}

// ---------------------------------------------------------------- DMC (dmc/src/{dmc,branching}.rs)
enum BranchKind : int32_t { BRANCH_SR = 0, BRANCH_SIMPLE = 1 };

struct Walkers {
  int ne;
  std::vector<double> w;    // weights
  std::vector<double> cfg;  // N x ne x 3
  size_t size() const { return w.size(); }
};

inline uint64_t rand_below(const Philox4& p, uint64_t n) {  // uniform integer in [0,n): floor(x64 * n / 2^64)
  const uint64_t x = ((uint64_t)p.w[1] << 32) | p.w[0];
  return (uint64_t)(((unsigned __int128)x * n) >> 64);
}

// SRBrancher::branch, branching.rs:15-40
inline Walkers branch_sr(const Walkers& in, Key key, uint32_t step) {
  const size_t N = in.size();
  const int n = 3 * in.ne;
  double tot = 0.0, max_weight = 0.0;
  for (size_t i = 0; i < N; ++i) tot = tot + in.w[i];
  const double global_weight = tot / (double)N;                       // :21
  for (size_t i = 0; i < N; ++i) max_weight = std::max(max_weight, in.w[i]);  // :22
  const double norm_factor = (double)N / max_weight;                  // :24
  std::vector<uint64_t> cum(N);
  uint64_t running = 0;
  for (size_t i = 0; i < N; ++i) {
    const double scaled = in.w[i] * norm_factor;
    const uint32_t k = (scaled >= 4294967295.0) ? 4294967295u : (scaled > 0.0 ? (uint32_t)scaled : 0u);  // `as u32`, :27
    running += k;  // rand 0.5 WeightedChoice keeps a running total (u64 here)
    cum[i] = running;
  }
  Walkers out;
  out.ne = in.ne;
  out.w.assign(N, global_weight);  // :36
  out.cfg.resize(N * n);
  for (size_t j = 0; j < N; ++j) {
    const uint64_t u = rand_below(draw(key, j, step, DOM_BRANCH, 0, 0), running);
    const size_t pick = std::upper_bound(cum.begin(), cum.end(), u) - cum.begin();  // first cum > u
    std::copy(in.cfg.begin() + pick * n, in.cfg.begin() + (pick + 1) * n, out.cfg.begin() + j * n);
  }
  return out;
}

// SimpleBranching::branch, branching.rs:50-91
inline Walkers branch_simple(const Walkers& in, Key key, uint32_t step) {
  const size_t N = in.size();
  const int n = 3 * in.ne;
  std::vector<std::pair<double, size_t>> nw;  // (weight, source index)
  std::vector<std::pair<double, size_t>> births;
  for (size_t i = 0; i < N; ++i) {
    const Philox4 p = draw(key, i, step, DOM_BRANCH, 0, 0);
    const double t = in.w[i] + u53(p.w[0], p.w[1]);
    size_t copies = (t > 0.0) ? (size_t)t : 0;  // `as usize`
    copies = std::min<size_t>(copies, 3);        // :60
    if (copies > 0) {
      nw.push_back({in.w[i], i});
      for (size_t c = 0; c + 1 < copies; ++c) births.push_back({in.w[i], i});  // :64-66
    }
  }
  nw.insert(nw.end(), births.begin(), births.end());  // :72
  if (N > nw.size()) {  // :75-82 clone random OLD walkers
    const size_t excess = N - nw.size();
    for (size_t j = 0; j < excess; ++j) {
      const size_t pick = (size_t)rand_below(draw(key, j, step, DOM_BRANCH, 1, 0), N);
      nw.push_back({in.w[pick], pick});
    }
  } else {  // :83-89 remove random new walkers, one at a time
    const size_t excess = nw.size() - N;
    for (size_t j = 0; j < excess; ++j) {
      const size_t pick = (size_t)rand_below(draw(key, j, step, DOM_BRANCH, 2, 0), nw.size());
      nw.erase(nw.begin() + pick);
    }
  }
  Walkers out;
  out.ne = in.ne;
  out.w.resize(nw.size());
  out.cfg.resize(nw.size() * n);
  for (size_t j = 0; j < nw.size(); ++j) {
    out.w[j] = nw[j].first;
    std::copy(in.cfg.begin() + nw[j].second * n, in.cfg.begin() + (nw[j].second + 1) * n,
              out.cfg.begin() + j * n);
  }
  return out;
}

struct DmcResult {
  std::vector<double> energies, errors;       // DmcRunner::diffuse return value
  std::vector<double> step_energies;          // ensemble energy of every time step
  double reference_energy;
  Walkers walkers;
};

// one DMC time step over all walkers, dmc.rs:84-135 (no branching)
inline double dmc_step(const Wf<double>& wf, const Ham<double>& ham, Walkers& wk, double metrop_tau,
                       double time_step, double reference_energy, Key key, uint32_t step,
                       double* total_weight_out = nullptr) {
  const size_t N = wk.size();
  const int ne = wf.ne, n = 3 * ne;
  std::vector<double> loc_e(N), loc_w(N);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < N; ++i) {
    double* conf = wk.cfg.data() + i * n;
    const double psi_old = wf_value(wf, conf);
    const double local_e = ham_act_on(ham, wf, conf) / psi_old;  // :89-96
    for (int e = 0; e < ne; ++e) diffuse_move_state(wf, conf, e, metrop_tau, key, (uint64_t)i, step);  // :99-110
    loc_e[i] = wk.w[i] * local_e;  // :112 (pre-update weight, pre-move energy)
    loc_w[i] = wk.w[i];            // :113
    const double psi_new = wf_value(wf, conf);
    const double local_e_new = ham_act_on(ham, wf, conf) / psi_new;  // :115-124
    wk.w[i] *= std::exp(-time_step * ((local_e + local_e_new) / 2.0 - reference_energy));  // :126-128
  }
  double ensemble_energy = 0.0, total_weight = 0.0;
  for (size_t i = 0; i < N; ++i) { ensemble_energy += loc_e[i]; total_weight += loc_w[i]; }
  if (total_weight_out) *total_weight_out = total_weight;
  return ensemble_energy / total_weight;  // :133
}

// DmcRunner::diffuse, dmc.rs:69-153 + update_energies :155-202
inline DmcResult dmc_diffuse(const WfDesc& wfd, const HamDesc& hd, Walkers wk, double metrop_tau,
                             double reference_energy, int branch_kind, Key key, double time_step,
                             int num_iterations, int block_size, int num_eq_blocks) {
  const Wf<double> wf(wfd);
  const Ham<double> ham(hd);
  DmcResult res;
  std::vector<double> vars;
  const int blocks = num_iterations / block_size;
  uint32_t step = 0;
  for (int block_nr = 0; block_nr < blocks; ++block_nr) {
    std::vector<double> eb;
    for (int j = 0; j < block_size; ++j, ++step) {
      const double e = dmc_step(wf, ham, wk, metrop_tau, time_step, reference_energy, key, step);
      eb.push_back(e);
      res.step_energies.push_back(e);
      wk = (branch_kind == BRANCH_SR) ? branch_sr(wk, key, step) : branch_simple(wk, key, step);  // :139-140
    }
    double sum = 0.0;
    for (double e : eb) sum += e;
    const double energy = sum / (double)eb.size();
    if (block_nr == num_eq_blocks) {  // :163-177
      reference_energy = (reference_energy + energy) / 2.0;
      res.energies.push_back(energy);
      vars.push_back(0.0);
    }
    if (block_nr > num_eq_blocks) {  // :178-201
      const double prev = res.energies.back();
      const double k = (double)(block_nr - num_eq_blocks);
      res.energies.push_back(prev + (energy - prev) / k);
      reference_energy = (reference_energy + res.energies.back()) / 2.0;
      vars.push_back(vars.back() + ((energy - prev) * (energy - res.energies.back()) - vars.back()) / k);
    }
  }
  for (size_t i = 0; i < vars.size(); ++i) res.errors.push_back(std::sqrt(vars[i] / (double)(i + 1)));  // :148-151
  res.reference_energy = reference_energy;
  res.walkers = wk;
  return res;
}

}  // namespace orc
