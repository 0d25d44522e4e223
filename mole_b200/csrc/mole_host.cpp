// Host side of the C ABI: seed derivation, the accumulator finaliser, the optimizers and the
// drivers (Runner / VmcRunner / DmcRunner).  Everything here is O(P^3) or O(iterations) work on
// reduced moments; all per-walker arithmetic happens in the CUDA kernels (mole_api.cu).
// Product code: never includes anything from oracle/.
#include <cmath>
#include <cstring>
#include <deque>
#include <vector>
#include <algorithm>
#include "mole_internal.h"
#define MOLE_HOST_ONLY   // only the host-callable Philox helpers of mole_rng.cuh are needed here
#include "mole_rng.cuh"

RngKey mole_key_from_seed(const uint8_t seed[32]) {
  uint32_t s[8];
  for (int i = 0; i < 8; ++i)
    s[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
           ((uint32_t)seed[4 * i + 3] << 24);
  return RngKey{s[0] ^ s[2] ^ s[4] ^ s[6], s[1] ^ s[3] ^ s[5] ^ s[7]};
}

int mole_oo_index(int P, int k, int l) { return k * P - k * (k - 1) / 2 + (l - k); }

uint64_t mole_el_signature(const WfParams& w, const HamParams& h) {
  uint64_t x = 1469598103934665603ull;
  auto eat = [&x](const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { x ^= b[i]; x *= 1099511628211ull; }
  };
  eat(&w.kind, sizeof(w.kind)); eat(&w.ne, sizeof(w.ne)); eat(w.p, sizeof(w.p)); eat(w.geom, sizeof(w.geom));
  eat(&h.kind, sizeof(h.kind)); eat(&h.n_ions, sizeof(h.n_ions)); eat(h.ion_pos, sizeof(h.ion_pos)); eat(h.ion_z, sizeof(h.ion_z));
  eat(&h.frequency, sizeof(h.frequency));
  return x ? x : 1;
}

// ------------------------------------------------------------------ optimizers
struct mole_opt_s {
  int kind, np;
  double step_size, momentum_parameter;
  int history;
  uint32_t compat;
  double sr_diag_scale = 1.0 + 1e-2;   // optimizers.rs:225-231
  double sr_diag_shift = 0.0;          // mole_opt_set_sr_regularization
  std::vector<double> momentum, momentum_prev, grad_prev, pars_prev;
  std::deque<std::vector<double>> s, y;
  size_t iter = 0;
};

static double vdot(const std::vector<double>& a, const std::vector<double>& b) {
  double r = 0.0;
  for (size_t i = 0; i < a.size(); ++i) r += a[i] * b[i];
  return r;
}

// the reduced moments every optimizer needs, from either source: the packed accumulators (P <= MOLE_ACC_MAX_PARAMS)
// or the Gram matrix of the per-sample rows (1, E_L, O_1 .. O_P) of a large-P kind
struct Moments {
  int np = 0;
  double n = 0.0, sum_e = 0.0;
  std::vector<double> o, oe, oo;   // oo: full np x np, symmetric
  bool finite() const {
    bool ok = std::isfinite(n) && std::isfinite(sum_e);
    for (double v : o) ok = ok && std::isfinite(v);
    for (double v : oe) ok = ok && std::isfinite(v);
    for (double v : oo) ok = ok && std::isfinite(v);
    return ok;
  }
};

static Moments moments_from_acc(const mole_acc_host* a, int np) {
  Moments m;
  m.np = np; m.n = a->n_samples; m.sum_e = a->sum_e;
  m.o.assign(a->sum_o, a->sum_o + np);
  m.oe.assign(a->sum_oe, a->sum_oe + np);
  m.oo.assign((size_t)np * np, 0.0);
  for (int k = 0; k < np; ++k)
    for (int l = 0; l < np; ++l) m.oo[(size_t)k * np + l] = a->sum_oo[mole_oo_index(np, std::min(k, l), std::max(k, l))];
  return m;
}

static Moments moments_from_gram(int n_cols, const double* g) {
  Moments m;
  const int np = n_cols - 2;
  m.np = np; m.n = g[0]; m.sum_e = g[1];
  m.o.resize(np); m.oe.resize(np); m.oo.resize((size_t)np * np);
  for (int k = 0; k < np; ++k) {
    m.o[k] = g[2 + k];
    m.oe[k] = g[(size_t)n_cols + 2 + k];
    for (int l = 0; l < np; ++l) m.oo[(size_t)k * np + l] = g[(size_t)(2 + k) * n_cols + 2 + l];
  }
  return m;
}

// gradient of the energy from the reduced moments: g_k = 2 (<O_k E> - <O_k><E>)   (util.rs:37-45)
static bool energy_gradient(const Moments& a, std::vector<double>& g) {
  if (!(a.n > 0.0)) return false;
  const double n = a.n, ebar = a.sum_e / n;
  g.assign(a.np, 0.0);
  for (int k = 0; k < a.np; ++k) g[k] = 2.0 * (a.oe[k] / n - (a.o[k] / n) * ebar);
  return true;
}

// S_kl = <O_k O_l> - <O_k><O_l>, diagonal scaled by 1.01   (optimizers.rs:191-233)
static void sr_matrix(const mole_opt_s* o, const Moments& a, std::vector<double>& S) {
  const int np = o->np;
  const double n = a.n;
  S.assign((size_t)np * np, 0.0);
  double avg_sum = 0.0;
  for (int k = 0; k < np; ++k) avg_sum += a.o[k] / n;
  for (int k = 0; k < np; ++k)
    for (int l = 0; l < np; ++l) {
      double v = a.oo[(size_t)k * np + l] / n;
      if (o->compat & MOLE_COMPAT_SR_SUBTRACT) v -= avg_sum * avg_sum;   // sum_ij o_i o_j removed from every element (:219-224)
      else v -= (a.o[k] / n) * (a.o[l] / n);
      S[(size_t)k * np + l] = v;
    }
  for (int k = 0; k < np; ++k) S[(size_t)k * np + k] = S[(size_t)k * np + k] * o->sr_diag_scale + o->sr_diag_shift;
}

// dense solve with partial pivoting; stands in for LAPACK dsytrf/dsytrs (optimizers.rs:251).  Singular means a
// pivot below n eps max|S| (or NaN): a rank-deficient S amplifies the noise of g without bound, so the step is
// refused (Error::LinalgError) instead of applied.
static bool solve_dense(std::vector<double> a, std::vector<double> b, int n, std::vector<double>& x) {
  double amax = 0.0;
  for (double v : a) amax = std::max(amax, std::fabs(v));
  if (!(amax > 0.0) || !std::isfinite(amax)) return false;
  const double tiny = (double)n * 2.220446049250313e-16 * amax;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[(size_t)r * n + c]) > std::fabs(a[(size_t)piv * n + c])) piv = r;
    const double d = a[(size_t)piv * n + c];
    if (!(std::fabs(d) > tiny)) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) std::swap(a[(size_t)c * n + k], a[(size_t)piv * n + k]);
      std::swap(b[c], b[piv]);
    }
    for (int r = c + 1; r < n; ++r) {
      const double f = a[(size_t)r * n + c] / a[(size_t)c * n + c];
      for (int k = c; k < n; ++k) a[(size_t)r * n + k] -= f * a[(size_t)c * n + k];
      b[r] -= f * b[c];
    }
  }
  x.assign(n, 0.0);
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= a[(size_t)r * n + k] * x[k];
    x[r] = s / a[(size_t)r * n + r];
  }
  return true;
}

static void lbfgs_push(mole_opt_s* o, const std::vector<double>& pars, const std::vector<double>& grad) {  // :148-159
  std::vector<double> sv(o->np), yv(o->np);
  for (int i = 0; i < o->np; ++i) { sv[i] = pars[i] - o->pars_prev[i]; yv[i] = grad[i] - o->grad_prev[i]; }
  if ((int)o->s.size() >= o->history) o->s.pop_front();
  o->s.push_back(sv);
  if ((int)o->y.size() >= o->history) o->y.pop_front();
  o->y.push_back(yv);
}

static std::vector<double> lbfgs_direction(mole_opt_s* o, const std::vector<double>& g) {  // :121-146
  const int np = o->np;
  std::vector<double> p(np), alphas;
  for (int i = 0; i < np; ++i) p[i] = -g[i];
  for (int q = (int)o->s.size() - 1; q >= 0; --q) {
    const double alpha = vdot(o->s[q], p) / vdot(o->s[q], o->y[q]);
    for (int i = 0; i < np; ++i) p[i] -= alpha * o->y[q][i];
    alphas.push_back(alpha);
  }
  double scale = 1e-10;
  if (o->iter != 0) {
    double tot = 0.0;
    for (size_t q = 0; q < o->s.size(); ++q) tot += vdot(o->s[q], o->y[q]) / vdot(o->y[q], o->y[q]);
    scale = tot / (double)std::min<size_t>(o->iter, (size_t)o->history);
  }
  for (int i = 0; i < np; ++i) p[i] *= scale;
  const size_t m = std::min(alphas.size(), o->s.size());
  for (size_t q = 0; q < m; ++q) {
    const double c = alphas[alphas.size() - 1 - q] - vdot(o->y[q], p) / vdot(o->y[q], o->s[q]);
    for (int i = 0; i < np; ++i) p[i] += c * o->s[q][i];
  }
  return p;
}

// ------------------------------------------------------------------ population rebalancing (host plan)
// shares[r] = slots of the global population that rank r's walkers fill after rebalancing: proportional to the rank's
// total weight, rounded systematically with ONE shared draw u in [0, 1) so that they add up to the global count
void mole_rebalance_shares(int nranks, const double* totals, const int64_t* counts, double u, int64_t* shares) {
  double T = 0.0;
  int64_t N = 0;
  for (int r = 0; r < nranks; ++r) { T += totals[r]; N += counts[r]; }
  double cum = 0.0;
  int64_t prev = 0;
  for (int r = 0; r < nranks; ++r) {
    cum += totals[r];
    int64_t edge = (r == nranks - 1) ? N : (int64_t)std::floor(cum * (double)N / T + u);
    edge = std::min<int64_t>(std::max<int64_t>(edge, prev), N);
    shares[r] = edge - prev;
    prev = edge;
  }
}

// surplus copies (share > count) travel to free slots (share < count), matched greedily in rank order.  A sender
// keeps copies 0 .. count-1 and sends copies count .. share-1 (rows 0.. of its send buffer); a receiver fills its
// slots share .. count-1 (rows 0.. of its receive buffer)
std::vector<MoleMove> mole_rebalance_moves(int nranks, const int64_t* counts, const int64_t* shares) {
  std::vector<MoleMove> mv;
  std::vector<int64_t> sent(nranks, 0), got(nranks, 0);
  int d = 0;
  for (int s = 0; s < nranks; ++s) {
    int64_t left = shares[s] - counts[s];
    while (left > 0) {
      while (d < nranks && (counts[d] - shares[d]) - got[d] <= 0) ++d;
      if (d >= nranks) return mv;   // cannot happen: the shares add up to the counts
      const int64_t room = (counts[d] - shares[d]) - got[d];
      const int64_t n = std::min(left, room);
      mv.push_back(MoleMove{s, d, sent[s], got[d], n});
      sent[s] += n; got[d] += n; left -= n;
    }
  }
  return mv;
}

extern "C" {

// the plan alone, for tests and for callers that bring their own transport
int32_t mole_rebalance_plan(int32_t nranks, const double* totals, const int64_t* counts, double u, int64_t* shares,
                            int64_t* moves /* nranks * nranks, [src][dst] counts; nullable */) {
  if (nranks < 1 || !totals || !counts || !shares || !(u >= 0.0 && u < 1.0)) return MOLE_ERR_INVALID_ARG;
  double T = 0.0;
  for (int r = 0; r < nranks; ++r) {
    if (!(totals[r] >= 0.0) || !std::isfinite(totals[r]) || counts[r] < 0) return MOLE_ERR_INVALID_ARG;
    T += totals[r];
  }
  if (!(T > 0.0)) return mole_set_error(nullptr, MOLE_ERR_DATA_ACCESS, "mole_rebalance: the ensemble has no weight left");
  mole_rebalance_shares(nranks, totals, counts, u, shares);
  if (moves) {
    for (int i = 0; i < nranks * nranks; ++i) moves[i] = 0;
    for (const MoleMove& m : mole_rebalance_moves(nranks, counts, shares)) moves[m.src * nranks + m.dst] += m.count;
  }
  return MOLE_OK;
}

int32_t mole_derive_seed(const uint8_t master[32], uint32_t n, uint8_t out[32]) {
  if (!master || !out) return MOLE_ERR_INVALID_ARG;
  const RngKey k = mole_key_from_seed(master);
  for (uint32_t half = 0; half < 2; ++half) {
    const Philox4 r = mole_draw(k, n, 0, DOM_SEED, 0, half);
    const uint32_t w[4] = {r.a, r.b, r.c, r.d};
    for (int i = 0; i < 4; ++i)
      for (int b = 0; b < 4; ++b) out[16 * half + 4 * i + b] = (uint8_t)(w[i] >> (8 * b));
  }
  return MOLE_OK;
}

int32_t mole_acc_finalize(const mole_acc_host* a, double* energy, double* error, double* acceptance, double* grad) {
  if (!a) return MOLE_ERR_INVALID_ARG;
  if (!(a->n_samples > 0.0)) return mole_set_error(nullptr, MOLE_ERR_DATA_ACCESS, "no \"Energy\" samples accumulated");
  const double mean = a->sum_e / a->n_samples;
  if (energy) *energy = mean;
  if (error) {   // vmc.rs:150-170
    const double bms = a->sum_b2 / a->n_blocks;
    *error = std::sqrt((bms - mean * mean) / (a->n_blocks - 1.0));
  }
  if (acceptance) *acceptance = a->n_moves > 0.0 ? a->n_accept / a->n_moves : 0.0;   // vmc.rs:97
  if (grad) {
    std::vector<double> g;
    energy_gradient(moments_from_acc(a, a->n_params), g);
    for (int k = 0; k < a->n_params; ++k) grad[k] = g[k];
  }
  return MOLE_OK;
}

// the same from the Gram matrix of a large-P kind: mean energy and gradient (the blocking error of such a run comes
// from the ten scalar accumulators, mole_acc_finalize)
int32_t mole_gram_finalize(int32_t n_cols, const double* gram, double* energy, double* grad) {
  if (!gram || n_cols < 3) return MOLE_ERR_INVALID_ARG;
  const Moments m = moments_from_gram(n_cols, gram);
  if (!(m.n > 0.0)) return mole_set_error(nullptr, MOLE_ERR_DATA_ACCESS, "no samples in the Gram matrix");
  if (energy) *energy = m.sum_e / m.n;
  if (grad) {
    std::vector<double> g;
    energy_gradient(m, g);
    for (int k = 0; k < m.np; ++k) grad[k] = g[k];
  }
  return MOLE_OK;
}

int32_t mole_opt_create(int32_t kind, int32_t np, double step, double mom, int32_t history, uint32_t compat, mole_opt_t* out) {
  if (!out || kind < MOLE_OPT_SD || kind > MOLE_OPT_SR || np < 1 || np > MOLE_WF_MAX_PARAMS)
    return mole_set_error(nullptr, MOLE_ERR_INVALID_ARG, "mole_opt_create: bad arguments");
  mole_opt_s* o = new mole_opt_s();
  o->kind = kind; o->np = np; o->step_size = step; o->momentum_parameter = mom; o->history = history; o->compat = compat;
  o->momentum.assign(np, 0.0); o->momentum_prev.assign(np, 0.0);
  o->grad_prev.assign(np, 1e-5); o->pars_prev.assign(np, 1e-5);   // optimizers.rs:109-114
  *out = o;
  return MOLE_OK;
}
int32_t mole_opt_destroy(mole_opt_t o) { delete o; return MOLE_OK; }

int32_t mole_opt_sr_matrix(mole_opt_t o, const mole_acc_host* a, double* S) {
  if (!o || !a || !S) return MOLE_ERR_INVALID_ARG;
  if (a->n_params != o->np) return mole_set_error(nullptr, MOLE_ERR_SHAPE, "accumulator / optimizer parameter count mismatch");
  std::vector<double> m;
  sr_matrix(o, moments_from_acc(a, o->np), m);
  std::copy(m.begin(), m.end(), S);
  return MOLE_OK;
}

static int32_t opt_step_unchecked(mole_opt_t o, const double* pars, const Moments& a, double* deltap) {
  const int np = o->np;
  std::vector<double> g;
  if (!energy_gradient(a, g)) return mole_set_error(nullptr, MOLE_ERR_DATA_ACCESS, "no samples accumulated");
  switch (o->kind) {
    case MOLE_OPT_SD:   // optimizers.rs:21-29
      for (int i = 0; i < np; ++i) deltap[i] = -(o->step_size * g[i]);
      return MOLE_OK;
    case MOLE_OPT_MOMENTUM:   // :50-59
      for (int i = 0; i < np; ++i) o->momentum[i] -= o->step_size * g[i];
      for (int i = 0; i < np; ++i) deltap[i] = o->momentum_parameter * o->momentum[i];
      return MOLE_OK;
    case MOLE_OPT_NESTEROV:   // :82-93
      o->momentum_prev = o->momentum;
      for (int i = 0; i < np; ++i) o->momentum[i] = o->momentum_parameter * o->momentum[i] + o->step_size * g[i];
      for (int i = 0; i < np; ++i)
        deltap[i] = -(o->momentum_parameter * o->momentum_prev[i] + (1.0 + o->momentum_parameter) * o->momentum[i]);
      return MOLE_OK;
    case MOLE_OPT_LBFGS: {   // :163-178 (update_curvature_pairs is called twice per step upstream)
      const std::vector<double> pv(pars, pars + np);
      lbfgs_push(o, pv, g);
      const std::vector<double> p = lbfgs_direction(o, g);
      for (int i = 0; i < np; ++i) deltap[i] = -o->step_size * p[i];
      lbfgs_push(o, pv, g);
      o->grad_prev = g;
      o->pars_prev = pv;
      o->iter += 1;
      return MOLE_OK;
    }
    case MOLE_OPT_SR: {   // :237-252
      std::vector<double> S, rhs(np), x;
      sr_matrix(o, a, S);
      for (int i = 0; i < np; ++i) rhs[i] = -0.5 * g[i];
      if (!solve_dense(S, rhs, np, x)) return mole_set_error(nullptr, MOLE_ERR_LINALG, "singular SR matrix");
      for (int i = 0; i < np; ++i) deltap[i] = o->step_size * x[i];
      return MOLE_OK;
    }
  }
  return MOLE_ERR_INVALID_ARG;
}

static int32_t opt_step_checked(mole_opt_t o, const double* pars, const Moments& m, double* deltap) {
  MOLE_RANGE("mole_opt_step");
  if (!m.finite())
    return mole_set_error(nullptr, MOLE_ERR_DATA_ACCESS, "non-finite optimisation moments (see mole_ensemble_health); no update computed");
  std::vector<double> dp(o->np, 0.0);
  const int32_t rc = opt_step_unchecked(o, pars, m, dp.data());
  if (rc != MOLE_OK) return rc;
  for (int i = 0; i < o->np; ++i)
    if (!std::isfinite(dp[i]))
      return mole_set_error(nullptr, MOLE_ERR_LINALG, "parameter update is not finite (ill-conditioned solve); parameters left untouched");
  for (int i = 0; i < o->np; ++i) deltap[i] = dp[i];
  return MOLE_OK;
}

int32_t mole_opt_step(mole_opt_t o, const double* pars, const mole_acc_host* a, double* deltap) {
  if (!o || !pars || !a || !deltap) return MOLE_ERR_INVALID_ARG;
  if (a->n_params != o->np || o->np > MOLE_ACC_MAX_PARAMS)
    return mole_set_error(nullptr, MOLE_ERR_DATA_ACCESS, "\"Parameter gradient\" moments missing or of the wrong size");
  return opt_step_checked(o, pars, moments_from_acc(a, o->np), deltap);
}

// Optimizer::compute_parameter_update from the Gram matrix of the per-sample rows (1, E_L, O_1 .. O_P) of a large-P
// kind (n_cols = P + 2, row-major, mole_gram_get): the same optimizers on the same moments
int32_t mole_opt_step_gram(mole_opt_t o, const double* pars, int32_t n_cols, const double* gram, double* deltap) {
  if (!o || !pars || !gram || !deltap) return MOLE_ERR_INVALID_ARG;
  if (n_cols != o->np + 2)
    return mole_set_error(nullptr, MOLE_ERR_DATA_ACCESS, "\"Parameter gradient\" moments missing or of the wrong size");
  return opt_step_checked(o, pars, moments_from_gram(n_cols, gram), deltap);
}

int32_t mole_opt_sr_matrix_gram(mole_opt_t o, int32_t n_cols, const double* gram, double* S) {
  if (!o || !gram || !S) return MOLE_ERR_INVALID_ARG;
  if (n_cols != o->np + 2) return mole_set_error(nullptr, MOLE_ERR_SHAPE, "Gram matrix / optimizer parameter count mismatch");
  std::vector<double> m;
  sr_matrix(o, moments_from_gram(n_cols, gram), m);
  std::copy(m.begin(), m.end(), S);
  return MOLE_OK;
}

// S_kk <- S_kk * diag_scale + diag_shift before the solve.  The reference hard-codes (1.01, 0) (optimizers.rs:225-231),
// which is the default; an absolute shift is the usual SR stabiliser when two parameters are nearly redundant.
int32_t mole_opt_set_sr_regularization(mole_opt_t o, double diag_scale, double diag_shift) {
  if (!o || !(diag_scale > 0.0) || !(diag_shift >= 0.0)) return MOLE_ERR_INVALID_ARG;
  o->sr_diag_scale = diag_scale;
  o->sr_diag_shift = diag_shift;
  return MOLE_OK;
}

// ------------------------------------------------------------------ Runner::run (montecarlo.rs:24-46)
int32_t mole_runner_run(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, uint32_t observables, uint32_t compat,
                        int32_t steps, int32_t block_size, double* energy_trace, double* wfvalue_trace, double* kinetic_trace,
                        double* pgrad_trace, uint8_t* accept_trace) {
  if (!ens) return MOLE_ERR_INVALID_ARG;
  if (block_size < 1 || !(steps >= 2 * block_size))
    return mole_set_error(ens->ctx, MOLE_ERR_ASSERT, "assertion failed: steps >= 2 * block_size (montecarlo.rs:29)");
  const int blocks = steps / block_size;
  mole_sweep_args a;
  memset(&a, 0, sizeof(a));
  a.n_sweeps = blocks * block_size;
  a.n_discard = block_size;   // block 0 is equilibration (:36)
  a.block_size = block_size;
  a.observables = observables;
  a.compat = compat;
  a.energy_trace = energy_trace; a.wfvalue_trace = wfvalue_trace; a.kinetic_trace = kinetic_trace;
  a.pgrad_trace = pgrad_trace; a.accept_trace = accept_trace;
  return mole_sweep(ens, wf, m, op, &a);
}

// Runner::run with a logger (montecarlo.rs:31-43, Log: montecarlo/src/traits.rs:44-47): one launch per
// block, the callback sees the block-level reductions (differences of the device accumulators).
int32_t mole_runner_run_logged(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, uint32_t observables,
                               uint32_t compat, int32_t steps, int32_t block_size, uint32_t sweep_flags, mole_log_fn log,
                               void* user) {
  if (!ens) return MOLE_ERR_INVALID_ARG;
  if (block_size < 1 || !(steps >= 2 * block_size))
    return mole_set_error(ens->ctx, MOLE_ERR_ASSERT, "assertion failed: steps >= 2 * block_size (montecarlo.rs:29)");
  const int blocks = steps / block_size;
  mole_acc_host prev, cur, first;
  int32_t rc;
  if ((rc = mole_acc_get(ens, &prev)) != MOLE_OK) return rc;
  first = prev;
  for (int b = 0; b < blocks; ++b) {
    mole_sweep_args a;
    memset(&a, 0, sizeof(a));
    a.n_sweeps = block_size;
    a.n_discard = b == 0 ? block_size : 0;   // block 0 is equilibration (:36)
    a.block_size = block_size;
    a.observables = observables;
    a.compat = compat;
    a.flags = sweep_flags;
    if (b > 1 && (sweep_flags & MOLE_SWEEP_KEEP_SERIES)) a.flags = (sweep_flags & ~MOLE_SWEEP_KEEP_SERIES) | MOLE_SWEEP_APPEND_SERIES;
    if ((rc = mole_sweep(ens, wf, m, op, &a)) != MOLE_OK) return rc;
    if (!log) continue;
    if ((rc = mole_acc_get(ens, &cur)) != MOLE_OK) return rc;
    if (b == 0) { prev = cur; continue; }
    mole_block_log d;
    memset(&d, 0, sizeof(d));
    d.block_nr = b;
    d.block_size = block_size;
    d.n_samples = cur.n_samples - prev.n_samples;
    const double ns = d.n_samples > 0 ? d.n_samples : 1.0, nt = cur.n_samples - first.n_samples;
    d.block_energy = (cur.sum_e - prev.sum_e) / ns;
    d.running_energy = (cur.sum_e - first.sum_e) / (nt > 0 ? nt : 1.0);
    d.block_kinetic = (cur.sum_t - prev.sum_t) / ns;
    d.block_wfvalue = (cur.sum_psi - prev.sum_psi) / ns;
    const double mv = cur.n_moves - prev.n_moves;
    d.acceptance = mv > 0 ? (cur.n_accept - prev.n_accept) / mv : 0.0;
    log(user, &d);
    prev = cur;
  }
  return MOLE_OK;
}

// block-size schedule of the reference's blocking analysis (scripts/statfor.rs:59-66):
// MIN_LEFT = 20, NSIZES = 100, sizes 1, 1+step, ... <= n/20 with step = max(n/20/100, 1)
int32_t mole_series_block_sizes(int64_t n, int32_t* sizes, int32_t* n_sizes) {
  if (!n_sizes || n < 0) return MOLE_ERR_INVALID_ARG;
  const int64_t large = n / 20, step = std::max<int64_t>(large / 100, 1);
  int32_t k = 0;
  for (int64_t s = 1; s <= large; s += step) {
    if (sizes) sizes[k] = (int32_t)s;
    ++k;
  }
  *n_sizes = k;
  return MOLE_OK;
}

// ------------------------------------------------------------------ VmcRunner::run_optimization (vmc.rs:43-106)
int32_t mole_vmc_run_optimization(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, mole_opt_t opt,
                                  const uint8_t master_seed[32], int32_t iters, int64_t total_samples, int32_t block_size,
                                  uint32_t compat, uint32_t flags, double* energies, double* errors, double* acceptance,
                                  double* param_history) {
  if (!ens || !wf || !m || !op || !opt || !master_seed || !energies || !errors) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = ens->ctx;
  const int np = wf->p.np;
  if (np < 1) return mole_set_error(ctx, MOLE_ERR_FUNC, "wavefunction kind has no Optimize impl");
  const int64_t nworkers = ens->W * (int64_t)ctx->nranks;
  const int64_t steps = total_samples / nworkers;   // :50
  if (steps > 0x7fffffff) return mole_set_error(ctx, MOLE_ERR_INVALID_ARG, "too many steps per walker");
  int32_t rc;
  if (flags & MOLE_VMC_RESTART_EACH_ITER)
    if ((rc = mole_ensemble_snapshot(ens)) != MOLE_OK) return rc;
  for (int it = 0; it < iters; ++it) {
    uint8_t seed[32];
    mole_derive_seed(master_seed, (uint32_t)it, seed);   // :59-61
    if ((rc = mole_ensemble_reseed(ens, seed)) != MOLE_OK) return rc;
    if (flags & MOLE_VMC_RESTART_EACH_ITER)
      if ((rc = mole_ensemble_restore(ens)) != MOLE_OK) return rc;   // every clone starts from the master cfg (:56)
    if ((rc = mole_acc_reset(ens)) != MOLE_OK) return rc;
    rc = mole_runner_run(ens, wf, m, op, MOLE_OBS_ENERGY | MOLE_OBS_PGRAD | MOLE_OBS_WFVALUE, compat, (int32_t)steps, block_size,
                         nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc != MOLE_OK) return rc;
    if (ctx->nranks > 1 && (rc = mole_acc_allreduce(ens)) != MOLE_OK) return rc;   // concatenate_worker_data (:108-130)
    mole_acc_host acc;
    if ((rc = mole_acc_get(ens, &acc)) != MOLE_OK) return rc;
    double e, err, accp;
    if ((rc = mole_acc_finalize(&acc, &e, &err, &accp, nullptr)) != MOLE_OK) return rc;   // :80-83
    energies[it] = e;
    errors[it] = err;
    if (acceptance) acceptance[it] = accp;
    double pars[MOLE_WF_MAX_PARAMS], dp[MOLE_WF_MAX_PARAMS];
    mole_wf_get_parameters(wf, pars);
    if (np > MOLE_ACC_MAX_PARAMS) {   // large-P kind: the moments are the Gram matrix of the per-sample rows
      if (ctx->nranks > 1 && (rc = mole_gram_allreduce(ens)) != MOLE_OK) return rc;
      int32_t nc = 0;
      std::vector<double> gram((size_t)(np + 2) * (np + 2));
      if ((rc = mole_gram_get(ens, &nc, gram.data())) != MOLE_OK) return rc;
      if ((rc = mole_opt_step_gram(opt, pars, nc, gram.data(), dp)) != MOLE_OK) return rc;
    } else
    if ((rc = mole_opt_step(opt, pars, &acc, dp)) != MOLE_OK) return rc;   // :85-89
    mole_wf_update_parameters(wf, dp);                                     // :91
    if (param_history) {
      mole_wf_get_parameters(wf, pars);
      for (int k = 0; k < np; ++k) param_history[(size_t)it * np + k] = pars[k];
    }
  }
  return MOLE_OK;
}

// ------------------------------------------------------------------ DmcRunner::diffuse (dmc.rs:69-203)
int32_t mole_dmc_diffuse(mole_ens_t ens, mole_wf_t wf, mole_metrop_t m, mole_op_t op, int32_t branch_kind, double time_step,
                         double* reference_energy, int32_t num_iterations, int32_t block_size, int32_t num_eq_blocks,
                         double* energies, double* errors, int32_t* n_out, double* step_energies) {
  if (!ens || !wf || !m || !op || !reference_energy || !energies || !errors || !n_out || block_size < 1)
    return MOLE_ERR_INVALID_ARG;
  const int blocks = num_iterations / block_size;
  std::vector<double> en, vars;
  double e_ref = *reference_energy;
  int32_t rc;
  int64_t t = 0;
  std::vector<double> se((size_t)block_size);
  for (int block_nr = 0; block_nr < blocks; ++block_nr) {
    // the whole block is enqueued without host reads: E_ref only changes between blocks (:143-145)
    if ((rc = mole_dmc_block(ens, wf, m, op, branch_kind, time_step, e_ref, block_size, se.data())) != MOLE_OK) return rc;
    // multi-rank: the ranks are population islands inside a block; when their weights per walker have drifted more than
    // 5 % apart they become one equal-weight population again (mole_rebalance).  Every rank takes the same decision.
    if (branch_kind == MOLE_BRANCH_SR && ens->ctx->nranks > 1) {
      double ratio = 1.0;
      if ((rc = mole_dmc_island_imbalance(ens, &ratio)) != MOLE_OK) return rc;
      if (ratio > MOLE_REBALANCE_RATIO && (rc = mole_rebalance(ens)) != MOLE_OK) return rc;
    }
    double sum = 0.0;
    for (int j = 0; j < block_size; ++j, ++t) {
      sum += se[j];                // :133-135
      if (step_energies) step_energies[t] = se[j];
    }
    const double energy = sum / (double)block_size;
    if (block_nr == num_eq_blocks) {   // :163-177
      e_ref = (e_ref + energy) / 2.0;
      en.push_back(energy);
      vars.push_back(0.0);
    }
    if (block_nr > num_eq_blocks) {   // :178-201
      const double prev = en.back();
      const double k = (double)(block_nr - num_eq_blocks);
      en.push_back(prev + (energy - prev) / k);
      e_ref = (e_ref + en.back()) / 2.0;
      vars.push_back(vars.back() + ((energy - prev) * (energy - en.back()) - vars.back()) / k);
    }
  }
  for (size_t i = 0; i < en.size(); ++i) {
    energies[i] = en[i];
    errors[i] = std::sqrt(vars[i] / (double)(i + 1));   // :148-151
  }
  *n_out = (int32_t)en.size();
  *reference_energy = e_ref;
  return MOLE_OK;
}

}  // extern "C"
