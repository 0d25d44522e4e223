// ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path (mole_b200/).
//
// CPU restatement of the reference's trial wavefunctions and local operators.
// The mole library ships no wavefunctions (README.md:33-37); the closed forms live
// in its examples/ and tests/, and each function below cites the lines it follows.
// Arithmetic is written in the reference's order (norm_l2 = sqrt of a sequential sum
// of squares, `powi(n)` = repeated multiplication, no FMA contraction: build with
// -ffp-contract=off).  Templated on the scalar so that (a) double gives the oracle,
// (b) orc::Counted gives the algorithmic op counts of DESIGN.md (F_alg).
//
// Parity status: the reference cannot be built here (Rust nightly + MKL, no toolchain,
// SURVEY.md §8(c)); this file is pinned against the golden vectors of SURVEY.md §8(c)
// and against 40-digit mpmath evaluations (tests/golden/make_golden.py).  The LCAO kinds (7-9) and the
// Slater-Jastrow kind have no upstream implementation at all: they are pinned by the mpmath fixtures and,
// for LCAO, by reduction to the reference's He and STO closed forms (tests/test_oracle_golden.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include <stdexcept>

namespace orc {

// ---------------------------------------------------------------- counting scalar
struct OpCounts { uint64_t add = 0, mul = 0, div = 0, sqrt = 0, exp = 0, log = 0, cmp = 0; };
inline OpCounts& op_counts() { static thread_local OpCounts c; return c; }

struct Counted {
  double v;
  Counted() : v(0) {}
  Counted(double x) : v(x) {}
  explicit operator double() const { return v; }
};
inline Counted operator+(Counted a, Counted b) { op_counts().add++; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { op_counts().add++; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { op_counts().mul++; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { op_counts().div++; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline Counted& operator+=(Counted& a, Counted b) { a = a + b; return a; }
inline Counted& operator-=(Counted& a, Counted b) { a = a - b; return a; }
inline Counted& operator*=(Counted& a, Counted b) { a = a * b; return a; }
inline bool operator>(Counted a, Counted b) { op_counts().cmp++; return a.v > b.v; }
inline bool operator<(Counted a, Counted b) { op_counts().cmp++; return a.v < b.v; }

inline double r_exp(double x) { return std::exp(x); }
inline double r_sqrt(double x) { return std::sqrt(x); }
inline double r_val(double x) { return x; }
inline Counted r_exp(Counted x) { op_counts().exp++; return Counted(std::exp(x.v)); }
inline Counted r_sqrt(Counted x) { op_counts().sqrt++; return Counted(std::sqrt(x.v)); }
inline double r_val(Counted x) { return x.v; }

// ---------------------------------------------------------------- descriptors (POD, ctypes-visible)
enum WfKind : int32_t {
  WF_STO_1S = 0,          // examples/dmc.rs:95-149
  WF_GAUSSIAN = 1,        // examples/dmc.rs:34-92, tests/sho_optimize.rs:52-111
  WF_STO_PRODUCT = 2,     // examples/helium_atom_singlet.rs:35-118, tests/helium_lcao.rs
  WF_H2_HL_STO = 3,       // examples/hydrogen_molecule.rs:65-168
  WF_H2P_PRODUCT = 4,     // tests/hydrogen_molecular_ion_lcao.rs:51-98
  WF_SLATER_JASTROW = 5,  // SURVEY.md §8(c) synthetic config 5 + theory/jastrow.tex (erratum applied)
  WF_CONSTANT = 6,        // src/metropolis/src/metrop.rs:225-255 (WaveFunctionMock)
  // LCAO determinants over a hydrogen-1s basis: the API the reference's tests name but keep commented out
  // (tests/helium_lcao.rs:94-101, tests/hydrogen_molecular_ion_lcao.rs:103-107).  No upstream implementation exists;
  // the restatement is built from the reference's STO (hydrogen_molecular_ion_lcao.rs:25-49) shifted to each centre.
  WF_LCAO_1E_2C = 7,      // SingleDeterminant([phi_0]) over two centres (H2+)
  WF_LCAO_2E_1C = 8,      // two orbitals over one centre (He)
  WF_LCAO_2E_2C = 9,      // two orbitals over two centres (H2 molecular orbitals)
  // General case of the same API: SpinDeterminantProduct of n_orb = max(n_up, n_dn) LCAO orbitals
  // phi_k(r) = sum_c C[k][c] exp(-alpha_c |r - R_c|) over N_c <= 8 centres, both spins sharing the orbitals, times the
  // e-e Jastrow of theory/jastrow.tex (as WF_SLATER_JASTROW).  geom: [kappa, n_up, n_dn, N_c, -, -, -, -,
  // (R_c x, y, z, alpha_c) for c < N_c]; params: C[k][c] at k N_c + c, then b1..b4 (P = n_orb N_c + 4 <= 44).
  WF_LCAO_SJ = 10,
};

constexpr int WF_MAX_PARAMS = 48;
constexpr int WF_MAX_GEOM = 40;
constexpr int LSJ_MAX_CENTRES = 8;
constexpr int WF_MAX_ELEC = 10;

struct WfDesc {
  int32_t kind;
  int32_t n_elec;
  int32_t n_params;
  int32_t reserved;
  double params[WF_MAX_PARAMS];  // variational parameters (Optimize::parameters)
  double geom[WF_MAX_GEOM];      // fixed constants: H2/H2P: [R]; SJ: [kappa, n_up, n_dn]; CONSTANT: [value]
};

enum HamKind : int32_t {
  HAM_KINETIC = 0,       // src/operator/src/operator.rs:103-125
  HAM_IONIC_POT = 1,     // operator.rs:16-62
  HAM_ELEC_POT = 2,      // operator.rs:68-97
  HAM_IONIC = 3,         // operator.rs:130-149
  HAM_ELECTRONIC = 4,    // operator.rs:156-184
  HAM_HARMONIC = 5,      // examples/custom_operator.rs:30-61, tests/sho_optimize.rs:18-50
};
constexpr int HAM_MAX_IONS = 8;

struct HamDesc {
  int32_t kind;
  int32_t n_ions;
  double ion_pos[HAM_MAX_IONS * 3];
  int32_t ion_charge[HAM_MAX_IONS];
  double frequency;
};

// ---------------------------------------------------------------- helpers
// ndarray-linalg `norm_l2`: sqrt(sum_i x_i^2), sequential (SURVEY.md §8(c) third-party table).
template <class R>
inline R norm_l2(const R* x, int n) {
  R s = R(0.0);
  for (int i = 0; i < n; ++i) s += x[i] * x[i];
  return r_sqrt(s);
}

template <class R>
struct Wf {
  int kind, ne, np;
  R p[WF_MAX_PARAMS];
  double geom[WF_MAX_GEOM];
  explicit Wf(const WfDesc& d) : kind(d.kind), ne(d.n_elec), np(d.n_params) {
    for (int i = 0; i < WF_MAX_PARAMS; ++i) p[i] = R(d.params[i]);
    for (int i = 0; i < WF_MAX_GEOM; ++i) geom[i] = d.geom[i];
  }
};

// ---- single STO  (hydrogen_molecule.rs:38-62; hydrogen_molecular_ion_lcao.rs:25-49; dmc.rs:106-143)
template <class R> inline R sto_value(R alpha, const R* x) { return r_exp(-alpha * norm_l2(x, 3)); }
template <class R> inline void sto_gradient(R alpha, const R* x, R* g) {
  // -alpha * value / |x| * x          (hydrogen_molecule.rs:47-49)
  const R c = -alpha * sto_value(alpha, x) / norm_l2(x, 3);
  for (int k = 0; k < 3; ++k) g[k] = c * x[k];
}
template <class R> inline R sto_laplacian(R alpha, const R* x) {
  // alpha * value / |x| * (alpha*|x| - 2)   (hydrogen_molecule.rs:51-53)
  return alpha * sto_value(alpha, x) / norm_l2(x, 3) * (alpha * norm_l2(x, 3) - R(2.0));
}
template <class R> inline R sto_pgrad(R alpha, const R* x) {
  return -norm_l2(x, 3) * sto_value(alpha, x);  // hydrogen_molecule.rs:55-57
}

// ---- LCAO over a hydrogen-1s basis (kinds 7-9): chi_c(r) = STO(alpha)(r - R_c), phi_k = sum_c C[k][c] chi_c.
// geom: [mode, alpha = 1/width, R_0 (3), R_1 (3)]; params: C[k][c] at k * nc + c.  mode 0: psi = phi_0(x_0) phi_1(x_1)
// (SpinDeterminantProduct, n_up = 1); mode 1: psi = phi_0(x_0) phi_1(x_1) - phi_0(x_1) phi_1(x_0) (SingleDeterminant).
inline bool wf_is_lcao(int kind) { return kind >= WF_LCAO_1E_2C && kind <= WF_LCAO_2E_2C; }
inline int lcao_nc(int kind) { return kind == WF_LCAO_2E_1C ? 1 : 2; }
template <class R>
struct LcaoOrb {   // orbital k evaluated at one electron position
  R v, g[3], l;
};
template <class R>
inline LcaoOrb<R> lcao_orbital(const Wf<R>& wf, int k, const R* x) {
  const int nc = lcao_nc(wf.kind);
  const R al = R(wf.geom[1]);
  LcaoOrb<R> o;
  o.v = R(0.0); o.l = R(0.0);
  for (int q = 0; q < 3; ++q) o.g[q] = R(0.0);
  for (int c = 0; c < nc; ++c) {
    R d[3], gc[3];
    for (int q = 0; q < 3; ++q) d[q] = x[q] - R(wf.geom[2 + 3 * c + q]);
    const R C = wf.p[k * nc + c];
    sto_gradient(al, d, gc);
    o.v += C * sto_value(al, d);
    o.l += C * sto_laplacian(al, d);
    for (int q = 0; q < 3; ++q) o.g[q] += C * gc[q];
  }
  return o;
}
template <class R>
inline R lcao_chi(const Wf<R>& wf, int c, const R* x) {
  R d[3];
  for (int q = 0; q < 3; ++q) d[q] = x[q] - R(wf.geom[2 + 3 * c + q]);
  return sto_value(R(wf.geom[1]), d);
}

// ---- small dense determinant (partial pivoting), used only by the Slater-Jastrow oracle
template <class R>
inline R det_n(const R* a_in, int n) {
  R a[25];
  for (int i = 0; i < n * n; ++i) a[i] = a_in[i];
  R det = R(1.0);
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = std::fabs(r_val(a[c * n + c]));
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(r_val(a[r * n + c])) > best) { best = std::fabs(r_val(a[r * n + c])); piv = r; }
    if (best == 0.0) return R(0.0);
    if (piv != c) {
      for (int k = 0; k < n; ++k) { R t = a[c * n + k]; a[c * n + k] = a[piv * n + k]; a[piv * n + k] = t; }
      det = -det;
    }
    det *= a[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      const R f = a[r * n + c] / a[c * n + c];
      for (int k = c + 1; k < n; ++k) a[r * n + k] -= f * a[c * n + k];
    }
  }
  return det;
}

// ---- Slater-Jastrow building blocks (SURVEY.md §8(c) "Synthetic config 5")
// Orbital set: 0: exp(-z1 r)  1: r exp(-z2 r)  2,3,4: (x,y,z) exp(-z3 r).
// out: val[5], grad[5][3], lap[5], dzeta[5] (derivative w.r.t. the orbital's own exponent)
template <class R>
inline void sj_orbitals(const R* zeta, const R* x, int norb, R* val, R* grad, R* lap, R* dz) {
  const R r = norm_l2(x, 3);
  const R e1 = r_exp(-zeta[0] * r), e2 = r_exp(-zeta[1] * r), e3 = r_exp(-zeta[2] * r);
  for (int k = 0; k < norb; ++k) {
    if (k == 0) {
      val[0] = e1;
      for (int c = 0; c < 3; ++c) grad[c] = -zeta[0] * e1 * x[c] / r;
      lap[0] = zeta[0] * e1 * (zeta[0] - R(2.0) / r);
      dz[0] = -r * e1;
    } else if (k == 1) {
      val[1] = r * e2;
      for (int c = 0; c < 3; ++c) grad[3 + c] = (R(1.0) - zeta[1] * r) * e2 * x[c] / r;
      lap[1] = (zeta[1] * zeta[1] * r - R(4.0) * zeta[1] + R(2.0) / r) * e2;
      dz[1] = -r * r * e2;
    } else {
      const int a = k - 2;  // cartesian axis of the p orbital
      val[k] = x[a] * e3;
      for (int c = 0; c < 3; ++c) {
        R t = -zeta[2] * x[a] * x[c] / r;
        if (c == a) t = t + R(1.0);
        grad[3 * k + c] = e3 * t;
      }
      lap[k] = x[a] * e3 * (zeta[2] * zeta[2] - R(4.0) * zeta[2] / r);
      dz[k] = -r * x[a] * e3;
    }
  }
}

// Jastrow pair function in the scaled distance R=(1-exp(-kappa r))/kappa (theory/jastrow.tex:23-31)
template <class R>
struct JPair { R u, g, dg, Rs, e; };  // u(r), g=du/dr, dg=dg/dr, scaled distance, exp(-kappa r)
template <class R>
inline JPair<R> sj_pair(const R* b, double kappa, R r) {
  JPair<R> o;
  o.e = r_exp(-R(kappa) * r);
  o.Rs = (R(1.0) - o.e) / R(kappa);
  const R den = R(1.0) + b[1] * o.Rs;
  o.u = b[0] * o.Rs / den + b[2] * o.Rs * o.Rs + b[3] * o.Rs * o.Rs * o.Rs;
  const R du = b[0] / (den * den) + R(2.0) * b[2] * o.Rs + R(3.0) * b[3] * o.Rs * o.Rs;   // jastrow.tex:45-48
  const R d2u = -R(2.0) * b[0] * b[1] / (den * den * den) + R(2.0) * b[2] + R(6.0) * b[3] * o.Rs;
  o.g = o.e * du;                                   // jastrow.tex:68-71
  o.dg = -R(kappa) * o.g + o.e * o.e * d2u;         // jastrow.tex:82-88 with the erratum of SURVEY.md §8(c)
  return o;
}

template <class R>
struct SjParts {
  R det[2];                    // D_up, D_dn
  R ddet[WF_MAX_ELEC][3];      // grad_i D_s(i)   (row-replacement determinants)
  R d2det[WF_MAX_ELEC];        // lap_i  D_s(i)
  R dzdet[2][3];               // dD_s/dzeta_m
  R f;                         // Jastrow exponent
  R gf[WF_MAX_ELEC][3];        // grad_i f
  R lf[WF_MAX_ELEC];           // lap_i f
  R dbf[4];                    // df/db_m
};

template <class R>
inline void sj_parts(const Wf<R>& wf, const R* cfg, SjParts<R>& o) {
  const double kappa = wf.geom[0];
  const int nup = (int)wf.geom[1], ndn = (int)wf.geom[2];
  const R* zeta = wf.p;
  const R* b = wf.p + 3;
  const int nspin[2] = {nup, ndn};
  const int first[2] = {0, nup};
  for (int s = 0; s < 2; ++s) {
    const int n = nspin[s];
    R A[25], G[5][15], L[25], Z[25];
    for (int i = 0; i < n; ++i) {
      R val[5], grad[15], lap[5], dz[5];
      sj_orbitals(zeta, cfg + 3 * (first[s] + i), n, val, grad, lap, dz);
      for (int k = 0; k < n; ++k) {
        A[i * n + k] = val[k]; L[i * n + k] = lap[k]; Z[i * n + k] = dz[k];
        for (int c = 0; c < 3; ++c) G[i][3 * k + c] = grad[3 * k + c];
      }
    }
    o.det[s] = (n == 0) ? R(1.0) : det_n(A, n);
    for (int i = 0; i < n; ++i) {
      R B[25];
      for (int c = 0; c < 3; ++c) {
        for (int q = 0; q < n * n; ++q) B[q] = A[q];
        for (int k = 0; k < n; ++k) B[i * n + k] = G[i][3 * k + c];
        o.ddet[first[s] + i][c] = det_n(B, n);
      }
      for (int q = 0; q < n * n; ++q) B[q] = A[q];
      for (int k = 0; k < n; ++k) B[i * n + k] = L[i * n + k];
      o.d2det[first[s] + i] = det_n(B, n);
    }
    // dD/dzeta_m: D is multilinear in its columns; replace each column that depends on zeta_m
    for (int m = 0; m < 3; ++m) {
      R acc = R(0.0);
      for (int k = 0; k < n; ++k) {
        const int owner = (k == 0) ? 0 : (k == 1 ? 1 : 2);
        if (owner != m) continue;
        R B[25];
        for (int q = 0; q < n * n; ++q) B[q] = A[q];
        for (int i = 0; i < n; ++i) B[i * n + k] = Z[i * n + k];
        acc += det_n(B, n);
      }
      o.dzdet[s][m] = acc;
    }
  }
  const int ne = wf.ne;
  o.f = R(0.0);
  for (int m = 0; m < 4; ++m) o.dbf[m] = R(0.0);
  for (int i = 0; i < ne; ++i) { o.lf[i] = R(0.0); for (int c = 0; c < 3; ++c) o.gf[i][c] = R(0.0); }
  for (int i = 0; i < ne; ++i)
    for (int j = i + 1; j < ne; ++j) {
      R d[3];
      for (int c = 0; c < 3; ++c) d[c] = cfg[3 * i + c] - cfg[3 * j + c];
      const R r = norm_l2(d, 3);
      const JPair<R> pr = sj_pair(b, kappa, r);
      o.f += pr.u;
      for (int c = 0; c < 3; ++c) {
        const R t = pr.g * d[c] / r;
        o.gf[i][c] += t;
        o.gf[j][c] -= t;
      }
      const R lp = pr.g * R(2.0) / r + pr.dg;   // div(rhat g), jastrow.tex:91-97 corrected
      o.lf[i] += lp;
      o.lf[j] += lp;
      const R den = R(1.0) + b[1] * pr.Rs;
      o.dbf[0] += pr.Rs / den;
      o.dbf[1] -= b[0] * pr.Rs * pr.Rs / (den * den);      // jastrow.tex:109-112
      o.dbf[2] += pr.Rs * pr.Rs;                             // jastrow.tex:116-119
      o.dbf[3] += pr.Rs * pr.Rs * pr.Rs;
    }
}

// ---- general LCAO Slater-Jastrow (WF_LCAO_SJ): orbitals from the hydrogen-1s basis, determinants and Jastrow as above
template <class R>
struct LsjParts {
  R det[2];
  R ddet[WF_MAX_ELEC][3];
  R d2det[WF_MAX_ELEC];
  R dcdet[2][5 * LSJ_MAX_CENTRES];   // dD_s / dC[k][c]
  R f;
  R gf[WF_MAX_ELEC][3];
  R lf[WF_MAX_ELEC];
  R dbf[4];
};

inline int lsj_nc(const double* geom) { return (int)geom[3]; }
inline int lsj_norb(const double* geom) { return std::max((int)geom[1], (int)geom[2]); }

// chi_c and its gradient / laplacian at x: the reference's STO (hydrogen_molecular_ion_lcao.rs:25-49) at centre c
template <class R>
inline void lsj_basis(const Wf<R>& wf, const R* x, R* chi, R* gchi, R* lchi) {
  const int nc = lsj_nc(wf.geom);
  for (int c = 0; c < nc; ++c) {
    const double* g = wf.geom + 8 + 4 * c;
    const R al = R(g[3]);
    R d[3] = {x[0] - R(g[0]), x[1] - R(g[1]), x[2] - R(g[2])};
    const R r = norm_l2(d, 3);
    chi[c] = r_exp(-al * r);
    for (int q = 0; q < 3; ++q) gchi[3 * c + q] = -al * chi[c] * d[q] / r;
    lchi[c] = al * chi[c] * (al - R(2.0) / r);
  }
}

template <class R>
inline void lsj_parts(const Wf<R>& wf, const R* cfg, LsjParts<R>& o) {
  const double kappa = wf.geom[0];
  const int nup = (int)wf.geom[1], ndn = (int)wf.geom[2], nc = lsj_nc(wf.geom), norb = lsj_norb(wf.geom);
  const R* Cm = wf.p;                    // C[k][c]
  const R* b = wf.p + norb * nc;
  const int nspin[2] = {nup, ndn};
  const int first[2] = {0, nup};
  for (int s = 0; s < 2; ++s) {
    const int n = nspin[s];
    R A[25], G[5][15], L[25], X[5][LSJ_MAX_CENTRES];
    for (int i = 0; i < n; ++i) {
      R chi[LSJ_MAX_CENTRES], gchi[3 * LSJ_MAX_CENTRES], lchi[LSJ_MAX_CENTRES];
      lsj_basis(wf, cfg + 3 * (first[s] + i), chi, gchi, lchi);
      for (int c = 0; c < nc; ++c) X[i][c] = chi[c];
      for (int k = 0; k < n; ++k) {
        R v = R(0.0), l = R(0.0), g[3] = {R(0.0), R(0.0), R(0.0)};
        for (int c = 0; c < nc; ++c) {
          v += Cm[k * nc + c] * chi[c];
          l += Cm[k * nc + c] * lchi[c];
          for (int q = 0; q < 3; ++q) g[q] += Cm[k * nc + c] * gchi[3 * c + q];
        }
        A[i * n + k] = v; L[i * n + k] = l;
        for (int q = 0; q < 3; ++q) G[i][3 * k + q] = g[q];
      }
    }
    o.det[s] = (n == 0) ? R(1.0) : det_n(A, n);
    for (int i = 0; i < n; ++i) {
      R B[25];
      for (int q3 = 0; q3 < 3; ++q3) {
        for (int q = 0; q < n * n; ++q) B[q] = A[q];
        for (int k = 0; k < n; ++k) B[i * n + k] = G[i][3 * k + q3];
        o.ddet[first[s] + i][q3] = det_n(B, n);
      }
      for (int q = 0; q < n * n; ++q) B[q] = A[q];
      for (int k = 0; k < n; ++k) B[i * n + k] = L[i * n + k];
      o.d2det[first[s] + i] = det_n(B, n);
    }
    // dD/dC[k][c]: D is linear in column k, whose derivative is the column chi_c(r_i)
    for (int k = 0; k < norb; ++k)
      for (int c = 0; c < nc; ++c) {
        if (k >= n) { o.dcdet[s][k * nc + c] = R(0.0); continue; }
        R B[25];
        for (int q = 0; q < n * n; ++q) B[q] = A[q];
        for (int i = 0; i < n; ++i) B[i * n + k] = X[i][c];
        o.dcdet[s][k * nc + c] = det_n(B, n);
      }
  }
  const int ne = wf.ne;
  o.f = R(0.0);
  for (int m = 0; m < 4; ++m) o.dbf[m] = R(0.0);
  for (int i = 0; i < ne; ++i) { o.lf[i] = R(0.0); for (int c = 0; c < 3; ++c) o.gf[i][c] = R(0.0); }
  for (int i = 0; i < ne; ++i)
    for (int j = i + 1; j < ne; ++j) {
      R d[3];
      for (int c = 0; c < 3; ++c) d[c] = cfg[3 * i + c] - cfg[3 * j + c];
      const R r = norm_l2(d, 3);
      const JPair<R> pr = sj_pair(b, kappa, r);
      o.f += pr.u;
      for (int c = 0; c < 3; ++c) {
        const R t = pr.g * d[c] / r;
        o.gf[i][c] += t;
        o.gf[j][c] -= t;
      }
      const R lp = pr.g * R(2.0) / r + pr.dg;
      o.lf[i] += lp;
      o.lf[j] += lp;
      const R den = R(1.0) + b[1] * pr.Rs;
      o.dbf[0] += pr.Rs / den;
      o.dbf[1] -= b[0] * pr.Rs * pr.Rs / (den * den);
      o.dbf[2] += pr.Rs * pr.Rs;
      o.dbf[3] += pr.Rs * pr.Rs * pr.Rs;
    }
}

// ---------------------------------------------------------------- Function::value
template <class R>
R wf_value(const Wf<R>& wf, const R* cfg) {
  switch (wf.kind) {
    case WF_STO_1S:  // dmc.rs:108-111  (norm over the whole (1,3) array)
      return r_exp(-wf.p[0] * norm_l2(cfg, 3));
    case WF_GAUSSIAN: {  // dmc.rs:47-50: exp(-(|x|/a)^2)
      const R q = norm_l2(cfg, 3 * wf.ne) / wf.p[0];
      return r_exp(-(q * q));
    }
    case WF_STO_PRODUCT: {  // helium_atom_singlet.rs:63-70
      return r_exp(-wf.p[0] * (norm_l2(cfg, 3) + norm_l2(cfg + 3, 3)));
    }
    case WF_H2_HL_STO: {  // hydrogen_molecule.rs:91-100
      const R h = R(0.5) * R(wf.geom[0]);
      R a1[3] = {cfg[0] - h, cfg[1] - R(0.0), cfg[2] - R(0.0)};
      R b1[3] = {cfg[0] + h, cfg[1] + R(0.0), cfg[2] + R(0.0)};
      R a2[3] = {cfg[3] - h, cfg[4] - R(0.0), cfg[5] - R(0.0)};
      R b2[3] = {cfg[3] + h, cfg[4] + R(0.0), cfg[5] + R(0.0)};
      const R al = wf.p[0];
      return sto_value(al, a1) * sto_value(al, b2) + sto_value(al, b1) * sto_value(al, a2);
    }
    case WF_H2P_PRODUCT: {  // hydrogen_molecular_ion_lcao.rs:69-73
      const R h = R(0.5) * R(wf.geom[0]);
      R r1[3] = {cfg[0] - h, cfg[1], cfg[2]};
      R r2[3] = {cfg[0] + h, cfg[1], cfg[2]};
      return sto_value(wf.p[0], r1) * sto_value(wf.p[0], r2);
    }
    case WF_SLATER_JASTROW: {
      SjParts<R> s;
      sj_parts(wf, cfg, s);
      return s.det[0] * s.det[1] * r_exp(s.f);
    }
    case WF_LCAO_SJ: {
      LsjParts<R> s;
      lsj_parts(wf, cfg, s);
      return s.det[0] * s.det[1] * r_exp(s.f);
    }
    case WF_CONSTANT:  // metrop.rs:240-242
      return R(wf.geom[0]);
    case WF_LCAO_1E_2C:
      return lcao_orbital(wf, 0, cfg).v;
    case WF_LCAO_2E_1C:
    case WF_LCAO_2E_2C: {
      const R direct = lcao_orbital(wf, 0, cfg).v * lcao_orbital(wf, 1, cfg + 3).v;
      if (wf.geom[0] == 0.0) return direct;
      return direct - lcao_orbital(wf, 0, cfg + 3).v * lcao_orbital(wf, 1, cfg).v;
    }
  }
  throw std::runtime_error("wf_value: unknown kind");
}

// ---------------------------------------------------------------- Differentiate::gradient  (un-normalised)
template <class R>
void wf_gradient(const Wf<R>& wf, const R* cfg, R* out) {
  switch (wf.kind) {
    case WF_STO_1S: {  // dmc.rs:116-119: -alpha * value / |x| * x
      const R c = -wf.p[0] * wf_value(wf, cfg) / norm_l2(cfg, 3);
      for (int k = 0; k < 3; ++k) out[k] = c * cfg[k];
      return;
    }
    case WF_GAUSSIAN: {  // dmc.rs:56-59: -2 * value / a^2 * x
      const R a = wf.p[0];
      const R c = R(-2.0) * wf_value(wf, cfg) / (a * a);
      for (int k = 0; k < 3 * wf.ne; ++k) out[k] = c * cfg[k];
      return;
    }
    case WF_STO_PRODUCT: {  // helium_atom_singlet.rs:76-88
      const R al = wf.p[0];
      const R value = wf_value(wf, cfg);
      for (int e = 0; e < 2; ++e) {
        const R c = -al / norm_l2(cfg + 3 * e, 3) * value;
        for (int k = 0; k < 3; ++k) out[3 * e + k] = c * cfg[3 * e + k];
      }
      return;
    }
    case WF_H2_HL_STO: {  // hydrogen_molecule.rs:106-122
      const R h = R(0.5) * R(wf.geom[0]);
      R a1[3] = {cfg[0] - h, cfg[1] - R(0.0), cfg[2] - R(0.0)};
      R b1[3] = {cfg[0] + h, cfg[1] + R(0.0), cfg[2] + R(0.0)};
      R a2[3] = {cfg[3] - h, cfg[4] - R(0.0), cfg[5] - R(0.0)};
      R b2[3] = {cfg[3] + h, cfg[4] + R(0.0), cfg[5] + R(0.0)};
      const R al = wf.p[0];
      R ga1[3], gb1[3], ga2[3], gb2[3];
      sto_gradient(al, a1, ga1); sto_gradient(al, b1, gb1);
      sto_gradient(al, a2, ga2); sto_gradient(al, b2, gb2);
      const R vb2 = sto_value(al, b2), va2 = sto_value(al, a2);
      const R vb1 = sto_value(al, b1), va1 = sto_value(al, a1);
      for (int k = 0; k < 3; ++k) {
        out[k] = R(0.0) + (vb2 * ga1[k] + va2 * gb1[k]);       // :114-115
        out[3 + k] = R(0.0) + (vb1 * ga2[k] + va1 * gb2[k]);   // :118-119
      }
      return;
    }
    case WF_H2P_PRODUCT: {  // hydrogen_molecular_ion_lcao.rs:79-83
      const R h = R(0.5) * R(wf.geom[0]);
      R r1[3] = {cfg[0] - h, cfg[1], cfg[2]};
      R r2[3] = {cfg[0] + h, cfg[1], cfg[2]};
      R g1[3], g2[3];
      sto_gradient(wf.p[0], r1, g1); sto_gradient(wf.p[0], r2, g2);
      const R v1 = sto_value(wf.p[0], r1), v2 = sto_value(wf.p[0], r2);
      for (int k = 0; k < 3; ++k) out[k] = v1 * g2[k] + v2 * g1[k];
      return;
    }
    case WF_SLATER_JASTROW: {
      SjParts<R> s;
      sj_parts(wf, cfg, s);
      const int nup = (int)wf.geom[1];
      const R J = r_exp(s.f);
      for (int i = 0; i < wf.ne; ++i) {
        const R other = (i < nup) ? s.det[1] : s.det[0];
        for (int c = 0; c < 3; ++c)
          out[3 * i + c] = s.ddet[i][c] * other * J + s.det[0] * s.det[1] * J * s.gf[i][c];
      }
      return;
    }
    case WF_LCAO_SJ: {
      LsjParts<R> s;
      lsj_parts(wf, cfg, s);
      const int nup = (int)wf.geom[1];
      const R J = r_exp(s.f);
      for (int i = 0; i < wf.ne; ++i) {
        const R other = (i < nup) ? s.det[1] : s.det[0];
        for (int c = 0; c < 3; ++c)
          out[3 * i + c] = s.ddet[i][c] * other * J + s.det[0] * s.det[1] * J * s.gf[i][c];
      }
      return;
    }
    case WF_LCAO_1E_2C: {
      const LcaoOrb<R> a = lcao_orbital(wf, 0, cfg);
      for (int q = 0; q < 3; ++q) out[q] = a.g[q];
      return;
    }
    case WF_LCAO_2E_1C:
    case WF_LCAO_2E_2C: {
      const LcaoOrb<R> a0 = lcao_orbital(wf, 0, cfg), b1 = lcao_orbital(wf, 1, cfg + 3);
      for (int q = 0; q < 3; ++q) { out[q] = a0.g[q] * b1.v; out[3 + q] = a0.v * b1.g[q]; }
      if (wf.geom[0] != 0.0) {
        const LcaoOrb<R> a1 = lcao_orbital(wf, 0, cfg + 3), b0 = lcao_orbital(wf, 1, cfg);
        for (int q = 0; q < 3; ++q) { out[q] -= a1.v * b0.g[q]; out[3 + q] -= a1.g[q] * b0.v; }
      }
      return;
    }
    case WF_CONSTANT:  // metrop.rs:248-250: unimplemented!()
      throw std::runtime_error("WaveFunctionMock::gradient is unimplemented in the reference");
  }
  throw std::runtime_error("wf_gradient: unknown kind");
}

// ---------------------------------------------------------------- Differentiate::laplacian  (un-normalised, summed over electrons)
template <class R>
R wf_laplacian(const Wf<R>& wf, const R* cfg) {
  switch (wf.kind) {
    case WF_STO_1S: {  // dmc.rs:121-124
      const R al = wf.p[0];
      return al * wf_value(wf, cfg) / norm_l2(cfg, 3) * (al * norm_l2(cfg, 3) - R(2.0));
    }
    case WF_GAUSSIAN: {  // dmc.rs:61-64
      const R a = wf.p[0];
      const R n = norm_l2(cfg, 3 * wf.ne);
      return wf_value(wf, cfg) * (R(4.0) * (n * n) - R(6.0) * (a * a)) / ((a * a) * (a * a));
    }
    case WF_STO_PRODUCT: {  // helium_atom_singlet.rs:90-98
      const R al = wf.p[0];
      const R val = wf_value(wf, cfg);
      const R n1 = norm_l2(cfg, 3), n2 = norm_l2(cfg + 3, 3);
      return al / n1 * val * (al * n1 - R(2.0)) + al / n2 * val * (al * n2 - R(2.0));
    }
    case WF_H2_HL_STO: {  // hydrogen_molecule.rs:124-135
      const R h = R(0.5) * R(wf.geom[0]);
      R a1[3] = {cfg[0] - h, cfg[1] - R(0.0), cfg[2] - R(0.0)};
      R b1[3] = {cfg[0] + h, cfg[1] + R(0.0), cfg[2] + R(0.0)};
      R a2[3] = {cfg[3] - h, cfg[4] - R(0.0), cfg[5] - R(0.0)};
      R b2[3] = {cfg[3] + h, cfg[4] + R(0.0), cfg[5] + R(0.0)};
      const R al = wf.p[0];
      return (sto_value(al, b2) * sto_laplacian(al, a1) + sto_value(al, a2) * sto_laplacian(al, b1)) +
             (sto_value(al, b1) * sto_laplacian(al, a2) + sto_value(al, a1) * sto_laplacian(al, b2));
    }
    case WF_H2P_PRODUCT: {  // hydrogen_molecular_ion_lcao.rs:85-91
      const R h = R(0.5) * R(wf.geom[0]);
      R r1[3] = {cfg[0] - h, cfg[1], cfg[2]};
      R r2[3] = {cfg[0] + h, cfg[1], cfg[2]};
      R g1[3], g2[3];
      sto_gradient(wf.p[0], r1, g1); sto_gradient(wf.p[0], r2, g2);
      R dot = R(0.0);
      for (int k = 0; k < 3; ++k) dot += g1[k] * g2[k];
      return sto_value(wf.p[0], r1) * sto_laplacian(wf.p[0], r2) +
             sto_value(wf.p[0], r2) * sto_laplacian(wf.p[0], r1) + R(2.0) * dot;
    }
    case WF_SLATER_JASTROW: {
      SjParts<R> s;
      sj_parts(wf, cfg, s);
      const int nup = (int)wf.geom[1];
      const R J = r_exp(s.f);
      const R DD = s.det[0] * s.det[1];
      R lap = R(0.0);
      for (int i = 0; i < wf.ne; ++i) {
        const R other = (i < nup) ? s.det[1] : s.det[0];
        R cross = R(0.0), g2 = R(0.0);
        for (int c = 0; c < 3; ++c) { cross += s.ddet[i][c] * s.gf[i][c]; g2 += s.gf[i][c] * s.gf[i][c]; }
        lap += s.d2det[i] * other * J + R(2.0) * cross * other * J + DD * J * (s.lf[i] + g2);
      }
      return lap;
    }
    case WF_LCAO_SJ: {
      LsjParts<R> s;
      lsj_parts(wf, cfg, s);
      const int nup = (int)wf.geom[1];
      const R J = r_exp(s.f);
      const R DD = s.det[0] * s.det[1];
      R lap = R(0.0);
      for (int i = 0; i < wf.ne; ++i) {
        const R other = (i < nup) ? s.det[1] : s.det[0];
        R cross = R(0.0), g2 = R(0.0);
        for (int c = 0; c < 3; ++c) { cross += s.ddet[i][c] * s.gf[i][c]; g2 += s.gf[i][c] * s.gf[i][c]; }
        lap += s.d2det[i] * other * J + R(2.0) * cross * other * J + DD * J * (s.lf[i] + g2);
      }
      return lap;
    }
    case WF_LCAO_1E_2C:
      return lcao_orbital(wf, 0, cfg).l;
    case WF_LCAO_2E_1C:
    case WF_LCAO_2E_2C: {
      const LcaoOrb<R> a0 = lcao_orbital(wf, 0, cfg), b1 = lcao_orbital(wf, 1, cfg + 3);
      const R direct = a0.l * b1.v + a0.v * b1.l;
      if (wf.geom[0] == 0.0) return direct;
      const LcaoOrb<R> a1 = lcao_orbital(wf, 0, cfg + 3), b0 = lcao_orbital(wf, 1, cfg);
      return direct - (a1.v * b0.l + a1.l * b0.v);
    }
    case WF_CONSTANT:  // metrop.rs:252-254
      return R(1.0);
  }
  throw std::runtime_error("wf_laplacian: unknown kind");
}

// ---------------------------------------------------------------- Optimize::parameter_gradient  (un-normalised d psi / d p_k)
template <class R>
void wf_parameter_gradient(const Wf<R>& wf, const R* cfg, R* out) {
  switch (wf.kind) {
    case WF_STO_1S:  // dmc.rs:128-130
      out[0] = -norm_l2(cfg, 3) * wf_value(wf, cfg);
      return;
    case WF_GAUSSIAN: {  // dmc.rs:74-79
      const R a = wf.p[0];
      const R n = norm_l2(cfg, 3 * wf.ne);
      out[0] = wf_value(wf, cfg) * R(2.0) * (n * n) / (a * a * a);
      return;
    }
    case WF_STO_PRODUCT:  // helium_atom_singlet.rs:102-105
      out[0] = -wf_value(wf, cfg) * (norm_l2(cfg, 3) + norm_l2(cfg + 3, 3));
      return;
    case WF_H2_HL_STO: {  // hydrogen_molecule.rs:139-154
      const R h = R(0.5) * R(wf.geom[0]);
      R a1[3] = {cfg[0] - h, cfg[1] - R(0.0), cfg[2] - R(0.0)};
      R b1[3] = {cfg[0] + h, cfg[1] + R(0.0), cfg[2] + R(0.0)};
      R a2[3] = {cfg[3] - h, cfg[4] - R(0.0), cfg[5] - R(0.0)};
      R b2[3] = {cfg[3] + h, cfg[4] + R(0.0), cfg[5] + R(0.0)};
      const R al = wf.p[0];
      out[0] = sto_value(al, a1) * sto_pgrad(al, b2) + sto_pgrad(al, a1) * sto_value(al, b2) +
               sto_value(al, b1) * sto_pgrad(al, a2) + sto_pgrad(al, b1) * sto_value(al, a2);
      return;
    }
    case WF_H2P_PRODUCT:  // no Optimize impl in tests/hydrogen_molecular_ion_lcao.rs
      return;
    case WF_SLATER_JASTROW: {
      SjParts<R> s;
      sj_parts(wf, cfg, s);
      const R J = r_exp(s.f);
      for (int m = 0; m < 3; ++m)
        out[m] = (s.dzdet[0][m] * s.det[1] + s.det[0] * s.dzdet[1][m]) * J;
      for (int m = 0; m < 4; ++m) out[3 + m] = s.det[0] * s.det[1] * J * s.dbf[m];  // jastrow.tex:123-126
      return;
    }
    case WF_LCAO_SJ: {
      LsjParts<R> s;
      lsj_parts(wf, cfg, s);
      const int nco = lsj_norb(wf.geom) * lsj_nc(wf.geom);
      const R J = r_exp(s.f);
      for (int m = 0; m < nco; ++m) out[m] = (s.dcdet[0][m] * s.det[1] + s.det[0] * s.dcdet[1][m]) * J;
      for (int m = 0; m < 4; ++m) out[nco + m] = s.det[0] * s.det[1] * J * s.dbf[m];
      return;
    }
    case WF_LCAO_1E_2C:
      for (int c = 0; c < 2; ++c) out[c] = lcao_chi(wf, c, cfg);
      return;
    case WF_LCAO_2E_1C:
    case WF_LCAO_2E_2C: {
      const int nc = lcao_nc(wf.kind);
      const R a0 = lcao_orbital(wf, 0, cfg).v, b1 = lcao_orbital(wf, 1, cfg + 3).v;
      for (int c = 0; c < nc; ++c) {
        out[c] = lcao_chi(wf, c, cfg) * b1;             // d / d C[0][c]
        out[nc + c] = a0 * lcao_chi(wf, c, cfg + 3);    // d / d C[1][c]
      }
      if (wf.geom[0] != 0.0) {
        const R a1 = lcao_orbital(wf, 0, cfg + 3).v, b0 = lcao_orbital(wf, 1, cfg).v;
        for (int c = 0; c < nc; ++c) {
          out[c] -= lcao_chi(wf, c, cfg + 3) * b0;
          out[nc + c] -= a1 * lcao_chi(wf, c, cfg);
        }
      }
      return;
    }
    case WF_CONSTANT:
      return;
  }
  throw std::runtime_error("wf_parameter_gradient: unknown kind");
}

// ---------------------------------------------------------------- operators (src/operator/src/operator.rs)
template <class R>
struct Ham {
  HamDesc d;
  double ionic_repulsion;
  explicit Ham(const HamDesc& desc) : d(desc), ionic_repulsion(0.0) {
    // IonicPotential::new, operator.rs:40-55
    double pot = 0.0;
    for (int i = 0; i < d.n_ions; ++i)
      for (int j = i + 1; j < d.n_ions; ++j) {
        double s[3];
        for (int k = 0; k < 3; ++k) s[k] = d.ion_pos[3 * j + k] - d.ion_pos[3 * i + k];
        pot += (double)(d.ion_charge[i] * d.ion_charge[j]) / norm_l2(s, 3);
      }
    ionic_repulsion = pot;
  }
};

// IonicPotential::value, operator.rs:25-36
template <class R>
R ionic_potential(const Ham<R>& h, const R* cfg, int ne) {
  R pot = R(0.0);
  for (int i = 0; i < h.d.n_ions; ++i)
    for (int j = 0; j < ne; ++j) {
      R s[3];
      for (int k = 0; k < 3; ++k) s[k] = cfg[3 * j + k] - R(h.d.ion_pos[3 * i + k]);
      pot -= R((double)h.d.ion_charge[i]) / norm_l2(s, 3);
    }
  return pot + R(h.ionic_repulsion);
}
// ElectronicPotential::value, operator.rs:80-90
template <class R>
R electronic_potential(const R* cfg, int ne) {
  R pot = R(0.0);
  for (int i = 0; i < ne; ++i)
    for (int j = i + 1; j < ne; ++j) {
      R s[3];
      for (int k = 0; k < 3; ++k) s[k] = cfg[3 * i + k] - cfg[3 * j + k];
      pot += R(1.0) / norm_l2(s, 3);
    }
  return pot;
}

// LocalOperator::act_on -> H psi (NOT divided by psi)
template <class R>
R ham_act_on(const Ham<R>& h, const Wf<R>& wf, const R* cfg) {
  const int ne = wf.ne;
  switch (h.d.kind) {
    case HAM_KINETIC:  // operator.rs:122-124
      return R(-0.5) * wf_laplacian(wf, cfg);
    case HAM_IONIC_POT:  // operator.rs:59-61
      return ionic_potential(h, cfg, ne) * wf_value(wf, cfg);
    case HAM_ELEC_POT:  // operator.rs:94-96
      return electronic_potential(cfg, ne) * wf_value(wf, cfg);
    case HAM_IONIC:  // operator.rs:146-148
      return R(-0.5) * wf_laplacian(wf, cfg) + ionic_potential(h, cfg, ne) * wf_value(wf, cfg);
    case HAM_ELECTRONIC:  // operator.rs:181-183
      return R(-0.5) * wf_laplacian(wf, cfg) + ionic_potential(h, cfg, ne) * wf_value(wf, cfg) +
             electronic_potential(cfg, ne) * wf_value(wf, cfg);
    case HAM_HARMONIC: {  // custom_operator.rs:52-60
      const R ke = R(-0.5) * wf_laplacian(wf, cfg);
      const R n = norm_l2(cfg, 3 * ne);
      const R w = R(h.d.frequency);
      const R pe = R(0.5) * (w * w) * (n * n) * wf_value(wf, cfg);
      return ke + pe;
    }
  }
  throw std::runtime_error("ham_act_on: unknown kind");
}

}  // namespace orc
