// Metropolis::move_state for one electron of the Slater-Jastrow kind (included from mole_sj.cuh).
//
// Semantics: src/metropolis/src/metrop.rs:60-96 (box) and :150-212 (diffusion) with the Frobenius norms over
// ALL electrons' drift.  With d = x' - x of the moved electron e, V_j the drift of electron j at the current
// point and V'_j at the trial point (V_j tau enters the norms):
//   -ln t_low  = [ |d - V_e tau|^2  + tau^2 sum_{j != e} |V_j|^2  ] / 2 tau = ( |d|^2 - 2 tau d.V_e ) / 2 tau + (tau/2) Q
//   -ln t_high = [ |-d - V'_e tau|^2 + tau^2 sum_{j != e} |V'_j|^2 ] / 2 tau = -ln t_low + (tau/2) dQ + d.(V_e + V'_e)
// where Q = sum_j |V_j|^2 and dQ = sum_j (|V'_j|^2 - |V_j|^2) = sum_j dV_j . (V_j + V'_j).  The acceptance
//   A = t_high psi'^2 / (t_low psi^2) = ratio^2 exp(2 df - (tau/2) dQ - d.(V_e + V'_e))
// therefore needs only DIFFERENCES of the drift, which is also how the drift is carried:
//   same spin,  j != e:  V'_j = V_j + d(grad f_j) - (v_j / ratio) H_j,   H_j = sum_k grad phi_k(r_j) Minv[k][e]
//   other spin:          V'_j = V_j + d(grad f_j)
//   j = e:               V'_e = grad f_e(x') + H_e(x') / ratio
// (Sherman-Morrison: Minv'[.][j] = Minv[.][j] - Minv[.][e] v_j / ratio, v_j = phi(x') . Minv[.][j]).  No norm is
// re-summed and neither grad ln D nor grad f is stored (the measurement re-forms both).  Q is carried only for the
// range guard below.
//
// The three independent transcendental chains of a move (orbital radial part and the two Jastrow pairs
// of each lane) go through the batched, stage-major math of mole_math.cuh and every lane sums the
// mailbox rows itself (four warp syncs per move: proposal, pair/orbital exchange, drift differences, commit).
#pragma once

// |x| < bound, decided on the high word by the integer ALU (NaN and Inf fail): bound_hi = high word of the bound
MOLE_D bool sj_abs_below(double x, int bound_hi) { return (__double2hiint(x) & 0x7fffffff) < bound_hi; }
// 2^-133 <= |x| < 2^133 (about 1e-40 .. 1e40), on the exponent field; 0, denormals, Inf and NaN fail
MOLE_D bool sj_mid_range(double x) { return (unsigned)(((__double2hiint(x) >> 20) & 0x7ff) - (1023 - 133)) < 266u; }

template <int METROP>
MOLE_D bool sj_move(const SjConst& c, SjLane& L, int el, const MoveDraw& d, double param, double sd, double inv2tau,
                    uint32_t compat) {
  const int spin = L.ph;
  const int n = sj_spin_n(c, spin);
  const bool isown = (L.gl == el);
  const int sid_e = spin * 5 + el;
  double* const mb = L.sm + SJ_OFF_MB;
  const double* const minv = L.sm + SJ_OFF_MINV + spin * 25;
  // this lane's column (cg) and the moved electron's column (ce) of the inverse
  double cg[5], ce[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) { cg[k] = isown ? 0.0 : minv[k * 5 + L.gl]; ce[k] = minv[k * 5 + el]; }   // owner: see the column update
  // ---- A: the owner proposes, everybody reads the trial point and the displacement
  if (isown && L.wr) {
    double dv[3];
#if MOLE_SJ_PARK_DRAW
    const double* const pk = L.sm + SJ_OFF_PARK + L.gl;            // this lane's draws (sj_sweep_moves parked them)
    const double dr[4] = {pk[40], pk[45], pk[50], pk[55]};
#else
    const double dr[4] = {d.a, d.b, d.c, d.u};
#endif
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double xi = dr[q];
      double xn;
      if (METROP == MOLE_METROP_BOX) {
        const double lo = -0.5 * param, scale = 0.5 * param - lo;
        xn = L.x[0][q] + (lo + scale * xi);                                      // metrop.rs:63-68
      } else {
        xn = (L.x[0][q] + L.V[0][q] * param) + sd * xi;                          // metrop.rs:155-160
      }
      dv[q] = xn - L.x[0][q];
      mb[MB_XN + q] = xn;
      mb[MB_XN + 4 + q] = dv[q];
    }
    mb[MB_XN + 3] = dr[3];
    if (METROP == MOLE_METROP_DIFFUSE) {
      const double dd = fma(dv[2], dv[2], fma(dv[1], dv[1], dv[0] * dv[0]));
      const double dV = fma(dv[2], L.V[0][2], fma(dv[1], L.V[0][1], dv[0] * L.V[0][0]));
      mb[MB_XN + 7] = fma(-2.0 * param, dV, dd);                                 // |d|^2 - 2 tau d.V_e
    }
  }
  sj_sync();
  double xn[3], dmv[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) { xn[q] = mb[MB_XN + q]; dmv[q] = mb[MB_XN + 4 + q]; }
  const double u_acc = mb[MB_XN + 3];

  // ---- B: three independent chains, batched: |x'| and the two pair distances -> rsqrt -> exp
  bool pv[2];
  int pid[2];
  double dn[2][3], r2v[3], rv[3], iv[3], ea[3], ev[3];
  r2v[0] = fma(xn[2], xn[2], fma(xn[1], xn[1], xn[0] * xn[0]));
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int b = (t ^ L.ph) * 5 + L.gl;
    pv[t] = L.val[t] && !(t == 0 && isown);
    pid[t] = max(sj_pidx(sid_e, b), 0);   // branch-free; masked entries read finite (zero-initialised) cache slots
#pragma unroll
    for (int q = 0; q < 3; ++q) dn[t][q] = xn[q] - L.x[t][q];
    r2v[1 + t] = pv[t] ? fma(dn[t][2], dn[t][2], fma(dn[t][1], dn[t][1], dn[t][0] * dn[t][0])) : 1.0;
  }
  m_sqrt_rsqrt_n<3>(r2v, rv, iv);
  ea[0] = -(L.gl == 0 ? c.z1 : (L.gl == 1 ? c.z2 : c.z3)) * rv[0];
  ea[1] = -c.kappa * rv[1];
  ea[2] = -c.kappa * rv[2];
  m_exp_n<3, true>(ea, ev);
  if (L.gl < 3 && L.wr) mb[MB_E + L.gl] = ev[0];                 // one orbital exponential per lane, shared below
  // pair functions (theory/jastrow.tex:23-31,45-48,68-71), the two pairs stage by stage; the Laplacian term is
  // formed at measurement time from the cached g/r and R
  double Rs[2], den[2], iden[2], pu[2], pgr[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) { Rs[t] = (1.0 - ev[1 + t]) * c.ikappa; den[t] = fma(c.b2, Rs[t], 1.0); }
  m_rcp_n<2>(den, iden);
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const double R2 = Rs[t] * Rs[t];
    pu[t] = fma(R2, fma(c.b4, Rs[t], c.b3), (c.b1 * Rs[t]) * iden[t]);
    const double du = fma(c.b1, iden[t] * iden[t], fma(3.0 * c.b4, R2, 2.0 * c.b3 * Rs[t]));
    pgr[t] = (ev[1 + t] * du) * iv[1 + t];
  }
  // d(grad f_b) of this lane's two electrons and this lane's share of df and of grad f_e(x')
  double dfl = 0.0, ge[3] = {0.0, 0.0, 0.0}, dgf[2][3];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const double u_old = L.sm[SJ_OFF_PC + pid[t]], gr_old = L.sm[SJ_OFF_PC + SJ_NPAIR + pid[t]];
    const double m = pv[t] ? 1.0 : 0.0;
    dfl = fma(m, pu[t] - u_old, dfl);
    const double gn_ = m * pgr[t], go_ = m * gr_old, gd_ = go_ - gn_;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      // grad_b f changes by the (b,e) term -(x_e - x_b) g/r: old term back in, new term out, with
      // x_e - x_b = dn - d at the old point: go (dn - d) - gn dn = (go - gn) dn - go d
      dgf[t][q] = fma(-go_, dmv[q], gd_ * dn[t][q]);
      ge[q] = fma(gn_, dn[t][q], ge[q]);
    }
  }
  if (L.wr) {                                                    // reduction inputs: df and grad_e f (from scratch)
    mb[MB_RIN + L.gl] = dfl;
    mb[MB_RIN + 5 + L.gl] = ge[0];
    mb[MB_RIN + 10 + L.gl] = ge[1];
    mb[MB_RIN + 15 + L.gl] = ge[2];
  }
  sj_sync();

  // ---- C: every lane sums the four rows (lane order 0..4), determinant ratio, Sherman-Morrison column
  double on[5];
  on[0] = rv[0]; on[1] = iv[0];
  on[2] = mb[MB_E]; on[3] = mb[MB_E + 1]; on[4] = mb[MB_E + 2];
  double red[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double* r = mb + MB_RIN + 5 * j;
    red[j] = ((r[0] + r[1]) + (r[2] + r[3])) + r[4];             // fixed order, three dependent adds instead of four
  }
  const double df = red[0];
  double phin[5];
  sj_phi(xn, on, n, phin);
  // phi'.Minv columns as two partial dot products each (dependent depth 4 instead of 6)
  const double v = fma(phin[2], cg[2], fma(phin[1], cg[1], phin[0] * cg[0])) + fma(phin[4], cg[4], phin[3] * cg[3]);
  const double ratio = fma(phin[2], ce[2], fma(phin[1], ce[1], phin[0] * ce[0])) + fma(phin[4], ce[4], phin[3] * ce[3]);
  const double inv_ratio = m_rcp(ratio);
  // Sherman-Morrison column as ONE fma for every lane: the moved electron's own column is ce / ratio = 0 - ce (-1/ratio)
  // (its cg was loaded as zeros), the others are cg - ce (v / ratio)
  const double vr = isown ? -inv_ratio : v * inv_ratio;
  double mt[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) mt[k] = fma(-ce[k], vr, cg[k]);
  // H = sum_k grad phi_k Minv[k][e] at this lane's slot-0 electron (the owner: at the trial point); the trial drift
  // of slot 0 is base - vr H with base = V + d(grad f) (owner: the freshly summed grad f_e), of slot 1 V + d(grad f)
  double H[3], Vt[2][3];
  bool acc;
  const bool node = !(ratio > 0.0);                              // signum(psi') != signum(psi) or NaN, :178-180
  if (METROP == MOLE_METROP_DIFFUSE) {
    sj_gradlnD(c, isown ? xn : L.x[0], isown ? on : L.orb[0], ce, H);
    // absent electrons keep their (finite, arbitrary) drift: their differences are masked to zero
    double dq[3], dt[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double base = isown ? red[1 + q] : L.V[0][q] + dgf[0][q];
      Vt[0][q] = L.val[0] ? fma(-vr, H[q], base) : L.V[0][q];
      Vt[1][q] = L.V[1][q] + dgf[1][q];                          // dgf of an absent electron is an exact zero
      const double d0 = Vt[0][q] - L.V[0][q], s0 = Vt[0][q] + L.V[0][q], s1 = Vt[1][q] + L.V[1][q];
      dq[q] = fma(d0, s0, dgf[1][q] * s1);
      dt[q] = (isown ? dmv[q] : 0.0) * s0;
    }
    if (L.wr) {
      mb[MB_RIN + 20 + L.gl] = (dq[0] + dq[1]) + dq[2];
      if (isown) mb[MB_RIN + 25] = (dt[0] + dt[1]) + dt[2];
    }
    sj_sync();
    // ---- D: A = ratio^2 exp(2 df - (tau/2) dQ - d.(V_e + V'_e)): ONE exponential and no division wherever none of
    // the reference's intermediate products can leave the normal range (t_high, t_low >= e^-200, psi^2 and psi'^2
    // within 1e-80 .. 1e80): the same real number as (t_high psi'^2) / (t_low psi^2), relative difference ~1e-16.
    // Everywhere else (walkers next to a node: t denormal or 0, t psi^2 underflowing, 0/0 = NaN) the reference's own
    // sequence of operations on the absolute psi is evaluated, with Q re-summed exactly first; that branch is taken
    // by the whole warp when any of its walkers needs it (it holds a group reduction).
    const double* rq = mb + MB_RIN + 20;
    const double dQ = ((rq[0] + rq[1]) + (rq[2] + rq[3])) + rq[4];
    const double dse = mb[MB_RIN + 25], t1 = mb[MB_XN + 7];
    const double hq = fma(0.5 * param, dQ, dse);                   // (-ln t_high) - (-ln t_low)
    double sls = fma(t1, inv2tau, 0.5 * param * L.Q);              // -ln t_low
    double shs = sls + hq;                                         // -ln t_high
    const double r2 = ratio * ratio;
    const bool fast = sj_abs_below(shs, 0x40690000) && sj_abs_below(sls, 0x40690000) && sj_abs_below(df, 0x40490000) &&
                      sj_abs_below(L.fj, 0x40490000) && sj_mid_range(L.psi) && sj_mid_range(r2);
    double A = sj_clamp_acceptance(r2 * m_exp(fma(2.0, df, -hq)), compat);
    if (__any_sync(SJ_FULL, L.wr && !fast)) {
      const double Qx = sj_q_exact(L);
      L.Q = Qx;
      if (!fast) {
        sls = fma(t1, inv2tau, 0.5 * param * Qx);
        shs = sls + hq;
        const double targ[4] = {df, -shs, -sls, L.fj};
        double tv[4];
        m_exp_n<4>(targ, tv);
        const double psi_old = L.psi * tv[3], psi_new = psi_old * (ratio * tv[0]);
        A = sj_clamp_acceptance((tv[1] * (psi_new * psi_new)) / (tv[2] * (psi_old * psi_old)), compat);   // :195
      }
    }
    acc = !node && (A > u_acc);
    if (acc) L.Q += dQ;
  } else {
    const double qb = ratio * m_exp(df);
    acc = sj_clamp_acceptance(qb * qb, compat) > u_acc;            // metrop.rs:80
    if (acc) {                                                     // the drift is only needed by later samples
      sj_gradlnD(c, isown ? xn : L.x[0], isown ? on : L.orb[0], ce, H);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double base = isown ? red[1 + q] : L.V[0][q] + dgf[0][q];
        Vt[0][q] = L.val[0] ? fma(-vr, H[q], base) : L.V[0][q];
        Vt[1][q] = L.V[1][q] + dgf[1][q];
      }
    }
  }
  if (acc) {
    if (isown) {
#pragma unroll
      for (int qq = 0; qq < 3; ++qq) {
        L.x[0][qq] = xn[qq];
#if MOLE_SJ_POS_IN_MOVE
        if (L.wr) L.sm[SJ_OFF_POS + 3 * sid_e + qq] = xn[qq];        // the shared copy sj_gradf reads
#endif
      }
#pragma unroll
      for (int qq = 0; qq < 5; ++qq) L.orb[0][qq] = on[qq];
    }
#pragma unroll
    for (int qq = 0; qq < 3; ++qq) { L.V[0][qq] = Vt[0][qq]; L.V[1][qq] = Vt[1][qq]; }
    if (L.wr) {                                                    // (a phantom group past W owns its slot too; testing
      double* mw = L.sm + SJ_OFF_MINV + spin * 25;
#pragma unroll
      for (int k = 0; k < 5; ++k) mw[k * 5 + L.gl] = mt[k];
    }
    L.psi *= ratio;                                                // psi = L.psi exp(L.fj), folded by sj_fold_psi
    L.fj += df;
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (pv[t] && L.wr) {                                           // w < W here would be re-derived for every move)
        double* pc = L.sm + SJ_OFF_PC + pid[t];
        pc[0] = pu[t]; pc[SJ_NPAIR] = pgr[t]; pc[2 * SJ_NPAIR] = iv[1 + t]; pc[3 * SJ_NPAIR] = Rs[t];
      }
  }
  sj_sync();
  return acc;
}
