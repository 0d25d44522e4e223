// Metropolis::move_state for one electron of the Slater-Jastrow kind (included from mole_sj.cuh).
//
// The three independent transcendental chains of a move (orbital radial part and the two Jastrow pairs
// of each lane) go through the batched, stage-major math of mole_math.cuh and every lane sums the
// mailbox rows itself (four warp syncs per move: proposal, pair/orbital exchange, drift norms, commit).
// The kernel is bound by the FP64 pipe plus exposed dependency latency: what counts
// here is the number of FP64 instructions and the length of dependent chains, not integer work or syncs.
// Semantics: src/metropolis/src/metrop.rs:60-96 (box) and :150-212 (diffusion), Frobenius norms over
// ALL electrons' drift; the accept test is documented at phase D.
#pragma once

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
  // grad ln D of the own electrons is carried in L.G (rebuilt by sj_refresh, updated on accepted moves)
  const double* const Gown = L.G[0];
  // ---- A: the owner proposes, everybody reads the trial point
  if (isown && L.wr) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double xi = q == 0 ? d.a : (q == 1 ? d.b : d.c);
      if (METROP == MOLE_METROP_BOX) {
        const double lo = -0.5 * param, scale = 0.5 * param - lo;
        mb[MB_XN + q] = L.x[0][q] + (lo + scale * xi);                              // metrop.rs:63-68
      } else {
        mb[MB_XN + q] = (L.x[0][q] + (Gown[q] + L.gf[0][q]) * param) + sd * xi;     // metrop.rs:155-160
      }
      mb[MB_XN + 4 + q] = L.x[0][q];
    }
    mb[MB_XN + 3] = d.u;
  }
  sj_sync();
  double xn[3], xo[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) { xn[q] = mb[MB_XN + q]; xo[q] = mb[MB_XN + 4 + q]; }
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
  // pair functions (theory/jastrow.tex:23-31,45-48,68-71,82-97), the two pairs stage by stage
  double Rs[2], den[2], iden[2], pu[2], pgr[2], plt[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) { Rs[t] = (1.0 - ev[1 + t]) * c.ikappa; den[t] = fma(c.b2, Rs[t], 1.0); }
  m_rcp_n<2>(den, iden);
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const double R2 = Rs[t] * Rs[t], id2 = iden[t] * iden[t], E = ev[1 + t];
    pu[t] = fma(R2, fma(c.b4, Rs[t], c.b3), (c.b1 * Rs[t]) * iden[t]);
    const double du = fma(c.b1, id2, fma(3.0 * c.b4, R2, 2.0 * c.b3 * Rs[t]));
    const double d2u = fma(-2.0 * c.b1 * c.b2, id2 * iden[t], fma(6.0 * c.b4, Rs[t], 2.0 * c.b3));
    const double g = E * du;
    pgr[t] = g * iv[1 + t];
    plt[t] = fma(2.0, pgr[t], fma(E * E, d2u, -c.kappa * g));   // div(rhat g) = 2 g/r + dg/dr
  }
  double dfl = 0.0, ge[3] = {0.0, 0.0, 0.0}, gft[2][3];
  const double dmv[3] = {xn[0] - xo[0], xn[1] - xo[1], xn[2] - xo[2]};
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const double u_old = L.sm[SJ_OFF_PC + pid[t]], gr_old = L.sm[SJ_OFF_PC + SJ_NPAIR + pid[t]];
    const double m = pv[t] ? 1.0 : 0.0;
    dfl = fma(m, pu[t] - u_old, dfl);
    const double gn_ = m * pgr[t], go_ = m * gr_old, gd_ = go_ - gn_;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      // grad_b f changes by the (b,e) term -(x_e - x_b) g/r: old term back in, new term out, with
      // x_e - x_b = dn - (x' - x) at the old point: gf + go (dn - dmove) - gn dn = gf + (go - gn) dn - go dmove
      gft[t][q] = fma(-go_, dmv[q], fma(gd_, dn[t][q], L.gf[t][q]));
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
  if (isown) { gft[0][0] = red[1]; gft[0][1] = red[2]; gft[0][2] = red[3]; }
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
  double Gt[3];
  bool acc;
  if (METROP == MOLE_METROP_DIFFUSE) {
    // Frobenius norms over ALL electrons' drift (metrop.rs:182-193)
    sj_gradlnD(c, isown ? xn : L.x[0], isown ? on : L.orb[0], mt, Gt);
    const double* const G1 = L.G[1];
    double ph[3], pl[3];                                           // per-component partial sums: short chains
    // absent electrons (slot validity) drop out through a zero time step: their drift is finite, their dx is 0
    const double tau0 = L.val[0] ? param : 0.0, tau1 = L.val[1] ? param : 0.0;
#pragma unroll
    for (int qq = 0; qq < 3; ++qq) {
      const double dx = isown ? xo[qq] - xn[qq] : 0.0;
      const double a0 = dx - (Gt[qq] + gft[0][qq]) * tau0, b0 = -dx - (Gown[qq] + L.gf[0][qq]) * tau0;
      const double a1 = (G1[qq] + gft[1][qq]) * tau1, b1 = (G1[qq] + L.gf[1][qq]) * tau1;
      ph[qq] = fma(a0, a0, a1 * a1);
      pl[qq] = fma(b0, b0, b1 * b1);
    }
    const double sh = (ph[0] + ph[1]) + ph[2], sl = (pl[0] + pl[1]) + pl[2];
    if (L.wr) { mb[MB_RIN + 20 + L.gl] = sh; mb[MB_RIN + 25 + L.gl] = sl; }
    sj_sync();
    // ---- D: acceptance A = (t_high psi'^2) / (t_low psi^2) (metrop.rs:182-195), t = exp(-s/2tau),
    // psi' = psi ratio exp(df).  Where none of the reference's intermediate products can leave the normal
    // range (bounds below: t >= e^-200, psi^2 and psi'^2 within 1e-207 .. 1e207) the three exponentials and the
    // division are ONE exponential, A = ratio^2 exp(2 df + (s_low - s_high)/2tau): the same real number,
    // relative difference ~1e-16.  Everywhere else (walkers next to a node: t denormal or 0, t psi^2
    // underflowing, 0/0 = NaN) the reference's own sequence of operations on the absolute psi is evaluated.
    const double* rh = mb + MB_RIN + 20;
    const double* rl = mb + MB_RIN + 25;
    const double shs = (((rh[0] + rh[1]) + (rh[2] + rh[3])) + rh[4]) * inv2tau;   // -ln t_high
    const double sls = (((rl[0] + rl[1]) + (rl[2] + rl[3])) + rl[4]) * inv2tau;   // -ln t_low
    const bool node = !(ratio > 0.0);                              // signum(psi') != signum(psi) or NaN, :178-180
    const double r2 = ratio * ratio, ap = fabs(L.psi);
    double A;
    if (fmax(shs, sls) < 200.0 && fabs(df) < 50.0 && fabs(L.fj) < 50.0 && ap > 1e-40 && ap < 1e40 && r2 > 1e-40 && r2 < 1e40) {
      A = sj_clamp_acceptance(r2 * m_exp(fma(2.0, df, sls - shs)), compat);
    } else {
      const double targ[4] = {df, -shs, -sls, L.fj};
      double tv[4];
      m_exp_n<4>(targ, tv);
      const double psi_old = L.psi * tv[3], psi_new = psi_old * (ratio * tv[0]);
      A = sj_clamp_acceptance((tv[1] * (psi_new * psi_new)) / (tv[2] * (psi_old * psi_old)), compat);   // :195
    }
    acc = !node && (A > u_acc);
  } else {
    const double qb = ratio * m_exp(df);
    acc = sj_clamp_acceptance(qb * qb, compat) > u_acc;            // metrop.rs:80
  }
  if (acc) {
    if (isown) {
#pragma unroll
      for (int qq = 0; qq < 3; ++qq) L.x[0][qq] = xn[qq];
#pragma unroll
      for (int qq = 0; qq < 5; ++qq) L.orb[0][qq] = on[qq];
    }
    if (METROP == MOLE_METROP_DIFFUSE) {
#pragma unroll
      for (int qq = 0; qq < 3; ++qq) L.G[0][qq] = Gt[qq];
    } else {
      sj_gradlnD(c, L.x[0], L.orb[0], mt, L.G[0]);
    }
    if (L.wr) {                                                    // (a phantom group past W owns its slot too; testing
      double* mw = L.sm + SJ_OFF_MINV + spin * 25;
#pragma unroll
      for (int k = 0; k < 5; ++k) mw[k * 5 + L.gl] = mt[k];
    }
#pragma unroll
    for (int qq = 0; qq < 3; ++qq) { L.gf[0][qq] = gft[0][qq]; L.gf[1][qq] = gft[1][qq]; }
    L.psi *= ratio;                                                // psi = L.psi exp(L.fj), folded by sj_fold_psi
    L.fj += df;
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (pv[t] && L.wr) {                                           // w < W here would be re-derived for every move)
        double* pc = L.sm + SJ_OFF_PC + pid[t];
        pc[0] = pu[t]; pc[SJ_NPAIR] = pgr[t]; pc[2 * SJ_NPAIR] = plt[t]; pc[3 * SJ_NPAIR] = iv[1 + t]; pc[4 * SJ_NPAIR] = Rs[t];
      }
  }
  sj_sync();
  return acc;
}
