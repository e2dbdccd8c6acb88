// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_solver.hpp).  Parity pinned to the reference's source text (tests/test_oracle_vs_reference_source.py).
// Operation order inside every expression follows the Fortran source left to right; build
// with -ffp-contract=off so no FMA is formed.
#include "oracle_solver.hpp"

namespace orc {

static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// vec_weight, mod_util.f90:729-743 : weight ON r2 (the neighbour)
static inline void vec_weight(double& wt, const double* r0, const double* r1, const double* r2) {
  double ra[3] = {r1[0] - r0[0], r1[1] - r0[1], r1[2] - r0[2]};
  double rb[3] = {r2[0] - r0[0], r2[1] - r0[1], r2[2] - r0[2]};
  double la = std::sqrt(ra[0] * ra[0] + ra[1] * ra[1] + ra[2] * ra[2]);
  double lb = std::sqrt(rb[0] * rb[0] + rb[1] * rb[1] + rb[2] * rb[2]);
  if (la + lb > 0.0) wt = la / (la + lb); else wt = 0.5;
}

// ---- matinv3 + calc_grad, mod_solver.f90:8-81 ----------------------------------------------
static void matinv3(const double A[3][3], double B[3][3]) {
  double det = (A[0][0] * A[1][1] * A[2][2] - A[0][0] * A[1][2] * A[2][1]
              - A[0][1] * A[1][0] * A[2][2] + A[0][1] * A[1][2] * A[2][0]
              + A[0][2] * A[1][0] * A[2][1] - A[0][2] * A[1][1] * A[2][0]);
  if (std::fabs(det) > 2.2250738585072014e-308) {
    double detinv = 1.0 / det;
    B[0][0] = +detinv * (A[1][1] * A[2][2] - A[1][2] * A[2][1]);
    B[1][0] = -detinv * (A[1][0] * A[2][2] - A[1][2] * A[2][0]);
    B[2][0] = +detinv * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    B[0][1] = -detinv * (A[0][1] * A[2][2] - A[0][2] * A[2][1]);
    B[1][1] = +detinv * (A[0][0] * A[2][2] - A[0][2] * A[2][0]);
    B[2][1] = -detinv * (A[0][0] * A[2][1] - A[0][1] * A[2][0]);
    B[0][2] = +detinv * (A[0][1] * A[1][2] - A[0][2] * A[1][1]);
    B[1][2] = -detinv * (A[0][0] * A[1][2] - A[0][2] * A[1][0]);
    B[2][2] = +detinv * (A[0][0] * A[1][1] - A[0][1] * A[1][0]);
  } else {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[i][j] = 0.0;
    det = A[0][0] + A[1][1] + A[2][2];
    double detinv = 1.0 / det;
    B[0][0] = detinv; B[1][1] = detinv; B[2][2] = detinv;
  }
}

void calc_grad(const double* phi, double* grad, const double* xc, const double* yc, const double* zc,
               const int* ef2nb_idx, const int* ef2nb1, int ne) {
  for (int e = 1; e <= ne; ++e) {
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, g[3] = {0, 0, 0};
    for (int idx = ef2nb_idx[e - 1]; idx <= ef2nb_idx[e] - 1; ++idx) {
      int enb, lfnb;
      get_idx(ef2nb1[idx - 1], enb, lfnb);
      double dr[3] = {xc[enb - 1] - xc[e - 1], yc[enb - 1] - yc[e - 1], zc[enb - 1] - zc[e - 1]};
      double wt = 1.0 / (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
      double dphi = phi[enb - 1] - phi[e - 1];
      for (int i = 0; i < 3; ++i) g[i] = g[i] + wt * dphi * dr[i];
      A[0][0] = A[0][0] + wt * dr[0] * dr[0];
      A[0][1] = A[0][1] + wt * dr[0] * dr[1];
      A[0][2] = A[0][2] + wt * dr[0] * dr[2];
      A[1][1] = A[1][1] + wt * dr[1] * dr[1];
      A[1][2] = A[1][2] + wt * dr[1] * dr[2];
      A[2][2] = A[2][2] + wt * dr[2] * dr[2];
    }
    A[1][0] = A[0][1]; A[2][0] = A[0][2]; A[2][1] = A[1][2];
    double Ai[3][3];
    matinv3(A, Ai);
    for (int i = 0; i < 3; ++i) {
      double c = 0.0;
      for (int j = 0; j < 3; ++j) c = c + Ai[i][j] * g[j];
      grad[3 * (long)(e - 1) + i] = c;
    }
  }
}

// ---- solve_gs / smoother_gs / calc_residual, mod_solver.f90:191-327 ------------------------
static inline double row_sum(int e, const double* phi, const double* anb, const double* b,
                             const int* ef2nb_idx, const int* ef2nb1) {
  double sumnb = b[e - 1];
  for (int idx = ef2nb_idx[e - 1]; idx <= ef2nb_idx[e] - 1; ++idx) {
    int enb = (int)((uint32_t)ef2nb1[idx - 1] >> num_face_bits);
    sumnb = sumnb + anb[idx - 1] * phi[enb - 1];
  }
  return sumnb;
}

SolveStat solve_gs(bool is_pc, double* phi, const double* ap, const double* anb, const double* b,
                   const int* ef2nb_idx, const int* ef2nb1, int ne, int nit) {
  double sor = 1.0;
  if (is_pc) sor = 1.02;
  SolveStat st;
  double res_i = 0.0;
  for (int e = 1; e <= ne; ++e) {
    double r = row_sum(e, phi, anb, b, ef2nb_idx, ef2nb1) - ap[e - 1] * phi[e - 1];
    res_i = res_i + r * r;
  }
  res_i = std::sqrt(res_i / ne);
  int it = 0;
  double res_f = res_i, res_target = res_i / 10.0, res_max = 0.0;  // res_max is undefined in the reference when it==0
  while (it < nit && res_f > res_target) {
    it = it + 1;
    for (int e = 1; e <= ne; ++e) {
      double sumnb = row_sum(e, phi, anb, b, ef2nb_idx, ef2nb1);
      phi[e - 1] = (sumnb + (sor - 1.0) * ap[e - 1] * phi[e - 1]) / ap[e - 1] / sor;
    }
    for (int e = ne; e >= 1; --e) {
      double sumnb = row_sum(e, phi, anb, b, ef2nb_idx, ef2nb1);
      phi[e - 1] = (sumnb + (sor - 1.0) * ap[e - 1] * phi[e - 1]) / ap[e - 1] / sor;
    }
    res_max = 0.0;
    res_f = 0.0;
    for (int e = 1; e <= ne; ++e) {
      double r = std::fabs(row_sum(e, phi, anb, b, ef2nb_idx, ef2nb1) - ap[e - 1] * phi[e - 1]);
      res_max = std::max(r, res_max);
      res_f = res_f + r * r;
    }
    res_f = std::sqrt(res_f / ne);
  }
  st.it = it; st.res_i = res_i; st.res_f = res_f; st.res_max = res_max;
  return st;
}

void smoother_gs(bool is_pc, double* phi, const double* ap, const double* anb, const double* b,
                 const int* ef2nb_idx, const int* ef2nb1, int ne, int nit) {
  double sor = 1.0;
  if (is_pc) sor = 1.02;
  for (int it = 1; it <= nit; ++it) {
    for (int e = 1; e <= ne; ++e) {
      double sumnb = row_sum(e, phi, anb, b, ef2nb_idx, ef2nb1);
      phi[e - 1] = (sumnb + (sor - 1.0) * ap[e - 1] * phi[e - 1]) / ap[e - 1] / sor;
    }
    for (int e = ne; e >= 1; --e) {
      double sumnb = row_sum(e, phi, anb, b, ef2nb_idx, ef2nb1);
      phi[e - 1] = (sumnb + (sor - 1.0) * ap[e - 1] * phi[e - 1]) / ap[e - 1] / sor;
    }
  }
}

void calc_residual(const double* phi, const double* ap, const double* anb, const double* b,
                   const int* ef2nb_idx, const int* ef2nb1, int ne, double& res, double& res_max) {
  res = 0.0;
  res_max = 0.0;
  for (int e = 1; e <= ne; ++e) {
    double r = row_sum(e, phi, anb, b, ef2nb_idx, ef2nb1) - ap[e - 1] * phi[e - 1];
    res_max = std::max(r, res_max);  // signed r: reference quirk, :248
    res = res + r * r;
  }
  res = std::sqrt(res / ne);
}

// ---- mod_subdomains.f90 -------------------------------------------------------------------
void construct_subdomains(Case& c) {  // :18-160
  Mesh& m = c.m;
  const int P = c.n_subdomains;
  c.subdomain.assign(P, Subdomain());
  c.intf.assign((size_t)P * P, Intf());
  std::vector<std::vector<int>> intf_tmp((size_t)P * P);
  std::vector<int> ef2nb_tmp((size_t)(2 * m.nf - m.nbf) + 1, 0);
  auto S = [&](int k) -> Subdomain& { return c.subdomain[k - 1]; };
  for (int k = 1; k <= P; ++k) {
    S(k).id = k;
    S(k).ne = m.g2gf_idx(k + 1) - m.g2gf_idx(k);
    int ne = S(k).ne;
    S(k).ap.alloc(ne, 0.0); S(k).b.alloc(ne, 0.0);
    int nf = 0, nbf = 0;
    for (int n = m.g2gf_idx(k); n <= m.g2gf_idx(k + 1) - 1; ++n) {
      int gf = m.g2gf_p(n);
      for (int idx = m.ef2nb_idx(gf); idx <= m.ef2nb_idx(gf + 1) - 1; ++idx) {
        int gfnb, lfnb;
        get_idx(m.ef2nb1(idx), gfnb, lfnb);
        int cnb = 0;
        if (lfnb != 0) cnb = m.gf2g(gfnb);
        if (cnb == 0) { c.I(k, k).ncs += 1; nbf += 1; }
        else if (k != cnb) { c.I(k, cnb).ncs += 1; nbf += 1; }
        nf += 1;
      }
    }
    nf = (nf + nbf) / 2;
    S(k).nbf = nbf; S(k).nf = nf;
    S(k).anb.alloc(2 * nf - nbf, 0.0);
    S(k).phic.alloc(ne + nbf, 0.0);
    S(k).ef2nb.alloc(2 * nf - nbf, 0);
    S(k).ef2nb_idx.alloc(ne + 1, 0);
    for (int cnb = 1; cnb <= P; ++cnb) {
      int ncs = c.I(k, cnb).ncs;
      c.I(k, cnb).c1 = k; c.I(k, cnb).c2 = cnb;
      c.I(k, cnb).index1.assign(ncs, 0);
      if (k != cnb) intf_tmp[(size_t)(k - 1) * P + (cnb - 1)].assign(ncs, 0);
    }
  }
  for (int k = 1; k <= P; ++k) for (int cnb = 1; cnb <= P; ++cnb) c.I(k, cnb).ncs = 0;
  for (int k = 1; k <= P; ++k) {  // :88-114
    int nbf = 0, ne = S(k).ne;
    S(k).ef2nb_idx(1) = 1;
    for (int n = m.g2gf_idx(k); n <= m.g2gf_idx(k + 1) - 1; ++n) {
      int g = n - m.g2gf_idx(k) + 1, gf = m.g2gf_p(n);
      int nl = m.ef2nb_idx(gf + 1) - m.ef2nb_idx(gf);
      S(k).ef2nb_idx(g + 1) = S(k).ef2nb_idx(g) + nl;
      for (int idx = m.ef2nb_idx(gf); idx <= m.ef2nb_idx(gf + 1) - 1; ++idx) {
        int lf = idx - m.ef2nb_idx(gf) + 1, gfnb, lfnb;
        get_idx(m.ef2nb1(idx), gfnb, lfnb);
        int cnb = 0;
        if (lfnb != 0) cnb = m.gf2g(gfnb);
        if (k == cnb) {
          int idx2 = m.ef2nb_idx(gfnb) + lfnb - 1;
          ef2nb_tmp[idx2] = index_t(g, lf);
        } else {
          nbf += 1;
          ef2nb_tmp[idx] = index_t(ne + nbf, 0);
        }
      }
    }
  }
  for (int k = 1; k <= P; ++k)  // :116-147
    for (int n = m.g2gf_idx(k); n <= m.g2gf_idx(k + 1) - 1; ++n) {
      int g = n - m.g2gf_idx(k) + 1, gf = m.g2gf_p(n);
      for (int idx = m.ef2nb_idx(gf); idx <= m.ef2nb_idx(gf + 1) - 1; ++idx) {
        int lf = idx - m.ef2nb_idx(gf) + 1, gfnb, lfnb;
        get_idx(m.ef2nb1(idx), gfnb, lfnb);
        int cnb = 0;
        if (lfnb != 0) cnb = m.gf2g(gfnb);
        int idx2 = S(k).ef2nb_idx(g) + lf - 1;
        S(k).ef2nb(idx2) = ef2nb_tmp[idx];
        if (cnb == 0) {
          Intf& I = c.I(k, k);
          I.ncs += 1;
          I.index1[I.ncs - 1] = index_t(g, lf);
        } else if (k != cnb) {
          Intf& I = c.I(k, cnb);
          I.ncs += 1;
          int cs = I.ncs;
          I.index1[cs - 1] = index_t(g, lf);
          std::vector<int>& T = intf_tmp[(size_t)(k - 1) * P + (cnb - 1)];
          if (k < cnb) T[cs - 1] = index_t(gfnb, lfnb);
          else T[cs - 1] = index_t(gf, lf);
        }
      }
    }
  for (int k = 1; k <= P; ++k)  // :149-156 align index1/index2 on the shared key
    for (int cnb = 1; cnb <= P; ++cnb) {
      if (k == cnb) continue;
      Intf& I = c.I(k, cnb);
      std::vector<int>& T = intf_tmp[(size_t)(k - 1) * P + (cnb - 1)];
      if (I.ncs > 0) qsort_key(T.data(), I.index1.data(), 1, I.ncs);
    }
}

static void assemble_coef(Case& c, const double* ap, const double* anb, const double* b, const double* phi) {  // :162-189
  Mesh& m = c.m;
  for (int k = 1; k <= c.n_subdomains; ++k) {
    Subdomain& s = c.subdomain[k - 1];
    for (int e = 1; e <= s.ne; ++e) {
      int g = m.g2gf_p(m.g2gf_idx(k) + e - 1);
      s.ap(e) = ap[g - 1];
      s.b(e) = b[g - 1];
      s.phic(e) = phi[g - 1];
      for (int idx = s.ef2nb_idx(e); idx <= s.ef2nb_idx(e + 1) - 1; ++idx) {
        int enb, lfnb;
        get_idx(s.ef2nb(idx), enb, lfnb);
        int lf = idx - s.ef2nb_idx(e) + 1;
        int idx2 = m.ef2nb_idx(g) + lf - 1;
        s.anb(idx) = anb[idx2 - 1];
        if (lfnb == 0) {
          int gnb, t;
          get_idx(m.ef2nb1(idx2), gnb, t);
          s.phic(enb) = phi[gnb - 1];
        }
      }
    }
  }
}

static void update_halos(Case& c, int i, int j) {  // :191-212, intf(i,j), i<j
  Intf& I1 = c.I(i, j);
  Intf& I2 = c.I(j, i);
  Subdomain& s1 = c.subdomain[i - 1];
  Subdomain& s2 = c.subdomain[j - 1];
  for (int cs = 1; cs <= I1.ncs; ++cs) {
    int g1, lf1, g2, lf2, gnb1, gnb2, tmp;
    get_idx(I1.index1[cs - 1], g1, lf1);
    get_idx(s1.ef2nb(s1.ef2nb_idx(g1) + lf1 - 1), gnb1, tmp);
    get_idx(I2.index1[cs - 1], g2, lf2);
    get_idx(s2.ef2nb(s2.ef2nb_idx(g2) + lf2 - 1), gnb2, tmp);
    s1.phic(gnb1) = s2.phic(g2);
    s2.phic(gnb2) = s1.phic(g1);
  }
}

static void update_phi(Case& c, double* phi) {  // :214-229
  Mesh& m = c.m;
  for (int k = 1; k <= c.n_subdomains; ++k) {
    Subdomain& s = c.subdomain[k - 1];
    for (int e = 1; e <= s.ne; ++e) phi[m.g2gf_p(m.g2gf_idx(k) + e - 1) - 1] = s.phic(e);
  }
}

// ---- multi_subdomain_solver / solve, mod_solver.f90:124-189,329-344 ------------------------
SolveStat multi_subdomain_solver(Case& c, bool is_pc, const double* ap, const double* anb,
                                 const double* b, double* phi, int nit) {
  const int P = c.n_subdomains;
  assemble_coef(c, ap, anb, b, phi);
  double res_i_tot = 0.0, res_f_tot = 0.0, res_max_tot = 0.0, res, res_max;
  for (int k = 1; k <= P; ++k) {
    Subdomain& s = c.subdomain[k - 1];
    calc_residual(s.phic.data(), s.ap.data(), s.anb.data(), s.b.data(), s.ef2nb_idx.data(), s.ef2nb.data(), s.ne, res, res_max);
    res_i_tot = res_i_tot + res * res;
    res_max_tot = std::max(res_max_tot, res_max);
  }
  res_i_tot = std::sqrt(res_i_tot / P);
  int it = 0;
  res_f_tot = res_i_tot;
  double res_target = res_i_tot / 10.0;
  while (it < nit && res_f_tot > res_target) {
    for (int k = 1; k <= P; ++k) {
      Subdomain& s = c.subdomain[k - 1];
      smoother_gs(is_pc, s.phic.data(), s.ap.data(), s.anb.data(), s.b.data(), s.ef2nb_idx.data(), s.ef2nb.data(), s.ne, 2);
    }
    it = it + 2;
    for (int i = 1; i <= P; ++i)
      for (int j = i + 1; j <= P; ++j) update_halos(c, i, j);
    if (it % 10 == 0) {
      for (int k = 1; k <= P; ++k) {
        Subdomain& s = c.subdomain[k - 1];
        calc_residual(s.phic.data(), s.ap.data(), s.anb.data(), s.b.data(), s.ef2nb_idx.data(), s.ef2nb.data(), s.ne, res, res_max);
        res_f_tot = res_f_tot + res * res;  // NOT reset first: reference quirk, :175
        res_max_tot = std::max(res_max_tot, res_max);
      }
      res_f_tot = std::sqrt(res_f_tot / P);
    }
  }
  update_phi(c, phi);
  SolveStat st;
  st.it = it; st.res_i = res_i_tot; st.res_f = res_f_tot; st.res_max = res_max_tot;
  return st;
}

SolveStat solve(Case& c, bool is_pc, const double* ap, const double* anb, const double* b, double* phi, int nit) {
  if (c.n_subdomains == 1)
    return solve_gs(is_pc, phi, ap, anb, b, c.m.ef2nb_idx.data(), c.m.ef2nb1.data(), c.m.ne, nit);
  return multi_subdomain_solver(c, is_pc, ap, anb, b, phi, nit);
}

// ---- mod_uvwp.f90 -------------------------------------------------------------------------
void calc_coef_uvw(Case& c) {  // :161-286
  Mesh& g = c.m;
  const double dt = c.dt;
  for (int e = 1; e <= g.ne; ++e) {
    double rp[3] = {g.xc(e), g.yc(e), g.zc(e)};
    c.ap(e) = 0.0;
    double sumf = 0.0, sumss[3] = {0, 0, 0}, sumdefc[3] = {0, 0, 0};
    for (int idx = g.ef2nb_idx(e); idx <= g.ef2nb_idx(e + 1) - 1; ++idx) {
      c.anb(idx) = 0.0;
      int enb, lfnb;
      get_idx(g.ef2nb1(idx), enb, lfnb);
      double f = 0.0, fnb = 0.0, d = 0.0;
      if (lfnb > 0) {
        int fg = g.ef2nb2(idx);
        int fg_sgn = sgn(fg);
        fg = std::abs(fg);
        long i = 3 * (long)fg - 2;
        double area = std::sqrt(g.aip(i) * g.aip(i) + g.aip(i + 1) * g.aip(i + 1) + g.aip(i + 2) * g.aip(i + 2));
        double norm[3] = {fg_sgn * g.aip(i) / area, fg_sgn * g.aip(i + 1) / area, fg_sgn * g.aip(i + 2) / area};
        double rip[3] = {g.rip(i), g.rip(i + 1), g.rip(i + 2)};
        double rpnb[3] = {g.xc(enb), g.yc(enb), g.zc(enb)};
        double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
        double ds = std::sqrt(dot3(dr, dr));
        double wt;
        vec_weight(wt, rip, rp, rpnb);
        double drip[3] = {rip[0] - rp[0], rip[1] - rp[1], rip[2] - rp[2]};
        double t = dot3(drip, norm);
        double rp_p[3] = {rip[0] - t * norm[0], rip[1] - t * norm[1], rip[2] - t * norm[2]};
        for (int k = 0; k < 3; ++k) drip[k] = rip[k] - rpnb[k];
        t = dot3(drip, norm);
        double rpnb_p[3] = {rip[0] - t * norm[0], rip[1] - t * norm[1], rip[2] - t * norm[2]};
        double dr_p[3] = {rpnb_p[0] - rp_p[0], rpnb_p[1] - rp_p[1], rpnb_p[2] - rp_p[2]};
        double ds_p = std::sqrt(dot3(dr_p, dr_p));
        f = -fg_sgn * c.mip(fg);
        fnb = std::max(f, 0.0);
        sumf = sumf + f;
        double muip = (1.0 - wt) * c.mu(e) + wt * c.mu(enb);
        d = muip * area / ds;
        double gip[3];
        for (int mm = 1; mm <= 3; ++mm) {  // secondary stress term
          gip[0] = (1.0 - wt) * c.gu(3 * (long)e - 3 + mm) + wt * c.gu(3 * (long)enb - 3 + mm);
          gip[1] = (1.0 - wt) * c.gv(3 * (long)e - 3 + mm) + wt * c.gv(3 * (long)enb - 3 + mm);
          gip[2] = (1.0 - wt) * c.gw(3 * (long)e - 3 + mm) + wt * c.gw(3 * (long)enb - 3 + mm);
          sumss[mm - 1] = sumss[mm - 1] + muip * area * dot3(gip, dr) / ds;
        }
        const A1<double>* G[3] = {&c.gu, &c.gv, &c.gw};  // deferred correction of real diffusion
        for (int q = 0; q < 3; ++q) {
          for (int k = 0; k < 3; ++k) gip[k] = (1.0 - wt) * (*G[q])(3 * (long)e - 2 + k) + wt * (*G[q])(3 * (long)enb - 2 + k);
          sumdefc[q] = sumdefc[q] + muip * area * (dot3(gip, dr_p) / ds_p - dot3(gip, dr) / ds);
        }
      }
      c.anb(idx) = d + fnb;
      c.ap(e) = c.ap(e) + d + fnb;
    }
    double ap0 = c.rho(e) * g.vol(e) / dt;
    c.ap(e) = c.ap(e) + ap0;
    c.bu(e) = ap0 * c.u0(e) + sumf * c.u(e) - g.vol(e) * c.gp(3 * (long)e - 2) + sumss[0] + sumdefc[0];
    c.bv(e) = ap0 * c.v0(e) + sumf * c.v(e) - g.vol(e) * c.gp(3 * (long)e - 1) + sumss[1] + sumdefc[1];
    c.bw(e) = ap0 * c.w0(e) + sumf * c.w(e) - g.vol(e) * c.gp(3 * (long)e) + sumss[2] + sumdefc[2];
  }
  for (size_t ibc = 0; ibc < c.bcs.size(); ++ibc)  // :243-273
    for (int enb = c.bcs[ibc].esec[0]; enb <= c.bcs[ibc].esec[1]; ++enb) {
      int e, lf;
      get_idx(std::abs(g.bs(enb)), e, lf);
      int idx = g.ef2nb_idx(e) + lf - 1;
      int fg = g.ef2nb2(idx);
      long i = 3 * (long)fg - 2;
      double area = std::sqrt(g.aip(i) * g.aip(i) + g.aip(i + 1) * g.aip(i + 1) + g.aip(i + 2) * g.aip(i + 2));
      double norm[3] = {g.aip(i) / area, g.aip(i + 1) / area, g.aip(i + 2) / area};
      double dr[3] = {g.xc(enb) - g.xc(e), g.yc(enb) - g.yc(e), g.zc(enb) - g.zc(e)};
      double ds = std::sqrt(dot3(dr, dr));
      double d = 0.0, f = 0.0;
      if (c.bcs[ibc].bc_type == "dirichlet") {
        f = 0.0;
        d = c.mu(e) * area / ds;
        double vbnc[3] = {c.u(enb), c.v(enb), c.w(enb)};
        double vrel[3] = {c.u(e), c.v(e), c.w(e)};
        double vn = dot3(vrel, norm);
        for (int k = 0; k < 3; ++k) vrel[k] = vrel[k] - vn * norm[k];
        for (int k = 0; k < 3; ++k) vrel[k] = vbnc[k] - vrel[k];
        c.bu(e) = c.bu(e) + d * vrel[0] - d * c.u(e);
        c.bv(e) = c.bv(e) + d * vrel[1] - d * c.v(e);
        c.bw(e) = c.bw(e) + d * vrel[2] - d * c.w(e);
      } else if (c.bcs[ibc].bc_type == "zero_flux") {
        f = 0.0;
        d = c.mu(e) * area / ds;
      } else {
        throw std::runtime_error("calc_coef_uvw: bc_type leaves d,f undefined in the reference");
      }
      c.ap(e) = c.ap(e) + d + f;
      c.anb(idx) = c.anb(idx) + d + f;
    }
  for (int e = 1; e <= g.ne; ++e) {  // :276-284
    c.d(e) = g.vol(e) / c.ap(e);
    c.dc(e) = c.ap(e);
    for (int idx = g.ef2nb_idx(e); idx <= g.ef2nb_idx(e + 1) - 1; ++idx) c.dc(e) = c.dc(e) - c.anb(idx);
    c.dc(e) = g.vol(e) / c.dc(e);
  }
}

void calc_coef_p(Case& c) {  // :289-368
  Mesh& g = c.m;
  for (int e = 1; e <= g.ne; ++e) {
    c.ap(e) = 0.0;
    double sumf = 0.0;
    double rp[3] = {g.xc(e), g.yc(e), g.zc(e)};
    for (int idx = g.ef2nb_idx(e); idx <= g.ef2nb_idx(e + 1) - 1; ++idx) {
      c.anb(idx) = 0.0;
      int enb, lfnb;
      get_idx(g.ef2nb1(idx), enb, lfnb);
      double d = 0.0;
      if (lfnb > 0) {
        int fg = g.ef2nb2(idx);
        int fg_sgn = sgn(fg);
        fg = std::abs(fg);
        long i = 3 * (long)fg - 2;
        double area = std::sqrt(g.aip(i) * g.aip(i) + g.aip(i + 1) * g.aip(i + 1) + g.aip(i + 2) * g.aip(i + 2));
        double norm[3] = {fg_sgn * g.aip(i) / area, fg_sgn * g.aip(i + 1) / area, fg_sgn * g.aip(i + 2) / area};
        double rip[3] = {g.rip(i), g.rip(i + 1), g.rip(i + 2)};
        double rpnb[3] = {g.xc(enb), g.yc(enb), g.zc(enb)};
        double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
        double wt;
        vec_weight(wt, rip, rp, rpnb);
        double f = -fg_sgn * c.mip(fg);
        sumf = sumf + f;
        double rhoip = (1.0 - wt) * c.rho(e) + wt * c.rho(enb);
        d = ((1.0 - wt) * c.dc(e) + wt * c.dc(enb)) / dot3(dr, norm) * rhoip * area;
      }
      c.anb(idx) = d;
      c.ap(e) = c.ap(e) + d;
    }
    c.b(e) = sumf;
  }
  for (size_t ibc = 0; ibc < c.bcs.size(); ++ibc)  // :344-366, d = 0 for both bc types
    for (int enb = c.bcs[ibc].esec[0]; enb <= c.bcs[ibc].esec[1]; ++enb) {
      int e, lf;
      get_idx(std::abs(g.bs(enb)), e, lf);
      int idx = g.ef2nb_idx(e) + lf - 1;
      int fg = g.ef2nb2(idx);
      double d = 0.0;
      c.ap(e) = c.ap(e) + d;
      c.anb(idx) = c.anb(idx) + d;
      c.b(e) = c.b(e) - c.mip(fg);
    }
}

void calc_mip(Case& c, bool lRhieChow) {  // :438-490
  Mesh& g = c.m;
  const double dt = c.dt;
  for (int fg = 1; fg <= g.nf; ++fg) {
    int e, lf, enb, lfnb;
    get_idx(g.s2g(fg), e, lf);
    int idx = g.ef2nb_idx(e) + lf - 1;
    get_idx(g.ef2nb1(idx), enb, lfnb);
    if (lfnb == 0) continue;
    double rp[3] = {g.xc(e), g.yc(e), g.zc(e)};
    double rpnb[3] = {g.xc(enb), g.yc(enb), g.zc(enb)};
    long i = 3 * (long)fg - 2;
    double area = std::sqrt(g.aip(i) * g.aip(i) + g.aip(i + 1) * g.aip(i + 1) + g.aip(i + 2) * g.aip(i + 2));
    double norm[3] = {g.aip(i) / area, g.aip(i + 1) / area, g.aip(i + 2) / area};
    double rip[3] = {g.rip(i), g.rip(i + 1), g.rip(i + 2)};
    double wt;
    vec_weight(wt, rip, rp, rpnb);
    double vec1[3] = {c.u(e), c.v(e), c.w(e)}, vec2[3] = {c.u(enb), c.v(enb), c.w(enb)}, velip[3];
    for (int k = 0; k < 3; ++k) velip[k] = (1.0 - wt) * vec1[k] + wt * vec2[k];
    double rhoip = c.rho(e) * (1.0 - wt) + c.rho(enb) * wt;
    c.mip(fg) = dot3(velip, norm) * rhoip * area;
    if (lRhieChow) {
      double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
      double gpip[3], velip0[3];
      for (int k = 0; k < 3; ++k) gpip[k] = (1.0 - wt) * c.gp(3 * (long)e - 2 + k) + wt * c.gp(3 * (long)enb - 2 + k);
      double dip = (1.0 - wt) * c.d(e) + wt * c.d(enb);
      double v1[3] = {c.u0(e), c.v0(e), c.w0(e)}, v2[3] = {c.u0(enb), c.v0(enb), c.w0(enb)};
      for (int k = 0; k < 3; ++k) velip0[k] = (1.0 - wt) * v1[k] + wt * v2[k];
      c.mip(fg) = c.mip(fg) - rhoip * area * dip / dot3(dr, norm) * (c.p(enb) - c.p(e) - dot3(gpip, dr))
                            - rhoip / dt * dip * (c.mip0(fg) - dot3(velip0, norm) * rhoip * area);
    }
  }
}

void adjust_pc(Case& c, double pref) {  // :136-158 (zeroth-order halo extrapolation branch)
  Mesh& g = c.m;
  for (int e = 1; e <= g.ne; ++e) c.phic(e) = c.phic(e) - pref;
  for (int enb = g.ne + 1; enb <= g.ne + g.nbf; ++enb) {
    int e, lf;
    get_idx(std::abs(g.bs(enb)), e, lf);
    c.phic(enb) = c.phic(e);
  }
}

void update_uvwp(Case& c) {  // :370-436 (cell-velocity correction is if(.false.))
  Mesh& g = c.m;
  for (int e = 1; e <= g.ne; ++e) {
    c.p(e) = c.p(e) + c.phic(e);
    for (int k = 0; k < 3; ++k) c.gp(3 * (long)e - 2 + k) = c.gp(3 * (long)e - 2 + k) + c.gpc(3 * (long)e - 2 + k);
  }
  for (int fg = 1; fg <= g.nf; ++fg) {
    int e, lf, enb, lfnb;
    get_idx(g.s2g(fg), e, lf);
    int idx = g.ef2nb_idx(e) + lf - 1;
    get_idx(g.ef2nb1(idx), enb, lfnb);
    if (lfnb == 0) continue;
    long i = 3 * (long)fg - 2;
    double area = std::sqrt(g.aip(i) * g.aip(i) + g.aip(i + 1) * g.aip(i + 1) + g.aip(i + 2) * g.aip(i + 2));
    double norm[3] = {g.aip(i) / area, g.aip(i + 1) / area, g.aip(i + 2) / area};
    double rp[3] = {g.xc(e), g.yc(e), g.zc(e)};
    double rpnb[3] = {g.xc(enb), g.yc(enb), g.zc(enb)};
    double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
    double rip[3] = {g.rip(i), g.rip(i + 1), g.rip(i + 2)};
    double wt;
    vec_weight(wt, rip, rp, rpnb);
    double dip = (1.0 - wt) * c.dc(e) + wt * c.dc(enb);
    double rhoip = (c.rho(e) + c.rho(enb)) / 2.0;
    double dmip = rhoip * area * dip * (c.phic(enb) - c.phic(e)) / dot3(dr, norm);
    c.mip(fg) = c.mip(fg) - dmip;
  }
}

// BC callbacks dirichlet0 / lid / symmetry, mod_uvwp.f90:493-570
static void energy_boundaries(Case& c);
static void scalar_boundaries(Case& c);
void update_boundaries(Case& c) {  // mod_physics.f90:38-50
  Mesh& g = c.m;
  for (size_t i = 0; i < c.bcs.size(); ++i) {
    BC& bc = c.bcs[i];
    bc.bc_type = (bc.kind == BC_SYMMETRY) ? "zero_flux" : "dirichlet";
    for (int e = bc.esec[0]; e <= bc.esec[1]; ++e) {
      int enb, lfnb;
      get_idx(std::abs(g.bs(e)), enb, lfnb);
      int idx = g.ef2nb_idx(enb) + lfnb - 1;
      int fg = g.ef2nb2(idx);
      if (bc.kind == BC_SYMMETRY) {
        long i3 = 3 * (long)fg - 2;
        double area = std::sqrt(g.aip(i3) * g.aip(i3) + g.aip(i3 + 1) * g.aip(i3 + 1) + g.aip(i3 + 2) * g.aip(i3 + 2));
        double norm[3] = {g.aip(i3) / area, g.aip(i3 + 1) / area, g.aip(i3 + 2) / area};
        double vel[3] = {c.u(enb), c.v(enb), c.w(enb)};
        double vn = dot3(vel, norm);
        double veln[3] = {vn * norm[0], vn * norm[1], vn * norm[2]};
        double velt[3] = {vel[0] - veln[0], vel[1] - veln[1], vel[2] - veln[2]};
        c.u(e) = velt[0] - 2 * veln[0];
        c.v(e) = velt[1] - 2 * veln[1];
        c.w(e) = velt[2] - 2 * veln[2];
      } else {
        c.u(e) = bc.uvw[0]; c.v(e) = bc.uvw[1]; c.w(e) = bc.uvw[2];
      }
      c.p(e) = c.p(enb);
      c.mip(fg) = 0.0;
    }
  }
  if (c.has_energy) energy_boundaries(c);  // :47
  if (c.has_scalar) scalar_boundaries(c);  // :45 (commented out in the reference)
}

void update_time(Case& c) {  // mod_physics.f90:101-112
  c.u0.d = c.u.d; c.v0.d = c.v.d; c.w0.d = c.w.d; c.mip0.d = c.mip.d;
  if (c.has_energy) c.h0.d = c.h.d;  // :110
  if (c.has_scalar) c.s0.d = c.s.d;  // :104 (commented out in the reference, whose scalar equation is never constructed)
}

// ---- energy equation (enthalpy phi = cp*T), src/equations/mod_energy.f90 ---------------------------------------------
void construct_energy(Case& c) {  // :14-48; init_properties: tc = 5, cp = 1000 (mod_properties.f90:88-89)
  Mesh& g = c.m;
  const long H = g.ne + g.nbf;
  c.tc.alloc(g.ne, 5.0); c.cp.alloc(g.ne, 1000.0);
  c.t.alloc(H, 273.0); c.gt.alloc(3 * H, 0.0); c.h.alloc(H, 0.0); c.h0.alloc(H, 0.0); c.gh.alloc(3 * H, 0.0);
  // :33 eqn%phi = eqn%t*prop%cp with arrays of ne+nbf and ne entries (non-conforming in the reference): the halo entries are
  // overwritten by the boundary callbacks before they are read, the cell entries are t*cp
  for (long e = 1; e <= H; ++e) c.h(e) = c.t(e) * c.cp(e <= g.ne ? e : 1);
  c.h0.d = c.h.d;
  c.has_energy = true;
}

static void energy_boundaries(Case& c) {  // lid / dirichlet0 of mod_energy.f90:173-212 ('top' -> 373 K, the others 273 K)
  Mesh& g = c.m;
  for (BC& bc : c.bcs)
    for (int e = bc.esec[0]; e <= bc.esec[1]; ++e) {
      int enb, lfnb;
      get_idx(std::abs(g.bs(e)), enb, lfnb);
      c.t(e) = bc.kind == BC_LID ? 373.0 : 273.0;
      c.h(e) = c.cp(enb) * c.t(e);
    }
}

void calc_coef_energy(Case& c) {  // :82-169
  Mesh& g = c.m;
  const double dt = c.dt;
  for (int e = 1; e <= g.ne; ++e) {
    c.ap(e) = 0.0;
    double sumf = 0.0, sumdefc = 0.0;
    double rp[3] = {g.xc(e), g.yc(e), g.zc(e)};
    for (int idx = g.ef2nb_idx(e); idx <= g.ef2nb_idx(e + 1) - 1; ++idx) {
      c.anb(idx) = 0.0;
      int enb, lfnb;
      get_idx(g.ef2nb1(idx), enb, lfnb);
      if (lfnb == 0) continue;
      int fg = g.ef2nb2(idx);
      int fg_sgn = sgn(fg);
      fg = std::abs(fg);
      long i = 3 * (long)fg - 2;
      double area = std::sqrt(g.aip(i) * g.aip(i) + g.aip(i + 1) * g.aip(i + 1) + g.aip(i + 2) * g.aip(i + 2));
      double norm[3] = {fg_sgn * g.aip(i) / area, fg_sgn * g.aip(i + 1) / area, fg_sgn * g.aip(i + 2) / area};
      double rip[3] = {g.rip(i), g.rip(i + 1), g.rip(i + 2)};
      double rpnb[3] = {g.xc(enb), g.yc(enb), g.zc(enb)};
      double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
      double wt;
      vec_weight(wt, rip, rp, rpnb);
      double f = -fg_sgn * c.mip(fg);
      double fnb = std::max(f, 0.0);
      sumf = sumf + f;
      double tci = (1.0 - wt) * c.tc(e) + wt * c.tc(enb);
      double cpi = (1.0 - wt) * c.cp(e) + wt * c.cp(enb);
      double d = tci / cpi / dot3(dr, norm) * area;
      double ghi[3], gti[3];
      for (int k = 0; k < 3; ++k) {
        ghi[k] = (1.0 - wt) * c.gh(3 * (long)e - 2 + k) + wt * c.gh(3 * (long)enb - 2 + k);
        gti[k] = (1.0 - wt) * c.gt(3 * (long)e - 2 + k) + wt * c.gt(3 * (long)enb - 2 + k);
      }
      sumdefc = sumdefc + tci * area * (dot3(gti, norm) - dot3(ghi, norm) / cpi);
      c.anb(idx) = d + fnb;
      c.ap(e) = c.ap(e) + d + fnb;
    }
    double ap0 = c.rho(e) * g.vol(e) / dt;
    c.ap(e) = c.ap(e) + ap0;
    c.b(e) = ap0 * c.h0(e) + sumf * c.h(e) + sumdefc;
  }
  double d = 0.0, f = 0.0;  // locals of the routine: a 'zero_flux' section would reuse the last values (never the case: both callbacks are 'dirichlet')
  for (size_t ibc = 0; ibc < c.bcs.size(); ++ibc)
    for (int enb = c.bcs[ibc].esec[0]; enb <= c.bcs[ibc].esec[1]; ++enb) {
      int e, lf;
      get_idx(std::abs(g.bs(enb)), e, lf);
      int idx = g.ef2nb_idx(e) + lf - 1;
      int fg = g.ef2nb2(idx);
      long i = 3 * (long)fg - 2;
      double area = std::sqrt(g.aip(i) * g.aip(i) + g.aip(i + 1) * g.aip(i + 1) + g.aip(i + 2) * g.aip(i + 2));
      double norm[3] = {g.aip(i) / area, g.aip(i + 1) / area, g.aip(i + 2) / area};
      double dr[3] = {g.xc(enb) - g.xc(e), g.yc(enb) - g.yc(e), g.zc(enb) - g.zc(e)};
      double ds = dot3(dr, norm);
      f = 0.0;
      d = c.tc(e) * area / ds / c.cp(e);
      c.ap(e) = c.ap(e) + d + f;
      c.anb(idx) = c.anb(idx) + d + f;
    }
}

SolveStat solve_energy(Case& c) {  // :59-80
  Mesh& g = c.m;
  calc_grad(c.t.data(), c.gt.data(), g.xc.data(), g.yc.data(), g.zc.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne);
  calc_grad(c.h.data(), c.gh.data(), g.xc.data(), g.yc.data(), g.zc.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne);
  calc_coef_energy(c);
  SolveStat st = solve_gs(false, c.h.data(), c.ap.data(), c.anb.data(), c.b.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne, c.nit);
  for (int e = 1; e <= g.ne; ++e) c.t(e) = c.h(e) / c.cp(e);  // calc_temperature, mod_properties.f90:214-222
  return st;
}

// ---- passive scalar, src/equations/mod_scalar.f90 ---------------------------------------------------------------------
void construct_scalar(Case& c, double dcoef, const double vel[3], const double* bc_value) {  // :16-46
  Mesh& g = c.m;
  const long H = g.ne + g.nbf;
  c.s.alloc(H, 0.0); c.s0.alloc(H, 0.0); c.gs.alloc(3 * H, 0.0);
  c.s_dcoef = dcoef;
  for (int k = 0; k < 3; ++k) c.s_vel[k] = vel[k];
  c.s_bc.assign(c.bcs.size(), 0.0);
  if (bc_value) for (size_t i = 0; i < c.bcs.size(); ++i) c.s_bc[i] = bc_value[i];
  c.has_scalar = true;
}

static void scalar_boundaries(Case& c) {  // dirichlet0 / dirichlet1, :129-155
  for (size_t i = 0; i < c.bcs.size(); ++i)
    for (int e = c.bcs[i].esec[0]; e <= c.bcs[i].esec[1]; ++e) c.s(e) = c.s_bc[i];
}

void calc_coef_scalar(Case& c) {  // :74-126
  Mesh& g = c.m;
  const double dt = c.dt;
  for (int e = 1; e <= g.ne; ++e) {
    c.ap(e) = 0.0;
    double sumf = 0.0;
    for (int idx = g.ef2nb_idx(e); idx <= g.ef2nb_idx(e + 1) - 1; ++idx) {
      int fg = g.ef2nb2(idx);
      int fg_sgn = sgn(fg);
      fg = std::abs(fg);
      long i = 3 * (long)fg - 2;
      double area = std::sqrt(g.aip(i) * g.aip(i) + g.aip(i + 1) * g.aip(i + 1) + g.aip(i + 2) * g.aip(i + 2));
      double norm[3] = {fg_sgn * g.aip(i) / area, fg_sgn * g.aip(i + 1) / area, fg_sgn * g.aip(i + 2) / area};
      int enb, lfnb;
      get_idx(g.ef2nb1(idx), enb, lfnb);
      double dr[3] = {g.xc(enb) - g.xc(e), g.yc(enb) - g.yc(e), g.zc(enb) - g.zc(e)};
      double mnorm[3] = {-norm[0], -norm[1], -norm[2]};
      double f = dot3(c.s_vel, mnorm) * area;
      double wnb = 0.0;
      if (f > 0.0) wnb = 1.0;
      double fnb = wnb * f;
      sumf = sumf + f;
      double d = c.s_dcoef / dot3(dr, dr) * dot3(dr, norm) * area;
      c.anb(idx) = d + fnb;
      c.ap(e) = c.ap(e) + d + fnb;
    }
    double ap0 = g.vol(e) / dt;
    c.ap(e) = c.ap(e) + ap0;
    c.b(e) = ap0 * c.s0(e) + sumf * c.s(e);
  }
}

SolveStat solve_scalar(Case& c) {  // :56-72 (eqn%name = 'scalar': sor = 1)
  Mesh& g = c.m;
  calc_coef_scalar(c);
  calc_grad(c.s.data(), c.gs.data(), g.xc.data(), g.yc.data(), g.zc.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne);
  return solve_gs(false, c.s.data(), c.ap.data(), c.anb.data(), c.b.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne, c.nit);
}

void solve_uvwp(Case& c, SolveStat st[4]) {  // :95-134
  Mesh& g = c.m;
  calc_coef_uvw(c);
  st[0] = solve_gs(false, c.u.data(), c.ap.data(), c.anb.data(), c.bu.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne, c.nit);
  st[1] = solve_gs(false, c.v.data(), c.ap.data(), c.anb.data(), c.bv.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne, c.nit);
  st[2] = solve_gs(false, c.w.data(), c.ap.data(), c.anb.data(), c.bw.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne, c.nit);
  calc_grad(c.u.data(), c.gu.data(), g.xc.data(), g.yc.data(), g.zc.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne);
  calc_grad(c.v.data(), c.gv.data(), g.xc.data(), g.yc.data(), g.zc.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne);
  calc_grad(c.w.data(), c.gw.data(), g.xc.data(), g.yc.data(), g.zc.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne);
  calc_mip(c, true);
  calc_coef_p(c);
  std::fill(c.phic.d.begin(), c.phic.d.end(), 0.0);  // set_a_0
  st[3] = solve(c, true, c.ap.data(), c.anb.data(), c.b.data(), c.phic.data(), c.nit);
  double pref = c.phic(c.pref_cell);  // phic(1) in the reference (:129)
  adjust_pc(c, pref);
  calc_grad(c.phic.data(), c.gpc.data(), g.xc.data(), g.yc.data(), g.zc.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne);
  update_uvwp(c);
}

void construct_physics(Case& c, int n_subdomains) {  // mod_physics.f90:52-75, mod_uvwp.f90:20-84
  Mesh& g = c.m;
  c.n_subdomains = n_subdomains;
  const long H = g.ne + g.nbf, Z = 2L * g.nf - g.nbf;
  c.ap.alloc(g.ne, 0.0); c.b.alloc(g.ne, 0.0); c.anb.alloc(Z, 0.0); c.phic.alloc(H, 0.0);
  if (n_subdomains > 1) construct_subdomains(c);
  c.rho.alloc(g.ne, 5.0); c.mu.alloc(g.ne, 0.01);  // init_properties, mod_properties.f90:86-87
  for (A1<double>* a : {&c.u, &c.v, &c.w, &c.u0, &c.v0, &c.w0, &c.p}) a->alloc(H, 0.0);
  for (A1<double>* a : {&c.gu, &c.gv, &c.gw, &c.gp, &c.gpc}) a->alloc(3 * H, 0.0);
  c.mip.alloc(g.nf, 0.0); c.mip0.alloc(g.nf, 0.0);
  for (A1<double>* a : {&c.bu, &c.bv, &c.bw, &c.d, &c.dc}) a->alloc(g.ne, 0.0);
  // BCs: eqn%bcs indexed by 2-D section order; make_bc binds by substring match
  // (mod_eqn_setup.f90:46-68); 'top' -> lid, the other five -> dirichlet0 (mod_uvwp.f90:73-78)
  c.bcs.assign(g.nintf_c2b, BC());
  static const char* names[6] = {"top", "west", "east", "south", "north", "bottom"};
  for (int k = 0; k < 6; ++k) {
    bool found = false;
    for (int i = 1; i <= g.nintf_c2b; ++i) {
      int s = g.intf2sec[i - 1];
      std::string sn = g.sectionName[s - 1];
      while (!sn.empty() && sn.back() == ' ') sn.pop_back();
      if (!sn.empty() && std::string(names[k]).find(sn) != std::string::npos) {
        BC& bc = c.bcs[i - 1];
        bc.idx = i; bc.name = names[k];
        bc.esec[0] = g.esec[2 * (s - 1)]; bc.esec[1] = g.esec[2 * (s - 1) + 1];
        bc.kind = (k == 0) ? BC_LID : BC_WALL;
        bc.uvw[0] = (k == 0) ? 1.0 : 0.0; bc.uvw[1] = 0.0; bc.uvw[2] = 0.0;
        found = true;
      }
    }
    if (!found) throw std::runtime_error(std::string("BC ") + names[k] + " not found!");
  }
  calc_mip(c, false);
  c.mip0.d = c.mip.d;
}

}  // namespace orc
