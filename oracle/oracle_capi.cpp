// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points for tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs (loaded with ctypes from oracle/oracle.py).
// Parity: pinned to the reference's source text, see oracle_setup.hpp.
#include "oracle_solver.hpp"
#include <chrono>
#include <cstring>

using namespace orc;

static thread_local std::string g_err;

#define ORC_TRY try {
#define ORC_CATCH(ret) } catch (const std::exception& ex) { g_err = ex.what(); return ret; }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_create(int nvx, const double* x, const double* y, const double* z, int nsec, const int* etype,
                 const int* esec, const char* names /* nsec*32 */, int ne2vx_max, int nelem,
                 const int* e2vx, int n_subdomains) {
  Case* c = nullptr;
  ORC_TRY
  c = new Case;
  Mesh& m = c->m;
  m.nvx = nvx; m.nsec = nsec; m.ne2vx_max = ne2vx_max; m.nelem = nelem;
  m.x.assign(x, x + nvx); m.y.assign(y, y + nvx); m.z.assign(z, z + nvx);
  m.etype.assign(etype, etype + nsec);
  m.esec.assign(esec, esec + 2 * nsec);
  for (int s = 0; s < nsec; ++s) m.sectionName.push_back(std::string(names + 32 * s, 32));
  m.e2vx.assign(e2vx, e2vx + (size_t)ne2vx_max * nelem);
  setup_mesh(m, n_subdomains);
  construct_physics(*c, n_subdomains);
  return c;
  } catch (const std::exception& ex) { g_err = ex.what(); delete c; return nullptr; }
}

// Same as orc_create but starting from already-built geometry_t arrays (the outputs of
// find_element_nb / calc_aip_xyzip_uns / calc_vol_cv_centers_uns), so that timing runs on big
// meshes skip the reference's O(nvx k^2) face matching.  The hot path, the RCB partition and
// construct_subdomains are still the oracle's own.
void* orc_create_from_geom(int ne, int nf, int nbf, const int* ef2nb_idx, const int* ef2nb_nb, const int* ef2nb_fg,
                           const int* s2g, const int* bs, const double* xc, const double* yc, const double* zc,
                           const double* aip, const double* rip, const double* vol, int nsec, const int* etype,
                           const int* esec, const char* names, int n_subdomains) {
  Case* c = nullptr;
  try {
    c = new Case;
    Mesh& m = c->m;
    m.nsec = nsec; m.ne = ne; m.nf = nf; m.nbf = nbf; m.nelem = ne + nbf;
    m.etype.assign(etype, etype + nsec);
    m.esec.assign(esec, esec + 2 * nsec);
    for (int s = 0; s < nsec; ++s) m.sectionName.push_back(std::string(names + 32 * s, 32));
    m.bs_idx.alloc(nsec + 1, 0);
    m.bs_idx(1) = 1;
    for (int s = 1; s <= nsec; ++s) m.bs_idx(s + 1) = m.esec[2 * (s - 1) + 1] + 1;
    for (int s = 1; s <= nsec; ++s) if (m.etype[s - 1] < 10) m.intf2sec.push_back(s);
    m.nintf_c2b = (int)m.intf2sec.size();
    const long Z = 2L * nf - nbf, H = ne + nbf;
    m.ef2nb_idx.alloc(ne + 1); std::copy(ef2nb_idx, ef2nb_idx + ne + 1, m.ef2nb_idx.data());
    m.ef2nb1.alloc(Z); std::copy(ef2nb_nb, ef2nb_nb + Z, m.ef2nb1.data());
    m.ef2nb2.alloc(Z); std::copy(ef2nb_fg, ef2nb_fg + Z, m.ef2nb2.data());
    m.s2g.alloc(nf + 1); std::copy(s2g, s2g + nf, m.s2g.data());
    m.bs.alloc(nbf, 0, ne + 1); std::copy(bs, bs + nbf, m.bs.data());
    m.xc.alloc(H); std::copy(xc, xc + H, m.xc.data());
    m.yc.alloc(H); std::copy(yc, yc + H, m.yc.data());
    m.zc.alloc(H); std::copy(zc, zc + H, m.zc.data());
    m.aip.alloc(3L * nf); std::copy(aip, aip + 3L * nf, m.aip.data());
    m.rip.alloc(3L * nf); std::copy(rip, rip + 3L * nf, m.rip.data());
    m.vol.alloc(ne); std::copy(vol, vol + ne, m.vol.data());
    m.n_subdomains = 1;
    // large timing runs: the reference block-order sort (mod_util.f90:1683) is quadratic in cells per block (hours at 128^3);
    // a stable order keeps the blocks and the per-iteration cost identical, only the sweep order inside a block differs
    if (n_subdomains > 1) rcb_partition(m, n_subdomains, ne > 300000);
    construct_physics(*c, n_subdomains);
    return c;
  } catch (const std::exception& ex) { g_err = ex.what(); delete c; return nullptr; }
}

void orc_destroy(void* h) { delete (Case*)h; }

int orc_dims(void* h, int* ne, int* nf, int* nbf, int* nbc) {
  Case* c = (Case*)h;
  *ne = c->m.ne; *nf = c->m.nf; *nbf = c->m.nbf; *nbc = (int)c->bcs.size();
  return 0;
}

double* orc_real(void* h, const char* name, long* n) {
  Case* c = (Case*)h;
  struct E { const char* k; A1<double>* a; };
  E tab[] = {{"xc", &c->m.xc}, {"yc", &c->m.yc}, {"zc", &c->m.zc}, {"aip", &c->m.aip}, {"rip", &c->m.rip},
             {"vol", &c->m.vol}, {"rho", &c->rho}, {"mu", &c->mu}, {"ap", &c->ap}, {"anb", &c->anb},
             {"b", &c->b}, {"phic", &c->phic}, {"u", &c->u}, {"v", &c->v}, {"w", &c->w}, {"p", &c->p},
             {"gu", &c->gu}, {"gv", &c->gv}, {"gw", &c->gw}, {"gp", &c->gp}, {"gpc", &c->gpc},
             {"mip", &c->mip}, {"mip0", &c->mip0}, {"u0", &c->u0}, {"v0", &c->v0}, {"w0", &c->w0},
             {"bu", &c->bu}, {"bv", &c->bv}, {"bw", &c->bw}, {"d", &c->d}, {"dc", &c->dc},
             {"tc", &c->tc}, {"cp", &c->cp}, {"t", &c->t}, {"gt", &c->gt}, {"h", &c->h}, {"h0", &c->h0}, {"gh", &c->gh},
             {"s", &c->s}, {"s0", &c->s0}, {"gs", &c->gs}};
  for (auto& e : tab)
    if (!std::strcmp(e.k, name)) { *n = e.a->size(); return e.a->data(); }
  *n = -1;
  return nullptr;
}

int* orc_int(void* h, const char* name, long* n) {
  Case* c = (Case*)h;
  struct E { const char* k; A1<int>* a; };
  E tab[] = {{"ef2nb_idx", &c->m.ef2nb_idx}, {"ef2nb_nb", &c->m.ef2nb1}, {"ef2nb_fg", &c->m.ef2nb2},
             {"s2g", &c->m.s2g}, {"bs", &c->m.bs}, {"gf2g", &c->m.gf2g}, {"g2gf_p", &c->m.g2gf_p},
             {"g2gf_idx", &c->m.g2gf_idx}};
  for (auto& e : tab)
    if (!std::strcmp(e.k, name)) {
      *n = e.a->size();
      if (!std::strcmp(name, "s2g")) *n = c->m.nf;  // allocated ns+1 in the reference
      return e.a->data();
    }
  *n = -1;
  return nullptr;
}

int orc_bc_table(void* h, int* esec, int* kind, double* uvw) {
  Case* c = (Case*)h;
  for (size_t i = 0; i < c->bcs.size(); ++i) {
    esec[2 * i] = c->bcs[i].esec[0]; esec[2 * i + 1] = c->bcs[i].esec[1];
    kind[i] = c->bcs[i].kind;
    for (int k = 0; k < 3; ++k) uvw[3 * i + k] = c->bcs[i].uvw[k];
  }
  return 0;
}

int orc_set_bc(void* h, int i, int kind, double u, double v, double w) {
  Case* c = (Case*)h;
  if (i < 0 || i >= (int)c->bcs.size()) return 1;
  c->bcs[i].kind = kind; c->bcs[i].uvw[0] = u; c->bcs[i].uvw[1] = v; c->bcs[i].uvw[2] = w;
  return 0;
}

int orc_set_param(void* h, const char* key, double v) {
  Case* c = (Case*)h;
  if (!std::strcmp(key, "dt")) c->dt = v;
  else if (!std::strcmp(key, "nit")) c->nit = (int)v;
  else if (!std::strcmp(key, "pref_cell")) c->pref_cell = (int)v;
  else return 1;
  return 0;
}

static void put_stats(const SolveStat* st, int n, double* out) {
  for (int i = 0; i < n; ++i) {
    out[4 * i] = st[i].it; out[4 * i + 1] = st[i].res_i; out[4 * i + 2] = st[i].res_f; out[4 * i + 3] = st[i].res_max;
  }
}

int orc_update_boundaries(void* h) { ORC_TRY update_boundaries(*(Case*)h); return 0; ORC_CATCH(1) }
int orc_update_time(void* h) { ORC_TRY update_time(*(Case*)h); return 0; ORC_CATCH(1) }
int orc_construct_energy(void* h) { ORC_TRY construct_energy(*(Case*)h); return 0; ORC_CATCH(1) }
int orc_solve_energy(void* h, double* out4) { ORC_TRY SolveStat st = solve_energy(*(Case*)h); if (out4) put_stats(&st, 1, out4); return 0; ORC_CATCH(1) }
int orc_construct_scalar(void* h, double dcoef, const double* vel, const double* bc_value) {
  ORC_TRY construct_scalar(*(Case*)h, dcoef, vel, bc_value); return 0; ORC_CATCH(1)
}
int orc_solve_scalar(void* h, double* out4) { ORC_TRY SolveStat st = solve_scalar(*(Case*)h); if (out4) put_stats(&st, 1, out4); return 0; ORC_CATCH(1) }
int orc_solve_uvwp(void* h, double* hist16) {
  ORC_TRY
  SolveStat st[4];
  solve_uvwp(*(Case*)h, st);
  if (hist16) put_stats(st, 4, hist16);
  return 0;
  ORC_CATCH(1)
}
// main.f90:50-63 ; returns seconds spent in the loop through *seconds (may be null)
int orc_run(void* h, int ntstep, int ncoef, double* hist, double* seconds) {
  ORC_TRY
  Case& c = *(Case*)h;
  auto t0 = std::chrono::steady_clock::now();
  for (int ts = 0; ts < ntstep; ++ts) {
    for (int ic = 0; ic < ncoef; ++ic) {
      update_boundaries(c);
      SolveStat st[4];
      solve_uvwp(c, st);
      if (hist) put_stats(st, 4, hist + 16 * ((size_t)ts * ncoef + ic));
    }
    update_time(c);
  }
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
  ORC_CATCH(1)
}

int orc_calc_coef_uvw(void* h) { ORC_TRY calc_coef_uvw(*(Case*)h); return 0; ORC_CATCH(1) }
int orc_calc_coef_p(void* h) { ORC_TRY calc_coef_p(*(Case*)h); return 0; ORC_CATCH(1) }
int orc_calc_mip(void* h, int lrc) { ORC_TRY calc_mip(*(Case*)h, lrc != 0); return 0; ORC_CATCH(1) }
int orc_adjust_pc(void* h) { ORC_TRY Case& c = *(Case*)h; adjust_pc(c, c.phic(1)); return 0; ORC_CATCH(1) }
int orc_update_uvwp(void* h) { ORC_TRY update_uvwp(*(Case*)h); return 0; ORC_CATCH(1) }
int orc_calc_grad(void* h, const double* phi, double* grad) {
  ORC_TRY
  Mesh& g = ((Case*)h)->m;
  calc_grad(phi, grad, g.xc.data(), g.yc.data(), g.zc.data(), g.ef2nb_idx.data(), g.ef2nb1.data(), g.ne);
  return 0;
  ORC_CATCH(1)
}
// solve(cname, subdomain, intf, geom, ap, anb, b, phi, nsubd, nit) with the case's subdomains
int orc_solve(void* h, int is_pc, const double* ap, const double* anb, const double* b, double* phi, int nit, double* out4) {
  ORC_TRY
  SolveStat st = solve(*(Case*)h, is_pc != 0, ap, anb, b, phi, nit);
  put_stats(&st, 1, out4);
  return 0;
  ORC_CATCH(1)
}

// flat-signature routines on caller arrays (1-based content, reference layout)
int orc_flat_calc_grad(const double* phi, double* grad, const double* xc, const double* yc, const double* zc,
                       const int* ef2nb_idx, const int* ef2nb1, int ne) {
  ORC_TRY calc_grad(phi, grad, xc, yc, zc, ef2nb_idx, ef2nb1, ne); return 0; ORC_CATCH(1)
}
int orc_flat_solve_gs(int is_pc, double* phi, const double* ap, const double* anb, const double* b,
                      const int* ef2nb_idx, const int* ef2nb1, int ne, int nit, double* out4) {
  ORC_TRY
  SolveStat st = solve_gs(is_pc != 0, phi, ap, anb, b, ef2nb_idx, ef2nb1, ne, nit);
  put_stats(&st, 1, out4);
  return 0;
  ORC_CATCH(1)
}
int orc_flat_smoother_gs(int is_pc, double* phi, const double* ap, const double* anb, const double* b,
                         const int* ef2nb_idx, const int* ef2nb1, int ne, int nit) {
  ORC_TRY smoother_gs(is_pc != 0, phi, ap, anb, b, ef2nb_idx, ef2nb1, ne, nit); return 0; ORC_CATCH(1)
}
int orc_flat_calc_residual(const double* phi, const double* ap, const double* anb, const double* b,
                           const int* ef2nb_idx, const int* ef2nb1, int ne, double* res, double* res_max) {
  ORC_TRY calc_residual(phi, ap, anb, b, ef2nb_idx, ef2nb1, ne, *res, *res_max); return 0; ORC_CATCH(1)
}

}  // extern "C"
