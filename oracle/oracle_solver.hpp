// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by, or called from the product.
// CPU restatement (FP64, one thread, -ffp-contract=off) of the reference's per-iteration hot
// path: src/equations/mod_uvwp.f90, src/modules/mod_solver.f90, src/modules/mod_subdomains.f90,
// src/modules/mod_physics.f90:38-112.  PARITY PINNED TO THE REFERENCE'S SOURCE TEXT: the reference has no goldens and
// no Fortran compiler exists here, so its unmodified files are executed by oracle/f90run/f90py.py and this restatement
// must reproduce every field, matrix and residual record of those runs bit for bit
// (tests/test_oracle_vs_reference_source.py, fixtures tests/golden/ref_*.npz); plus the analytic KATs of tests/test_oracle_kat.py.
#pragma once
#include "oracle_setup.hpp"

namespace orc {

// subdomain_t / intf_t, mod_subdomains.f90:6-16
struct Subdomain {
  int id = 0, ne = 0, nf = 0, nbf = 0;
  A1<int> ef2nb, ef2nb_idx;
  A1<double> ap, anb, b, phic;
};
struct Intf {
  int c1 = 0, c2 = 0, ncs = 0;
  std::vector<int> index1;         // index2 of (c,cnb) is index1 of (cnb,c)
};

// bc_t, mod_eqn_setup.f90:8-14 ; coef callback represented by `kind`
enum BcKind { BC_WALL = 0, BC_LID = 1, BC_SYMMETRY = 2 };
struct BC {
  int kind = BC_WALL;
  std::string bc_type;             // set by the callback: 'dirichlet' / 'zero_flux'
  int idx = 0;
  int esec[2] = {0, 0};
  std::string name;
  double uvw[3] = {0, 0, 0};       // lid velocity (reference: 1,0,0)
};

struct SolveStat { int it = 0; double res_i = 0, res_f = 0, res_max = 0; };

struct Case {
  Mesh m;
  // phys_t, mod_physics.f90:13-35
  double dt = 0.01;
  int pref_cell = 1;  // reference: pref = phic(1), mod_uvwp.f90:129; tests of renumbered meshes point it at the cell that was cell 1
  int ntstep = 10, ncoef = 3, nit = 100, n_subdomains = 4;
  A1<double> ap, anb, b, phic;
  std::vector<Subdomain> subdomain;
  std::vector<Intf> intf;          // (P,P) row-major: intf[(c-1)*P + (cnb-1)]
  A1<double> rho, mu;              // properties_t (mod_properties.f90:86-87)
  // uvwp_t, mod_uvwp.f90:6-16
  A1<double> u, v, w, p, gu, gv, gw, gp, gpc, mip, mip0, u0, v0, w0, bu, bv, bw, d, dc;
  std::vector<BC> bcs;
  // energy_t (mod_energy.f90:5-10; phi = enthalpy cp*T) and scalar_t (mod_scalar.f90:5-12), constructed on request
  bool has_energy = false, has_scalar = false;
  A1<double> tc, cp;               // properties_t (mod_properties.f90:88-89)
  A1<double> t, gt, h, h0, gh;     // energy: t, gt, phi, phi0, grad
  A1<double> s, s0, gs;            // scalar: phi, phi0, grad
  double s_dcoef = 1.0, s_vel[3] = {0.0, 0.0, -100.0};
  std::vector<double> s_bc;        // scalar: Dirichlet value per boundary section (dirichlet0 / dirichlet1)
  Intf& I(int c, int cnb) { return intf[(size_t)(c - 1) * n_subdomains + (cnb - 1)]; }
};

// construct_physics (mod_physics.f90:52-75) + construct_uvwp (mod_uvwp.f90:20-84)
void construct_physics(Case& c, int n_subdomains);
void construct_subdomains(Case& c);                       // mod_subdomains.f90:18-160
void update_boundaries(Case& c);                          // mod_physics.f90:38-50
void update_time(Case& c);                                // mod_physics.f90:101-112
void solve_uvwp(Case& c, SolveStat st[4]);                // mod_uvwp.f90:95-134
void calc_coef_uvw(Case& c);                              // mod_uvwp.f90:161-286
void calc_coef_p(Case& c);                                // mod_uvwp.f90:289-368
void calc_mip(Case& c, bool lRhieChow);                   // mod_uvwp.f90:438-490
void adjust_pc(Case& c, double pref);                     // mod_uvwp.f90:136-158
void update_uvwp(Case& c);                                // mod_uvwp.f90:370-436

void construct_energy(Case& c);                           // mod_energy.f90:14-48 (after construct_physics)
SolveStat solve_energy(Case& c);                          // mod_energy.f90:59-80
void calc_coef_energy(Case& c);                           // mod_energy.f90:82-169
void construct_scalar(Case& c, double dcoef, const double vel[3], const double* bc_value);  // mod_scalar.f90:16-46
SolveStat solve_scalar(Case& c);                          // mod_scalar.f90:56-72
void calc_coef_scalar(Case& c);                           // mod_scalar.f90:74-126

// flat-signature routines of mod_solver.f90 (arrays 1-based through the raw pointers: p[i-1])
void calc_grad(const double* phi, double* grad, const double* xc, const double* yc, const double* zc,
               const int* ef2nb_idx, const int* ef2nb1, int ne);                                   // :40-81
SolveStat solve_gs(bool is_pc, double* phi, const double* ap, const double* anb, const double* b,
                   const int* ef2nb_idx, const int* ef2nb1, int ne, int nit);                      // :255-327
void smoother_gs(bool is_pc, double* phi, const double* ap, const double* anb, const double* b,
                 const int* ef2nb_idx, const int* ef2nb1, int ne, int nit);                        // :191-228
void calc_residual(const double* phi, const double* ap, const double* anb, const double* b,
                   const int* ef2nb_idx, const int* ef2nb1, int ne, double& res, double& res_max); // :230-253
SolveStat multi_subdomain_solver(Case& c, bool is_pc, const double* ap, const double* anb,
                                 const double* b, double* phi, int nit);                           // :124-189
SolveStat solve(Case& c, bool is_pc, const double* ap, const double* anb, const double* b,
                double* phi, int nit);                                                             // :329-344

}  // namespace orc
