// Assembly kernels of one SIMPLE iteration (sm_100a).  Every kernel is an atomics-free gather:
// cell kernels write only per-cell / per-slot outputs, face kernels only per-face outputs —
// the same write pattern as the reference's loops, so results are deterministic.  Arithmetic
// follows the Fortran expressions operation by operation (compiled with -fmad=false), which
// keeps FP64 results bit-comparable with the reference's unfused evaluation.
#include "state.h"
#include "device_math.cuh"

namespace cfdl {

#define TPB 256

// ---- update_boundaries: BC callbacks dirichlet0 / lid / symmetry, mod_uvwp.f90:493-570 --------
__global__ void __launch_bounds__(TPB) bc_kernel(int B, int N, const int32_t* __restrict__ halo_cell,
                                                 const int32_t* __restrict__ halo_face, const int32_t* __restrict__ halo_bc,
                                                 const int32_t* __restrict__ bc_kind, const double* __restrict__ bc_uvw,
                                                 const double* __restrict__ aip, double* u, double* v, double* w, double* p,
                                                 double* mip) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < B; j += gridDim.x * blockDim.x) {
    int bc = halo_bc[j];
    if (bc < 0) continue;
    int c = halo_cell[j], f = halo_face[j], h = N + j;
    if (bc_kind[bc] == CFDL_BC_SYMMETRY) {
      double a[3];
      load3(aip, f, a);
      double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      double norm[3] = {a[0] / area, a[1] / area, a[2] / area};
      double vel[3] = {u[c], v[c], w[c]};
      double vn = dot3(vel, norm);
      double veln[3] = {vn * norm[0], vn * norm[1], vn * norm[2]};
      u[h] = (vel[0] - veln[0]) - 2 * veln[0];
      v[h] = (vel[1] - veln[1]) - 2 * veln[1];
      w[h] = (vel[2] - veln[2]) - 2 * veln[2];
    } else {
      u[h] = bc_uvw[3 * bc]; v[h] = bc_uvw[3 * bc + 1]; w[h] = bc_uvw[3 * bc + 2];
    }
    p[h] = p[c];
    mip[f] = 0.0;
  }
}

int k_update_boundaries(Handle* h) {
  if (h->B == 0) return CFDL_OK;
  bc_kernel<<<grid_for(h, h->B, TPB), TPB, 0, S(h)>>>(h->B, h->Nc, h->halo_cell, h->halo_face, h->halo_bc, h->bc_kind,
                                                            h->bc_uvw, h->aip, h->fld[CFDL_F_U], h->fld[CFDL_F_V],
                                                            h->fld[CFDL_F_W], h->fld[CFDL_F_P], h->fld[CFDL_F_MIP]);
  CFDL_CUDA(cudaGetLastError());
  if (h->has_energy || h->has_scalar) return k_transport_boundaries(h);  // mod_physics.f90:45,47
  return CFDL_OK;
}

// ---- calc_coef_uvw, mod_uvwp.f90:161-286 ------------------------------------------------------
struct UvwArgs {
  int N, Nc, Np;  // owned cells, owned+ghost cells (halo indices start at Nc), ELL column stride
  const int32_t *ell_nb, *ell_fs, *halo_bc, *bc_kind;
  const uint8_t* nfc;
  const double *xc, *yc, *zc, *aip, *rip, *vol, *rho, *mu;
  const double *u, *v, *w, *u0, *v0, *w0, *gu, *gv, *gw, *gp, *mip;
  double *ap, *anb, *bu, *bv, *bw, *d, *dc;
  double dt;
};

template <int K>
__global__ void __launch_bounds__(TPB) coef_uvw_kernel(const UvwArgs A) {
  const int N = A.N, Nc = A.Nc, Np = A.Np;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    const int n = A.nfc[c];
    const double rp[3] = {A.xc[c], A.yc[c], A.zc[c]};
    const double mu_e = A.mu[c];
    double gue[3], gve[3], gwe[3];
    load3(A.gu, c, gue); load3(A.gv, c, gve); load3(A.gw, c, gwe);
    double ap = 0.0, sumf = 0.0, sumss[3] = {0, 0, 0}, sumdefc[3] = {0, 0, 0};
    double anbk[K];
    int nbk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      anbk[k] = 0.0;
      nbk[k] = -1;
      if (k < n) {
        const int nb = A.ell_nb[(size_t)k * Np + c];
        const int fs = A.ell_fs[(size_t)k * Np + c];
        nbk[k] = nb;
        double d = 0.0, fnb = 0.0;
        if (nb < Nc) {  // lfnb > 0 (owned or ghost cell)
          const int f = abs(fs) - 1;
          const double sg = fs > 0 ? 1.0 : -1.0;
          double a[3], rip[3];
          load3(A.aip, f, a); load3(A.rip, f, rip);
          const double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
          const double norm[3] = {sg * a[0] / area, sg * a[1] / area, sg * a[2] / area};
          const double rpnb[3] = {A.xc[nb], A.yc[nb], A.zc[nb]};
          const double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
          const double ds = sqrt(dot3(dr, dr));
          const double wt = vec_weight(rip, rp, rpnb);
          double drip[3] = {rip[0] - rp[0], rip[1] - rp[1], rip[2] - rp[2]};
          double t = dot3(drip, norm);
          const double rp_p[3] = {rip[0] - t * norm[0], rip[1] - t * norm[1], rip[2] - t * norm[2]};
          drip[0] = rip[0] - rpnb[0]; drip[1] = rip[1] - rpnb[1]; drip[2] = rip[2] - rpnb[2];
          t = dot3(drip, norm);
          const double rpnb_p[3] = {rip[0] - t * norm[0], rip[1] - t * norm[1], rip[2] - t * norm[2]};
          const double dr_p[3] = {rpnb_p[0] - rp_p[0], rpnb_p[1] - rp_p[1], rpnb_p[2] - rp_p[2]};
          const double ds_p = sqrt(dot3(dr_p, dr_p));
          const double f_in = -sg * A.mip[f];  // inward flux
          fnb = fmax(f_in, 0.0);               // upwind bias
          sumf = sumf + f_in;
          const double muip = (1.0 - wt) * mu_e + wt * A.mu[nb];
          d = muip * area / ds;
          double gun[3], gvn[3], gwn[3];
          load3(A.gu, nb, gun); load3(A.gv, nb, gvn); load3(A.gw, nb, gwn);
          const double w1 = 1.0 - wt;
#pragma unroll
          for (int m = 0; m < 3; ++m) {  // secondary stress term, :213-218
            const double gip[3] = {w1 * gue[m] + wt * gun[m], w1 * gve[m] + wt * gvn[m], w1 * gwe[m] + wt * gwn[m]};
            sumss[m] = sumss[m] + muip * area * dot3(gip, dr) / ds;
          }
          {  // deferred correction of real diffusion, :219-225
            double gip[3] = {w1 * gue[0] + wt * gun[0], w1 * gue[1] + wt * gun[1], w1 * gue[2] + wt * gun[2]};
            sumdefc[0] = sumdefc[0] + muip * area * (dot3(gip, dr_p) / ds_p - dot3(gip, dr) / ds);
            gip[0] = w1 * gve[0] + wt * gvn[0]; gip[1] = w1 * gve[1] + wt * gvn[1]; gip[2] = w1 * gve[2] + wt * gvn[2];
            sumdefc[1] = sumdefc[1] + muip * area * (dot3(gip, dr_p) / ds_p - dot3(gip, dr) / ds);
            gip[0] = w1 * gwe[0] + wt * gwn[0]; gip[1] = w1 * gwe[1] + wt * gwn[1]; gip[2] = w1 * gwe[2] + wt * gwn[2];
            sumdefc[2] = sumdefc[2] + muip * area * (dot3(gip, dr_p) / ds_p - dot3(gip, dr) / ds);
          }
        }
        anbk[k] = d + fnb;
        ap = ap + d + fnb;
      }
    }
    const double vol = A.vol[c];
    const double ap0 = A.rho[c] * vol / A.dt;
    ap = ap + ap0;
    const double ue = A.u[c], ve = A.v[c], we = A.w[c];
    double bu = ap0 * A.u0[c] + sumf * ue - vol * A.gp[3 * (size_t)c] + sumss[0] + sumdefc[0];
    double bv = ap0 * A.v0[c] + sumf * ve - vol * A.gp[3 * (size_t)c + 1] + sumss[1] + sumdefc[1];
    double bw = ap0 * A.w0[c] + sumf * we - vol * A.gp[3 * (size_t)c + 2] + sumss[2] + sumdefc[2];
    // boundary faces, visited in halo order like the reference's BC loop (:243-273)
    int last = -1;
    for (int t = 0; t < K; ++t) {
      int best = 0x7fffffff, bk = -1;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (nbk[k] >= Nc && nbk[k] > last && nbk[k] < best) { best = nbk[k]; bk = k; }
      if (bk < 0) break;
      last = best;
      const int bc = A.halo_bc[best - Nc];
      if (bc < 0) continue;
      const int f = A.ell_fs[(size_t)bk * Np + c] - 1;  // always outward on boundary
      double a[3];
      load3(A.aip, f, a);
      const double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      const double norm[3] = {a[0] / area, a[1] / area, a[2] / area};
      const double dr[3] = {A.xc[best] - rp[0], A.yc[best] - rp[1], A.zc[best] - rp[2]};
      const double ds = sqrt(dot3(dr, dr));
      const double d = mu_e * area / ds;
      if (A.bc_kind[bc] != CFDL_BC_SYMMETRY) {  // 'dirichlet'
        const double vbnc[3] = {A.u[best], A.v[best], A.w[best]};
        double vrel[3] = {ue, ve, we};
        const double vn = dot3(vrel, norm);
        vrel[0] = vrel[0] - vn * norm[0]; vrel[1] = vrel[1] - vn * norm[1]; vrel[2] = vrel[2] - vn * norm[2];
        vrel[0] = vbnc[0] - vrel[0]; vrel[1] = vbnc[1] - vrel[1]; vrel[2] = vbnc[2] - vrel[2];
        bu = bu + d * vrel[0] - d * ue;
        bv = bv + d * vrel[1] - d * ve;
        bw = bw + d * vrel[2] - d * we;
      }
      ap = ap + d;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k == bk) anbk[k] = anbk[k] + d;
    }
    // Rhie-Chow d and SIMPLEC dc, :276-284
    double dcv = ap;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (k < n) { dcv = dcv - anbk[k]; A.anb[(size_t)k * Np + c] = anbk[k]; }
    A.ap[c] = ap;
    A.bu[c] = bu; A.bv[c] = bv; A.bw[c] = bw;
    A.d[c] = vol / ap;
    A.dc[c] = vol / dcv;
  }
}

int k_calc_coef_uvw(Handle* h, double dt) {
  UvwArgs A;
  A.N = h->N; A.Nc = h->Nc; A.Np = h->Np;
  A.ell_nb = h->ell_nb; A.ell_fs = h->ell_fs; A.halo_bc = h->halo_bc; A.bc_kind = h->bc_kind; A.nfc = h->nfc;
  A.xc = h->xc; A.yc = h->yc; A.zc = h->zc; A.aip = h->aip; A.rip = h->rip; A.vol = h->vol; A.rho = h->rho; A.mu = h->mu;
  A.u = h->fld[CFDL_F_U]; A.v = h->fld[CFDL_F_V]; A.w = h->fld[CFDL_F_W];
  A.u0 = h->fld[CFDL_F_U0]; A.v0 = h->fld[CFDL_F_V0]; A.w0 = h->fld[CFDL_F_W0];
  A.gu = h->fld[CFDL_F_GU]; A.gv = h->fld[CFDL_F_GV]; A.gw = h->fld[CFDL_F_GW]; A.gp = h->fld[CFDL_F_GP];
  A.mip = h->fld[CFDL_F_MIP];
  A.ap = h->fld[CFDL_F_AP]; A.anb = h->fld[CFDL_F_ANB]; A.bu = h->fld[CFDL_F_BU]; A.bv = h->fld[CFDL_F_BV];
  A.bw = h->fld[CFDL_F_BW]; A.d = h->fld[CFDL_F_D]; A.dc = h->fld[CFDL_F_DC];
  A.dt = dt;
  const int g = grid_for(h, h->N, TPB);
  if (h->K > 6) return fail(CFDL_ERR_UNSUPPORTED, "cells with more than 6 faces are not supported");
  h->pc_sumap_ok = false;  // ap, anb become the momentum matrix
  prof_begin(h, PROF_COEF_UVW);
  if (h->use_statics && h->fs_area) {  // on precomputed face statics (kernels_statics.cu)
    int rc = k_calc_coef_uvw_statics(h, dt);
    if (rc) return rc;
  } else {                             // statics = 0: the reference's form, geometry recomputed per face
    if (h->K <= 4) coef_uvw_kernel<4><<<g, TPB, 0, S(h)>>>(A);
    else coef_uvw_kernel<6><<<g, TPB, 0, S(h)>>>(A);
  }
  prof_end(h);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- calc_coef_p, mod_uvwp.f90:289-368 --------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(TPB) coef_p_kernel(int N, int Nc, int Np, const int32_t* __restrict__ ell_nb,
                                                     const int32_t* __restrict__ ell_fs, const uint8_t* __restrict__ nfc,
                                                     const int32_t* __restrict__ halo_bc,
                                                     const double* __restrict__ xc, const double* __restrict__ yc,
                                                     const double* __restrict__ zc, const double* __restrict__ aip,
                                                     const double* __restrict__ rip, const double* __restrict__ rho,
                                                     const double* __restrict__ dc, const double* __restrict__ mip,
                                                     double* __restrict__ ap_o, double* __restrict__ anb_o, double* __restrict__ b_o) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    const int n = nfc[c];
    const double rp[3] = {xc[c], yc[c], zc[c]};
    const double rho_e = rho[c], dc_e = dc[c];
    double ap = 0.0, sumf = 0.0;
    int nbk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      nbk[k] = -1;
      if (k < n) {
        const int nb = ell_nb[(size_t)k * Np + c];
        const int fs = ell_fs[(size_t)k * Np + c];
        nbk[k] = nb;
        double d = 0.0;
        if (nb < Nc) {
          const int f = abs(fs) - 1;
          const double sg = fs > 0 ? 1.0 : -1.0;
          double a[3], r[3];
          load3(aip, f, a); load3(rip, f, r);
          const double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
          const double norm[3] = {sg * a[0] / area, sg * a[1] / area, sg * a[2] / area};
          const double rpnb[3] = {xc[nb], yc[nb], zc[nb]};
          const double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
          const double wt = vec_weight(r, rp, rpnb);
          const double f_in = -sg * mip[f];
          sumf = sumf + f_in;
          const double rhoip = (1.0 - wt) * rho_e + wt * rho[nb];
          d = ((1.0 - wt) * dc_e + wt * dc[nb]) / dot3(dr, norm) * rhoip * area;
        }
        anb_o[(size_t)k * Np + c] = d;
        ap = ap + d;
      }
    }
    double b = sumf;
    int last = -1;  // boundary loop :344-366 : d = 0, b -= mip(fg), in halo order
    for (int t = 0; t < K; ++t) {
      int best = 0x7fffffff, bk = -1;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (nbk[k] >= Nc && nbk[k] > last && nbk[k] < best) { best = nbk[k]; bk = k; }
      if (bk < 0) break;
      last = best;
      if (halo_bc[best - Nc] < 0) continue;
      b = b - mip[ell_fs[(size_t)bk * Np + c] - 1];
    }
    ap_o[c] = ap;
    b_o[c] = b;
  }
}

int k_calc_coef_p(Handle* h) {
  const int g = grid_for(h, h->N, TPB);
#define CP_ARGS h->N, h->Nc, h->Np, h->ell_nb, h->ell_fs, h->nfc, h->halo_bc, h->xc, h->yc, h->zc, h->aip, h->rip, h->rho, \
                h->fld[CFDL_F_DC], h->fld[CFDL_F_MIP], h->fld[CFDL_F_AP], h->fld[CFDL_F_ANB], h->fld[CFDL_F_B]
  h->pc_sumap_ok = true;  // every form of calc_coef_p builds ap as the slot-order sum of the anb it stores
  prof_begin(h, PROF_COEF_P);
  if (h->use_statics && h->fs_area && h->K <= 6) { int rc = k_calc_coef_p_statics(h); if (rc) return rc; }
  else if (h->K <= 4) coef_p_kernel<4><<<g, TPB, 0, S(h)>>>(CP_ARGS);
  else if (h->K <= 6) coef_p_kernel<6><<<g, TPB, 0, S(h)>>>(CP_ARGS);
  else return fail(CFDL_ERR_UNSUPPORTED, "cells with more than 6 faces are not supported");
  prof_end(h);
#undef CP_ARGS
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- calc_mip (Rhie-Chow), mod_uvwp.f90:438-490 : one thread per interior face ----------------
struct MipArgs {
  int Fi;
  const int32_t *face_a, *face_b;
  const double *xc, *yc, *zc, *aip, *rip, *rho;
  const double *u, *v, *w, *u0, *v0, *w0, *p, *gp, *d, *mip0;
  double* mip;
  double dt;
  int rhie_chow;
};

__global__ void __launch_bounds__(TPB) mip_kernel(const MipArgs A) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < A.Fi; f += gridDim.x * blockDim.x) {
    const int e = A.face_a[f], nb = A.face_b[f];
    const double rp[3] = {A.xc[e], A.yc[e], A.zc[e]};
    const double rpnb[3] = {A.xc[nb], A.yc[nb], A.zc[nb]};
    double a[3], rip[3];
    load3(A.aip, f, a); load3(A.rip, f, rip);
    const double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    const double norm[3] = {a[0] / area, a[1] / area, a[2] / area};
    const double wt = vec_weight(rip, rp, rpnb);
    const double w1 = 1.0 - wt;
    const double velip[3] = {w1 * A.u[e] + wt * A.u[nb], w1 * A.v[e] + wt * A.v[nb], w1 * A.w[e] + wt * A.w[nb]};
    const double rhoip = A.rho[e] * w1 + A.rho[nb] * wt;
    double m = dot3(velip, norm) * rhoip * area;
    if (A.rhie_chow) {
      const double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
      double ge[3], gn[3];
      load3(A.gp, e, ge); load3(A.gp, nb, gn);
      const double gpip[3] = {w1 * ge[0] + wt * gn[0], w1 * ge[1] + wt * gn[1], w1 * ge[2] + wt * gn[2]};
      const double dip = w1 * A.d[e] + wt * A.d[nb];
      const double velip0[3] = {w1 * A.u0[e] + wt * A.u0[nb], w1 * A.v0[e] + wt * A.v0[nb], w1 * A.w0[e] + wt * A.w0[nb]};
      m = m - rhoip * area * dip / dot3(dr, norm) * (A.p[nb] - A.p[e] - dot3(gpip, dr))
            - rhoip / A.dt * dip * (A.mip0[f] - dot3(velip0, norm) * rhoip * area);
    }
    A.mip[f] = m;
  }
}

int k_calc_mip(Handle* h, bool rhie_chow, double dt) {
  if (h->Fi == 0) return CFDL_OK;
  MipArgs A;
  A.Fi = h->Fi; A.face_a = h->face_a; A.face_b = h->face_b;
  A.xc = h->xc; A.yc = h->yc; A.zc = h->zc; A.aip = h->aip; A.rip = h->rip; A.rho = h->rho;
  A.u = h->fld[CFDL_F_U]; A.v = h->fld[CFDL_F_V]; A.w = h->fld[CFDL_F_W];
  A.u0 = h->fld[CFDL_F_U0]; A.v0 = h->fld[CFDL_F_V0]; A.w0 = h->fld[CFDL_F_W0];
  A.p = h->fld[CFDL_F_P]; A.gp = h->fld[CFDL_F_GP]; A.d = h->fld[CFDL_F_D]; A.mip0 = h->fld[CFDL_F_MIP0];
  A.mip = h->fld[CFDL_F_MIP]; A.dt = dt; A.rhie_chow = rhie_chow ? 1 : 0;
  prof_begin(h, PROF_MIP);
  if (h->use_statics && h->fs_area) { int rc = k_calc_mip_statics(h, rhie_chow, dt); if (rc) return rc; }
  else mip_kernel<<<grid_for(h, h->Fi, TPB), TPB, 0, S(h)>>>(A);
  prof_end(h);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- adjust_pc, mod_uvwp.f90:129-130,136-158 --------------------------------------------------
__global__ void copy_scalar_kernel(double* dst, const double* src) { *dst = *src; }
__global__ void __launch_bounds__(TPB) shift_kernel(int N, double* pc, const double* __restrict__ pref) {
  const double r = *pref;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) pc[c] = pc[c] - r;
}
__global__ void __launch_bounds__(TPB) halo_copy_kernel(int B, int N, const int32_t* __restrict__ halo_cell, double* pc) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < B; j += gridDim.x * blockDim.x) pc[N + j] = pc[halo_cell[j]];
}

int k_adjust_pc(Handle* h) {
  double* pc = h->fld[CFDL_F_PC];
  if (h->prep.o2c[0] >= 0 && h->prep.o2c[0] < h->N)
    copy_scalar_kernel<<<1, 1, 0, S(h)>>>(h->scal, pc + h->prep.o2c[0]);  // pref = phic(1)
  int rc = comm_bcast(h, h->scal, 1, h->prep.ref_cell_owner);  // several GPUs: the owner of cell 1 provides pref
  if (rc) return rc;
  // ghost copies are shifted by the same subtraction, so they stay equal to their owners' values
  shift_kernel<<<grid_for(h, h->Nc, TPB), TPB, 0, S(h)>>>(h->Nc, pc, h->scal);
  if (h->B) halo_copy_kernel<<<grid_for(h, h->B, TPB), TPB, 0, S(h)>>>(h->B, h->Nc, h->halo_cell, pc);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- update_uvwp, mod_uvwp.f90:370-436 (cell-velocity correction is disabled there) ------------
__global__ void __launch_bounds__(TPB) correct_cells_kernel(int N, int Nc, double* p, const double* __restrict__ pc, double* gp,
                                                            const double* __restrict__ gpc) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < Nc; c += gridDim.x * blockDim.x) {
    p[c] = p[c] + pc[c];  // ghost cells too: their pc is current, so p stays equal to the owner's
    if (c >= N) continue;
    const size_t i = 3 * (size_t)c;
    gp[i] = gp[i] + gpc[i]; gp[i + 1] = gp[i + 1] + gpc[i + 1]; gp[i + 2] = gp[i + 2] + gpc[i + 2];
  }
}

__global__ void __launch_bounds__(TPB) correct_faces_kernel(int Fi, const int32_t* __restrict__ face_a, const int32_t* __restrict__ face_b,
                                                            const double* __restrict__ xc, const double* __restrict__ yc,
                                                            const double* __restrict__ zc, const double* __restrict__ aip,
                                                            const double* __restrict__ rip, const double* __restrict__ rho,
                                                            const double* __restrict__ dc, const double* __restrict__ pc, double* mip) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < Fi; f += gridDim.x * blockDim.x) {
    const int e = face_a[f], nb = face_b[f];
    double a[3], r[3];
    load3(aip, f, a); load3(rip, f, r);
    const double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    const double norm[3] = {a[0] / area, a[1] / area, a[2] / area};
    const double rp[3] = {xc[e], yc[e], zc[e]};
    const double rpnb[3] = {xc[nb], yc[nb], zc[nb]};
    const double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
    const double wt = vec_weight(r, rp, rpnb);
    const double dip = (1.0 - wt) * dc[e] + wt * dc[nb];
    const double rhoip = (rho[e] + rho[nb]) / 2.0;
    const double dmip = rhoip * area * dip * (pc[nb] - pc[e]) / dot3(dr, norm);
    mip[f] = mip[f] - dmip;
  }
}

int k_update_uvwp(Handle* h) {
  correct_cells_kernel<<<grid_for(h, h->Nc, TPB), TPB, 0, S(h)>>>(h->N, h->Nc, h->fld[CFDL_F_P], h->fld[CFDL_F_PC], h->fld[CFDL_F_GP],
                                                                       h->fld[CFDL_F_GPC]);
  if (h->Fi && h->use_statics && h->fs_area) { int rc = k_correct_faces_statics(h); if (rc) return rc; }
  else if (h->Fi)
    correct_faces_kernel<<<grid_for(h, h->Fi, TPB), TPB, 0, S(h)>>>(h->Fi, h->face_a, h->face_b, h->xc, h->yc, h->zc, h->aip,
                                                                          h->rip, h->rho, h->fld[CFDL_F_DC], h->fld[CFDL_F_PC],
                                                                          h->fld[CFDL_F_MIP]);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- calc_grad (weighted least squares) + matinv3, mod_solver.f90:8-81 -------------------------
__device__ __forceinline__ void matinv3_apply(const double A[6], const double g[3], double out[3]) {
  // A = {a11,a12,a13,a22,a23,a33}, symmetric
  const double a11 = A[0], a12 = A[1], a13 = A[2], a21 = A[1], a22 = A[3], a23 = A[4], a31 = A[2], a32 = A[4], a33 = A[5];
  double det = (a11 * a22 * a33 - a11 * a23 * a32 - a12 * a21 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a13 * a22 * a31);
  double b11, b12, b13, b21, b22, b23, b31, b32, b33;
  if (fabs(det) > 2.2250738585072014e-308) {
    const double detinv = 1.0 / det;
    b11 = +detinv * (a22 * a33 - a23 * a32);
    b21 = -detinv * (a21 * a33 - a23 * a31);
    b31 = +detinv * (a21 * a32 - a22 * a31);
    b12 = -detinv * (a12 * a33 - a13 * a32);
    b22 = +detinv * (a11 * a33 - a13 * a31);
    b32 = -detinv * (a11 * a32 - a12 * a31);
    b13 = +detinv * (a12 * a23 - a13 * a22);
    b23 = -detinv * (a11 * a23 - a13 * a21);
    b33 = +detinv * (a11 * a22 - a12 * a21);
  } else {
    const double detinv = 1.0 / (a11 + a22 + a33);
    b11 = detinv; b22 = detinv; b33 = detinv;
    b12 = b13 = b21 = b23 = b31 = b32 = 0.0;
  }
  out[0] = 0.0 + b11 * g[0] + b12 * g[1] + b13 * g[2];
  out[1] = 0.0 + b21 * g[0] + b22 * g[1] + b23 * g[2];
  out[2] = 0.0 + b31 * g[0] + b32 * g[1] + b33 * g[2];
}

// grad_variant 0: the reference's form (weights and the 3x3 inverse rebuilt per call)
template <int K, int NF>
__global__ void __launch_bounds__(TPB) grad_kernel(int N, int Np, const int32_t* __restrict__ ell_nb, const uint8_t* __restrict__ nfc,
                                                   const double* __restrict__ xc, const double* __restrict__ yc,
                                                   const double* __restrict__ zc, const double* phi0, const double* phi1,
                                                   const double* phi2, double* g0, double* g1, double* g2) {
  const double* phis[3] = {phi0, phi1, phi2};
  double* gs[3] = {g0, g1, g2};
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    const int n = nfc[c];
    const double rp[3] = {xc[c], yc[c], zc[c]};
    double pe[NF], g[NF][3], A[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < NF; ++q) { pe[q] = phis[q][c]; g[q][0] = g[q][1] = g[q][2] = 0.0; }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k < n) {
        const int nb = ell_nb[(size_t)k * Np + c];
        const double dr[3] = {xc[nb] - rp[0], yc[nb] - rp[1], zc[nb] - rp[2]};
        const double wt = 1.0 / (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
#pragma unroll
        for (int q = 0; q < NF; ++q) {
          const double dphi = phis[q][nb] - pe[q];
          g[q][0] = g[q][0] + wt * dphi * dr[0];
          g[q][1] = g[q][1] + wt * dphi * dr[1];
          g[q][2] = g[q][2] + wt * dphi * dr[2];
        }
        A[0] = A[0] + wt * dr[0] * dr[0];
        A[1] = A[1] + wt * dr[0] * dr[1];
        A[2] = A[2] + wt * dr[0] * dr[2];
        A[3] = A[3] + wt * dr[1] * dr[1];
        A[4] = A[4] + wt * dr[1] * dr[2];
        A[5] = A[5] + wt * dr[2] * dr[2];
      }
    }
#pragma unroll
    for (int q = 0; q < NF; ++q) {
      double out[3];
      matinv3_apply(A, g[q], out);
      gs[q][3 * (size_t)c] = out[0]; gs[q][3 * (size_t)c + 1] = out[1]; gs[q][3 * (size_t)c + 2] = out[2];
    }
  }
}

// ---- calc_grad on precomputed least-squares statics (grad_variant 1) --------------------------------
// The matrix A = sum w dr dr^T, its inverse (matinv3, with the trace fallback) and the weights
// w = 1/|dr|^2 depend on the mesh only; the reference rebuilds them in each of its four calc_grad calls
// per iteration.  lsq_statics_kernel evaluates the very same expressions once (inverse entries b11..b33
// per cell, w per slot); grad_lsq_kernel then needs neither the division per face nor the inversion and
// produces the same bits (same operands, same operation order) for 15 more doubles read per cell.
template <int K>
__global__ void __launch_bounds__(TPB) lsq_statics_kernel(int N, int Np, const int32_t* __restrict__ ell_nb, const uint8_t* __restrict__ nfc,
                                                          const double* __restrict__ xc, const double* __restrict__ yc,
                                                          const double* __restrict__ zc, double* binv, double* wslot) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    const int n = nfc[c];
    const double rp[3] = {xc[c], yc[c], zc[c]};
    double A[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k < n) {
        const int nb = ell_nb[(size_t)k * Np + c];
        const double dr[3] = {xc[nb] - rp[0], yc[nb] - rp[1], zc[nb] - rp[2]};
        const double wt = 1.0 / (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
        wslot[(size_t)k * Np + c] = wt;
        A[0] = A[0] + wt * dr[0] * dr[0];
        A[1] = A[1] + wt * dr[0] * dr[1];
        A[2] = A[2] + wt * dr[0] * dr[2];
        A[3] = A[3] + wt * dr[1] * dr[1];
        A[4] = A[4] + wt * dr[1] * dr[2];
        A[5] = A[5] + wt * dr[2] * dr[2];
      }
    }
    // matinv3 (mod_solver.f90:8-38): the expressions of matinv3_apply, entries stored instead of applied
    const double a11 = A[0], a12 = A[1], a13 = A[2], a21 = A[1], a22 = A[3], a23 = A[4], a31 = A[2], a32 = A[4], a33 = A[5];
    const double det = (a11 * a22 * a33 - a11 * a23 * a32 - a12 * a21 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a13 * a22 * a31);
    double b[9];
    if (fabs(det) > 2.2250738585072014e-308) {
      const double detinv = 1.0 / det;
      b[0] = +detinv * (a22 * a33 - a23 * a32);
      b[3] = -detinv * (a21 * a33 - a23 * a31);
      b[6] = +detinv * (a21 * a32 - a22 * a31);
      b[1] = -detinv * (a12 * a33 - a13 * a32);
      b[4] = +detinv * (a11 * a33 - a13 * a31);
      b[7] = -detinv * (a11 * a32 - a12 * a31);
      b[2] = +detinv * (a12 * a23 - a13 * a22);
      b[5] = -detinv * (a11 * a23 - a13 * a21);
      b[8] = +detinv * (a11 * a22 - a12 * a21);
    } else {
      const double detinv = 1.0 / (a11 + a22 + a33);
      b[0] = detinv; b[4] = detinv; b[8] = detinv;
      b[1] = b[2] = b[3] = b[5] = b[6] = b[7] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) binv[(size_t)j * Np + c] = b[j];  // row-major b11 b12 b13 b21 ...
  }
}

template <int K, int NF>
__global__ void __launch_bounds__(TPB) grad_lsq_kernel(int N, int Np, const int32_t* __restrict__ ell_nb, const uint8_t* __restrict__ nfc,
                                                       const double* __restrict__ xc, const double* __restrict__ yc,
                                                       const double* __restrict__ zc, const double* __restrict__ binv,
                                                       const double* __restrict__ wslot, const double* phi0, const double* phi1,
                                                       const double* phi2, double* g0, double* g1, double* g2) {
  const double* phis[3] = {phi0, phi1, phi2};
  double* gs[3] = {g0, g1, g2};
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    const int n = nfc[c];
    const double rp[3] = {xc[c], yc[c], zc[c]};
    double pe[NF], g[NF][3];
#pragma unroll
    for (int q = 0; q < NF; ++q) { pe[q] = phis[q][c]; g[q][0] = g[q][1] = g[q][2] = 0.0; }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k < n) {
        const int nb = ell_nb[(size_t)k * Np + c];
        const double dr[3] = {xc[nb] - rp[0], yc[nb] - rp[1], zc[nb] - rp[2]};
        const double wt = wslot[(size_t)k * Np + c];
#pragma unroll
        for (int q = 0; q < NF; ++q) {
          const double dphi = phis[q][nb] - pe[q];
          g[q][0] = g[q][0] + wt * dphi * dr[0];
          g[q][1] = g[q][1] + wt * dphi * dr[1];
          g[q][2] = g[q][2] + wt * dphi * dr[2];
        }
      }
    }
    double b[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) b[j] = binv[(size_t)j * Np + c];
#pragma unroll
    for (int q = 0; q < NF; ++q) {
      gs[q][3 * (size_t)c] = 0.0 + b[0] * g[q][0] + b[1] * g[q][1] + b[2] * g[q][2];
      gs[q][3 * (size_t)c + 1] = 0.0 + b[3] * g[q][0] + b[4] * g[q][1] + b[5] * g[q][2];
      gs[q][3 * (size_t)c + 2] = 0.0 + b[6] * g[q][0] + b[7] * g[q][1] + b[8] * g[q][2];
    }
  }
}

// allocates and fills the least-squares statics on first use
static int ensure_lsq_statics(Handle* h) {
  if (h->lsq_binv) return CFDL_OK;
  double *b = nullptr, *w = nullptr;
  CFDL_CUDA(cudaMalloc(&b, sizeof(double) * 9 * (size_t)h->Np));
  h->allocs.push_back(b);
  CFDL_CUDA(cudaMalloc(&w, sizeof(double) * (size_t)h->K * h->Np));
  h->allocs.push_back(w);
  CFDL_CUDA(cudaMemsetAsync(b, 0, sizeof(double) * 9 * (size_t)h->Np, h->stream));
  CFDL_CUDA(cudaMemsetAsync(w, 0, sizeof(double) * (size_t)h->K * h->Np, h->stream));
  const int g = grid_for(h, h->N, TPB);
  if (h->K <= 4) lsq_statics_kernel<4><<<g, TPB, 0, S(h)>>>(h->N, h->Np, h->ell_nb, h->nfc, h->xc, h->yc, h->zc, b, w);
  else lsq_statics_kernel<6><<<g, TPB, 0, S(h)>>>(h->N, h->Np, h->ell_nb, h->nfc, h->xc, h->yc, h->zc, b, w);
  CFDL_CUDA(cudaGetLastError());
  h->lsq_binv = b; h->lsq_w = w;
  return CFDL_OK;
}

template <auto K4, auto K6, typename... Args>
static void launch_k46(Handle* h, int cells, Args... args) {
  if (h->K <= 4) K4<<<occ_grid<K4>(h, cells, TPB), TPB, 0, S(h)>>>(args...);
  else K6<<<occ_grid<K6>(h, cells, TPB), TPB, 0, S(h)>>>(args...);
}

// grad_variant 1 (default; B200, 128^3: 0.134 against 0.156 ms for u,v,w, 0.090 against 0.108 for one field; the
// locality order measured slower for both) = on the LSQ statics, 0 = the reference's form
static int grad1_launch(Handle* h, const double* phi, double* grad) {
  if (h->grad_variant != 0) {
    int rc = ensure_lsq_statics(h);
    if (rc) return rc;
    launch_k46<grad_lsq_kernel<4, 1>, grad_lsq_kernel<6, 1>>(h, h->N, h->N, h->Np, h->ell_nb, h->nfc, h->xc, h->yc, h->zc, h->lsq_binv, h->lsq_w, phi, phi, phi, grad, grad, grad);
  } else {
    launch_k46<grad_kernel<4, 1>, grad_kernel<6, 1>>(h, h->N, h->N, h->Np, h->ell_nb, h->nfc, h->xc, h->yc, h->zc, phi, phi, phi, grad, grad, grad);
  }
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}
static int grad3_launch(Handle* h) {
  const double *u = h->fld[CFDL_F_U], *v = h->fld[CFDL_F_V], *w = h->fld[CFDL_F_W];
  double *gu = h->fld[CFDL_F_GU], *gv = h->fld[CFDL_F_GV], *gw = h->fld[CFDL_F_GW];
  if (h->grad_variant != 0) {
    int rc = ensure_lsq_statics(h);
    if (rc) return rc;
    launch_k46<grad_lsq_kernel<4, 3>, grad_lsq_kernel<6, 3>>(h, h->N, h->N, h->Np, h->ell_nb, h->nfc, h->xc, h->yc, h->zc, h->lsq_binv, h->lsq_w, u, v, w, gu, gv, gw);
  } else {
    launch_k46<grad_kernel<4, 3>, grad_kernel<6, 3>>(h, h->N, h->N, h->Np, h->ell_nb, h->nfc, h->xc, h->yc, h->zc, u, v, w, gu, gv, gw);
  }
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

int k_calc_grad(Handle* h, const double* phi, double* grad) {
  if (h->K > 6) return fail(CFDL_ERR_UNSUPPORTED, "cells with more than 6 faces are not supported");
  prof_begin(h, PROF_GRAD1);
  int rc = grad1_launch(h, phi, grad);
  prof_end(h);
  return rc;
}

int k_calc_grad3(Handle* h) {  // the three calc_grad calls of mod_uvwp.f90:118-120 in one pass over the mesh
  if (h->K > 6) return fail(CFDL_ERR_UNSUPPORTED, "cells with more than 6 faces are not supported");
  prof_begin(h, PROF_GRAD);
  int rc = grad3_launch(h);
  prof_end(h);
  return rc;
}

int k_update_time(Handle* h) {
  const size_t hb = sizeof(double) * (size_t)h->H;
  CFDL_CUDA(cudaMemcpyAsync(h->fld[CFDL_F_U0], h->fld[CFDL_F_U], hb, cudaMemcpyDeviceToDevice, h->stream));
  CFDL_CUDA(cudaMemcpyAsync(h->fld[CFDL_F_V0], h->fld[CFDL_F_V], hb, cudaMemcpyDeviceToDevice, h->stream));
  CFDL_CUDA(cudaMemcpyAsync(h->fld[CFDL_F_W0], h->fld[CFDL_F_W], hb, cudaMemcpyDeviceToDevice, h->stream));
  CFDL_CUDA(cudaMemcpyAsync(h->fld[CFDL_F_MIP0], h->fld[CFDL_F_MIP], sizeof(double) * (size_t)h->F, cudaMemcpyDeviceToDevice, h->stream));
  if (h->has_energy || h->has_scalar) return k_transport_update_time(h);
  return CFDL_OK;
}

// ---- numbering conversions for host transfers ---------------------------------------------------
// dst[i*ncomp + q] = src[map[i]*ncomp + q]
__global__ void __launch_bounds__(TPB) gather_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                                     const int32_t* __restrict__ map, int64_t n, int ncomp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    for (int q = 0; q < ncomp; ++q) dst[i * ncomp + q] = src[(int64_t)map[i] * ncomp + q];
}
__global__ void __launch_bounds__(TPB) scatter_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                                      const int32_t* __restrict__ map, int64_t n, int ncomp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    for (int q = 0; q < ncomp; ++q) dst[(int64_t)map[i] * ncomp + q] = src[i * ncomp + q];
}
int k_gather(Handle* h, double* dst, const double* src, const int32_t* map, int64_t n, int ncomp) {
  if (n == 0) return CFDL_OK;
  gather_kernel<<<grid_for(h, n, TPB), TPB, 0, S(h)>>>(dst, src, map, n, ncomp);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}
int k_scatter(Handle* h, double* dst, const double* src, const int32_t* map, int64_t n, int ncomp) {
  if (n == 0) return CFDL_OK;
  scatter_kernel<<<grid_for(h, n, TPB), TPB, 0, S(h)>>>(dst, src, map, n, ncomp);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// anb between the reference's CSR slot order (ef2nb order, original cell numbering) and device ELL
__global__ void __launch_bounds__(TPB) csr_ell_kernel(int N, int Np, const int32_t* __restrict__ c2o, const int32_t* __restrict__ row_ptr,
                                                      double* ell, double* csr, int to_ell) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    const int e = c2o[c], r0 = row_ptr[e], n = row_ptr[e + 1] - r0;
    for (int k = 0; k < n; ++k) {
      if (to_ell) ell[(size_t)k * Np + c] = csr[r0 + k];
      else csr[r0 + k] = ell[(size_t)k * Np + c];
    }
  }
}
int k_csr_to_ell(Handle* h, double* ell, const double* csr) {
  csr_ell_kernel<<<grid_for(h, h->N, TPB), TPB, 0, S(h)>>>(h->N, h->Np, h->c2o, h->row_ptr, ell, const_cast<double*>(csr), 1);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}
int k_ell_to_csr(Handle* h, double* csr, const double* ell) {
  csr_ell_kernel<<<grid_for(h, h->N, TPB), TPB, 0, S(h)>>>(h->N, h->Np, h->c2o, h->row_ptr, const_cast<double*>(ell), csr, 0);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

}  // namespace cfdl
