// Assembly kernels on precomputed face statics (sm_100a).
//
// The reference recomputes, for every face of every cell in every iteration, quantities that
// depend on the mesh only: |A|, the unit normal, |dr|, the distance weights, the projected
// points and |dr_p| (src/equations/mod_uvwp.f90:188-201,321-326,401-407,461-464).  On the GPU
// those 5 square roots and 4 divisions per face made calc_coef_uvw instruction bound (ncu: XU pipe
// saturated, DRAM at 11 %).  face_statics_kernel evaluates exactly those expressions ONCE, in the
// owner's orientation; seen from the other cell every vector is the exact negation and every
// scalar identical (IEEE negation is exact), so the kernels below produce the same bits as the
// recomputing kernels in kernels_assembly.cu — the tests check that with array_equal.
#include <algorithm>
#include "state.h"
#include "device_math.cuh"

namespace cfdl {

#define TPB 256

struct FaceStatics {
  const double *area, *ds, *dsp, *dn, *wto, *wtn;
  const double *rds, *rdsp;  // RN(1/ds), RN(1/dsp): the reciprocals quot<true> needs (uvw_variant 5, 6)
  const double* rdn;         // RN(1/dn) for the quotients by dr.n in calc_mip, calc_coef_p and the face correction
  const double *n[3], *dr[3], *drp[3];
};

static FaceStatics statics_of(const Handle* h) {
  FaceStatics S;
  S.area = h->fs_area; S.ds = h->fs_ds; S.dsp = h->fs_dsp; S.dn = h->fs_dn; S.wto = h->fs_wto; S.wtn = h->fs_wtn;
  S.rds = h->fs_rds; S.rdsp = h->fs_rdsp; S.rdn = h->fs_rdn;
  for (int i = 0; i < 3; ++i) { S.n[i] = h->fs_n[i]; S.dr[i] = h->fs_dr[i]; S.drp[i] = h->fs_drp[i]; }
  return S;
}

__global__ void __launch_bounds__(TPB) face_statics_kernel(int Fi, const int32_t* __restrict__ face_a, const int32_t* __restrict__ face_b,
                                                           const double* __restrict__ xc, const double* __restrict__ yc,
                                                           const double* __restrict__ zc, const double* __restrict__ aip,
                                                           const double* __restrict__ rip_, double* area_o, double* ds_o, double* dsp_o,
                                                           double* dn_o, double* wto_o, double* wtn_o, double* n0, double* n1, double* n2,
                                                           double* d0, double* d1, double* d2, double* p0, double* p1, double* p2,
                                                           double* rds_o, double* rdsp_o, double* rdn_o) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < Fi; f += gridDim.x * blockDim.x) {
    const int e = face_a[f], nb = face_b[f];
    const double rp[3] = {xc[e], yc[e], zc[e]};
    const double rpnb[3] = {xc[nb], yc[nb], zc[nb]};
    double a[3], rip[3];
    load3(aip, f, a); load3(rip_, f, rip);
    const double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    const double norm[3] = {a[0] / area, a[1] / area, a[2] / area};
    const double dr[3] = {rpnb[0] - rp[0], rpnb[1] - rp[1], rpnb[2] - rp[2]};
    const double ds = sqrt(dot3(dr, dr));
    double drip[3] = {rip[0] - rp[0], rip[1] - rp[1], rip[2] - rp[2]};
    double t = dot3(drip, norm);
    const double rp_p[3] = {rip[0] - t * norm[0], rip[1] - t * norm[1], rip[2] - t * norm[2]};
    drip[0] = rip[0] - rpnb[0]; drip[1] = rip[1] - rpnb[1]; drip[2] = rip[2] - rpnb[2];
    t = dot3(drip, norm);
    const double rpnb_p[3] = {rip[0] - t * norm[0], rip[1] - t * norm[1], rip[2] - t * norm[2]};
    const double dr_p[3] = {rpnb_p[0] - rp_p[0], rpnb_p[1] - rp_p[1], rpnb_p[2] - rp_p[2]};
    const double dsp = sqrt(dot3(dr_p, dr_p));
    const double dn = dot3(dr, norm);
    area_o[f] = area; ds_o[f] = ds; dsp_o[f] = dsp; dn_o[f] = dn;
    rds_o[f] = 1.0 / ds; rdsp_o[f] = 1.0 / dsp; rdn_o[f] = 1.0 / dn;
    wto_o[f] = vec_weight(rip, rp, rpnb);   // weight seen from the owner
    wtn_o[f] = vec_weight(rip, rpnb, rp);   // weight seen from the neighbour
    n0[f] = norm[0]; n1[f] = norm[1]; n2[f] = norm[2];
    d0[f] = dr[0]; d1[f] = dr[1]; d2[f] = dr[2];
    p0[f] = dr_p[0]; p1[f] = dr_p[1]; p2[f] = dr_p[2];
  }
}

int k_face_statics(Handle* h) {
  if (h->Fi == 0) return CFDL_OK;
  face_statics_kernel<<<grid_for(h, h->Fi, TPB), TPB, 0, S(h)>>>(h->Fi, h->face_a, h->face_b, h->xc, h->yc, h->zc, h->aip, h->rip, h->fs_area,
                                                                 h->fs_ds, h->fs_dsp, h->fs_dn, h->fs_wto, h->fs_wtn, h->fs_n[0], h->fs_n[1],
                                                                 h->fs_n[2], h->fs_dr[0], h->fs_dr[1], h->fs_dr[2], h->fs_drp[0], h->fs_drp[1],
                                                                 h->fs_drp[2], h->fs_rds, h->fs_rdsp, h->fs_rdn);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- calc_coef_uvw on statics (mod_uvwp.f90:161-286) ------------------------------------------
struct UvwArgsS {
  int N, Nc, Np, ncol0;  // ncol0: cells of the first colour (paired order)
  const int32_t* order;  // owned cells in spatial order (locality-order variants)
  const int32_t *ell_nb, *ell_fs, *halo_bc, *bc_kind;
  const uint8_t* nfc;
  const double *xc, *yc, *zc, *aip, *vol, *rho, *mu;
  const double *u, *v, *w, *u0, *v0, *w0, *gu, *gv, *gw, *gp, *mip;
  double *ap, *anb, *bu, *bv, *bw, *d, *dc;
  double dt;
  FaceStatics S;
};

// The ten quotients per face share two divisors (|dr| and |dr_p|); their reciprocals come from the statics
// and every quotient is formed with quot<true> (one multiply + two FMAs: the correctly rounded quotient,
// device_math.cuh) — the bits of the reference's divisions for a third of their instructions.
// Round 2 timed fourteen forms of this routine on a B200 (profiles/r02_call1_bench_default_autotune.json:
// divisions 0.89 ms, in-kernel reciprocals 0.53, stored reciprocals 0.49, paired colour order 0.47, lean
// register variants 0.48-0.56, forced occupancy 0.53-0.59, locality order 0.46); the winner is what is left.
// A slot-parallel form on the statics (one thread per (cell, face slot), terms handed over in shared memory, one
// thread per cell summing in slot order; 56-80 registers, 37-56 % occupancy) was measured too: 0.91-1.16 ms — more
// than half of its stall cycles sit at the hand-over barrier — and was dropped.  So was a form that staged the ~23 operands
// of a face through per-thread shared-memory slots with cp.async, two faces in flight and no barrier (138 LDGSTS, 118
// registers, four CTAs of 128 threads per SM): 0.626 ms against 0.444 — the long-scoreboard stalls fell from 12.6 to 8.2
// cycles per issue, but the 8-byte asynchronous copies saturate the LSU / shared-memory path (profiles/r02_summary.md).
template <int K>
__device__ __forceinline__ void coef_uvw_statics_cell(const UvwArgsS& A, const int c) {
  const int Nc = A.Nc, Np = A.Np;
  const int n = A.nfc[c];
  const double mu_e = A.mu[c];
  double gue[3], gve[3], gwe[3];
  load3(A.gu, c, gue); load3(A.gv, c, gve); load3(A.gw, c, gwe);
  double ap = 0.0, sumf = 0.0, sumss[3] = {0, 0, 0}, sumdefc[3] = {0, 0, 0};
  double anbk[K];
  int nbk[K], fsk[K];
  // connectivity of all slots first (the ELL arrays are padded, so every slot of every cell can be read): the loads are
  // independent of one another and of nfc, so the kernel pays one memory latency for them instead of one per face in
  // front of the face's data loads (ncu, round 2: 77 % of the stall cycles were long-scoreboard waits on that chain)
#pragma unroll
  for (int k = 0; k < K; ++k) {
    nbk[k] = __ldg(&A.ell_nb[(size_t)k * Np + c]);
    fsk[k] = __ldg(&A.ell_fs[(size_t)k * Np + c]);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    anbk[k] = 0.0;
    if (k >= n) nbk[k] = -1;
    if (k < n) {
      const int nb = nbk[k];
      const int fs = fsk[k];
      double d = 0.0, fnb = 0.0;
      if (nb < Nc) {
        const int f = abs(fs) - 1;
        const bool own = fs > 0;
        const double sg = own ? 1.0 : -1.0;
        const double area = A.S.area[f], ds = A.S.ds[f], ds_p = A.S.dsp[f];
        const double wt = own ? A.S.wto[f] : A.S.wtn[f];
        const double dr[3] = {sg * A.S.dr[0][f], sg * A.S.dr[1][f], sg * A.S.dr[2][f]};
        const double dr_p[3] = {sg * A.S.drp[0][f], sg * A.S.drp[1][f], sg * A.S.drp[2][f]};
        const double f_in = -sg * A.mip[f];
        fnb = fmax(f_in, 0.0);
        sumf = sumf + f_in;
        const double muip = (1.0 - wt) * mu_e + wt * A.mu[nb];
        const double rds = A.S.rds[f], rdsp = A.S.rdsp[f];
        d = quot<true>(muip * area, ds, rds);
        double gun[3], gvn[3], gwn[3];
        load3(A.gu, nb, gun); load3(A.gv, nb, gvn); load3(A.gw, nb, gwn);
        const double w1 = 1.0 - wt;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const double gip[3] = {w1 * gue[m] + wt * gun[m], w1 * gve[m] + wt * gvn[m], w1 * gwe[m] + wt * gwn[m]};
          sumss[m] = sumss[m] + quot<true>(muip * area * dot3(gip, dr), ds, rds);
        }
        {
          double gip[3] = {w1 * gue[0] + wt * gun[0], w1 * gue[1] + wt * gun[1], w1 * gue[2] + wt * gun[2]};
          sumdefc[0] = sumdefc[0] + muip * area * (quot<true>(dot3(gip, dr_p), ds_p, rdsp) - quot<true>(dot3(gip, dr), ds, rds));
          gip[0] = w1 * gve[0] + wt * gvn[0]; gip[1] = w1 * gve[1] + wt * gvn[1]; gip[2] = w1 * gve[2] + wt * gvn[2];
          sumdefc[1] = sumdefc[1] + muip * area * (quot<true>(dot3(gip, dr_p), ds_p, rdsp) - quot<true>(dot3(gip, dr), ds, rds));
          gip[0] = w1 * gwe[0] + wt * gwn[0]; gip[1] = w1 * gwe[1] + wt * gwn[1]; gip[2] = w1 * gwe[2] + wt * gwn[2];
          sumdefc[2] = sumdefc[2] + muip * area * (quot<true>(dot3(gip, dr_p), ds_p, rdsp) - quot<true>(dot3(gip, dr), ds, rds));
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
  int last = -1;  // boundary faces in halo order (rare: geometry evaluated on the fly as in the reference)
  for (int t = 0; t < K; ++t) {
    int best = 0x7fffffff, bk = -1;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (nbk[k] >= Nc && nbk[k] > last && nbk[k] < best) { best = nbk[k]; bk = k; }
    if (bk < 0) break;
    last = best;
    const int bc = A.halo_bc[best - Nc];
    if (bc < 0) continue;
    const int f = A.ell_fs[(size_t)bk * Np + c] - 1;
    double a[3];
    load3(A.aip, f, a);
    const double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    const double norm[3] = {a[0] / area, a[1] / area, a[2] / area};
    const double dr[3] = {A.xc[best] - A.xc[c], A.yc[best] - A.yc[c], A.zc[best] - A.zc[c]};
    const double ds = sqrt(dot3(dr, dr));
    const double d = mu_e * area / ds;
    if (A.bc_kind[bc] != CFDL_BC_SYMMETRY) {
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
  double dcv = ap;
#pragma unroll
  for (int k = 0; k < K; ++k)
    if (k < n) { dcv = dcv - anbk[k]; A.anb[(size_t)k * Np + c] = anbk[k]; }
  A.ap[c] = ap;
  A.bu[c] = bu; A.bv[c] = bv; A.bw[c] = bw;
  A.d[c] = vol / ap;
  A.dc[c] = vol / dcv;
}

// Locality order (any number of colours): thread i takes the i-th cell of the base (natural | Morton) order,
// so the lanes of a warp hold cells of all colours that are neighbours in space — the two cells of a face read
// its statics in the same instruction or a few instructions apart (L1), where a colour-major sweep reads them
// a second time from DRAM, half a kernel later.  Per-cell arrays are touched as ncolors contiguous runs per warp.
template <int K>
__global__ void __launch_bounds__(TPB) coef_uvw_statics_kernel(const UvwArgsS A) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.N; i += gridDim.x * blockDim.x) coef_uvw_statics_cell<K>(A, A.order ? A.order[i] : i);
}

int k_calc_coef_uvw_statics(Handle* h, double dt) {
  UvwArgsS A;
  A.N = h->N; A.Nc = h->Nc; A.Np = h->Np;
  A.ell_nb = h->ell_nb; A.ell_fs = h->ell_fs; A.halo_bc = h->halo_bc; A.bc_kind = h->bc_kind; A.nfc = h->nfc;
  A.xc = h->xc; A.yc = h->yc; A.zc = h->zc; A.aip = h->aip; A.vol = h->vol; A.rho = h->rho; A.mu = h->mu;
  A.u = h->fld[CFDL_F_U]; A.v = h->fld[CFDL_F_V]; A.w = h->fld[CFDL_F_W];
  A.u0 = h->fld[CFDL_F_U0]; A.v0 = h->fld[CFDL_F_V0]; A.w0 = h->fld[CFDL_F_W0];
  A.gu = h->fld[CFDL_F_GU]; A.gv = h->fld[CFDL_F_GV]; A.gw = h->fld[CFDL_F_GW]; A.gp = h->fld[CFDL_F_GP];
  A.mip = h->fld[CFDL_F_MIP];
  A.ap = h->fld[CFDL_F_AP]; A.anb = h->fld[CFDL_F_ANB]; A.bu = h->fld[CFDL_F_BU]; A.bv = h->fld[CFDL_F_BV];
  A.bw = h->fld[CFDL_F_BW]; A.d = h->fld[CFDL_F_D]; A.dc = h->fld[CFDL_F_DC];
  A.dt = dt;
  A.S = statics_of(h);
  A.ncol0 = h->prep.ncolors == 2 ? h->prep.color_ptr[1] : h->N;
  A.order = h->loc_order;
  if (h->K <= 4) coef_uvw_statics_kernel<4><<<occ_grid<coef_uvw_statics_kernel<4>>(h, h->N, TPB), TPB, 0, S(h)>>>(A);
  else coef_uvw_statics_kernel<6><<<occ_grid<coef_uvw_statics_kernel<6>>(h, h->N, TPB), TPB, 0, S(h)>>>(A);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- calc_coef_p on statics (mod_uvwp.f90:289-368) ---------------------------------------------
struct CoefPArgs {
  int N, Nc, Np, ncol0;
  const int32_t* order;
  const int32_t *ell_nb, *ell_fs, *halo_bc;
  const uint8_t* nfc;
  const double *rho, *dc, *mip;
  double *ap, *anb, *b;
  FaceStatics S;
};

// the quotient by dr.n through the stored reciprocal and quot<true> (same bits as the division)
template <int K>
__device__ __forceinline__ void coef_p_statics_cell(const CoefPArgs& A, const int c) {
  const int Nc = A.Nc, Np = A.Np;
  const int n = A.nfc[c];
  const double rho_e = A.rho[c], dc_e = A.dc[c];
  double ap = 0.0, sumf = 0.0;
  int nbk[K], fsk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {  // connectivity of all slots first: independent loads, one latency (see coef_uvw_statics_cell)
    nbk[k] = __ldg(&A.ell_nb[(size_t)k * Np + c]);
    fsk[k] = __ldg(&A.ell_fs[(size_t)k * Np + c]);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (k >= n) nbk[k] = -1;
    if (k < n) {
      const int nb = nbk[k];
      const int fs = fsk[k];
      double d = 0.0;
      if (nb < Nc) {
        const int f = abs(fs) - 1;
        const bool own = fs > 0;
        const double sg = own ? 1.0 : -1.0;
        const double wt = own ? A.S.wto[f] : A.S.wtn[f];
        const double f_in = -sg * A.mip[f];
        sumf = sumf + f_in;
        const double rhoip = (1.0 - wt) * rho_e + wt * A.rho[nb];
        d = quot<true>((1.0 - wt) * dc_e + wt * A.dc[nb], A.S.dn[f], A.S.rdn[f]) * rhoip * A.S.area[f];
      }
      A.anb[(size_t)k * Np + c] = d;
      ap = ap + d;
    }
  }
  double b = sumf;
  int last = -1;
  for (int t = 0; t < K; ++t) {
    int best = 0x7fffffff, bk = -1;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (nbk[k] >= Nc && nbk[k] > last && nbk[k] < best) { best = nbk[k]; bk = k; }
    if (bk < 0) break;
    last = best;
    if (A.halo_bc[best - Nc] < 0) continue;
    b = b - A.mip[A.ell_fs[(size_t)bk * Np + c] - 1];
  }
  A.ap[c] = ap;
  A.b[c] = b;
}

// two-colour meshes: the paired colour order — a CTA takes TPB cells of the first colour and then the TPB cells
// at the same position of the second colour, on a mesh numbered with locality their neighbours, so the statics
// and mip of a face are read by both of its cells within one CTA's pass instead of half a kernel apart
// (B200, 128^3: 0.158 ms against 0.183 linear and 0.175 in the locality order)
template <int K>
__global__ void __launch_bounds__(TPB) coef_p_statics_paired_kernel(const CoefPArgs A) {
  const int n0 = A.ncol0, n1 = A.N - A.ncol0;
  const int nq = (max(n0, n1) + (int)blockDim.x - 1) / (int)blockDim.x;
  for (int q = blockIdx.x; q < nq; q += gridDim.x) {
    const int i = q * blockDim.x + threadIdx.x;
#pragma unroll 1
    for (int col = 0; col < 2; ++col)
      if (i < (col ? n1 : n0)) coef_p_statics_cell<K>(A, col ? n0 + i : i);
  }
}
// more colours: locality order (see coef_uvw_statics_kernel)
template <int K>
__global__ void __launch_bounds__(TPB) coef_p_statics_loc_kernel(const CoefPArgs A) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.N; i += gridDim.x * blockDim.x) coef_p_statics_cell<K>(A, A.order ? A.order[i] : i);
}

template <auto K4, auto K6>
static void launch_coef_p(Handle* h, const CoefPArgs& A, int cells) {
  if (h->K <= 4) K4<<<occ_grid<K4>(h, cells, TPB), TPB, 0, S(h)>>>(A);
  else K6<<<occ_grid<K6>(h, cells, TPB), TPB, 0, S(h)>>>(A);
}

int k_calc_coef_p_statics(Handle* h) {
  CoefPArgs A;
  A.N = h->N; A.Nc = h->Nc; A.Np = h->Np; A.ncol0 = h->prep.ncolors == 2 ? h->prep.color_ptr[1] : h->N;
  A.order = h->loc_order;
  A.ell_nb = h->ell_nb; A.ell_fs = h->ell_fs; A.halo_bc = h->halo_bc; A.nfc = h->nfc;
  A.rho = h->rho; A.dc = h->fld[CFDL_F_DC]; A.mip = h->fld[CFDL_F_MIP];
  A.ap = h->fld[CFDL_F_AP]; A.anb = h->fld[CFDL_F_ANB]; A.b = h->fld[CFDL_F_B];
  A.S = statics_of(h);
  if (h->prep.ncolors == 2) launch_coef_p<coef_p_statics_paired_kernel<4>, coef_p_statics_paired_kernel<6>>(h, A, std::max(A.ncol0, h->N - A.ncol0));
  else launch_coef_p<coef_p_statics_loc_kernel<4>, coef_p_statics_loc_kernel<6>>(h, A, h->N);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- calc_mip on statics (mod_uvwp.f90:438-490) ------------------------------------------------
// Faces visited from the cells: a thread takes one cell and the cell-cell faces that are
// numbered from it (ftouch; on a two-colour mesh every face of a first-colour cell).  Faces are
// numbered slot-major over exactly these cells, so for a fixed slot consecutive threads read
// consecutive statics and write consecutive mip entries, the cell's own fields are read once,
// and the neighbours' fields come from L2 instead of one DRAM pass per face slot (a thread per face
// streamed 2.5x the algorithmic bytes, round-1 ncu).  Each face evaluates the reference's expression on
// (owner, neighbour) in the reference's orientation, hence the same bits.
struct MipCellArgs {
  int n_cells, Np, K;
  const int32_t *ell_nb, *ell_fs;
  const uint8_t* ftouch;
  const double* rho;
  const double *u, *v, *w, *u0, *v0, *w0, *p, *gp, *d, *mip0;
  double* mip;
  double dt;
  int rhie_chow;
  FaceStatics S;
};

struct MipCellVals { double u, v, w, u0, v0, w0, p, d, rho, g[3]; };

__device__ __forceinline__ void mip_load_cell(const MipCellArgs& A, int c, bool rc, MipCellVals& x) {
  x.u = A.u[c]; x.v = A.v[c]; x.w = A.w[c]; x.rho = A.rho[c];
  if (rc) {
    x.u0 = A.u0[c]; x.v0 = A.v0[c]; x.w0 = A.w0[c]; x.p = A.p[c]; x.d = A.d[c];
    load3(A.gp, c, x.g);
  }
}

// both quotients of the Rhie-Chow term (by dr.n and by dt) through reciprocals and quot<true>
// HOIST: the connectivity of all slots is loaded before the faces are visited (one memory latency instead of one per face
// in front of the face's data; 12 more live registers, so two instead of three CTAs per SM)
template <int K, bool HOIST>
__global__ void __launch_bounds__(TPB, HOIST ? 2 : 3) mip_cells_kernel(const MipCellArgs A) {
  const double rdt = 1.0 / A.dt;
  const bool rc = A.rhie_chow != 0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < A.n_cells; c += gridDim.x * blockDim.x) {
    const unsigned mask = A.ftouch[c];
    if (!mask) continue;
    MipCellVals me;
    mip_load_cell(A, c, rc, me);
    int nbk[K], fsk[K];
    if (HOIST) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        nbk[k] = __ldg(&A.ell_nb[(size_t)k * A.Np + c]);
        fsk[k] = __ldg(&A.ell_fs[(size_t)k * A.Np + c]);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!(mask >> k & 1u)) continue;
      const int other = HOIST ? nbk[k] : A.ell_nb[(size_t)k * A.Np + c];
      const int fs = HOIST ? fsk[k] : A.ell_fs[(size_t)k * A.Np + c];
      const int f = abs(fs) - 1;
      MipCellVals ot;
      mip_load_cell(A, other, rc, ot);
      const bool own = fs > 0;  // this cell is the face's owner (reference orientation e -> nb)
      const MipCellVals& E = own ? me : ot;
      const MipCellVals& NB = own ? ot : me;
      const double area = A.S.area[f];
      const double norm[3] = {A.S.n[0][f], A.S.n[1][f], A.S.n[2][f]};
      const double wt = A.S.wto[f];
      const double w1 = 1.0 - wt;
      const double velip[3] = {w1 * E.u + wt * NB.u, w1 * E.v + wt * NB.v, w1 * E.w + wt * NB.w};
      const double rhoip = E.rho * w1 + NB.rho * wt;
      double m = dot3(velip, norm) * rhoip * area;
      if (rc) {
        const double dr[3] = {A.S.dr[0][f], A.S.dr[1][f], A.S.dr[2][f]};
        const double gpip[3] = {w1 * E.g[0] + wt * NB.g[0], w1 * E.g[1] + wt * NB.g[1], w1 * E.g[2] + wt * NB.g[2]};
        const double dip = w1 * E.d + wt * NB.d;
        const double velip0[3] = {w1 * E.u0 + wt * NB.u0, w1 * E.v0 + wt * NB.v0, w1 * E.w0 + wt * NB.w0};
        m = m - quot<true>(rhoip * area * dip, A.S.dn[f], A.S.rdn[f]) * (NB.p - E.p - dot3(gpip, dr))
              - quot<true>(rhoip, A.dt, rdt) * dip * (A.mip0[f] - dot3(velip0, norm) * rhoip * area);
      }
      A.mip[f] = m;
    }
  }
}

template <auto K4, auto K6>
static void launch_mip(Handle* h, const MipCellArgs& A) {
  if (h->K <= 4) K4<<<occ_grid<K4>(h, A.n_cells, TPB), TPB, 0, S(h)>>>(A);
  else K6<<<occ_grid<K6>(h, A.n_cells, TPB), TPB, 0, S(h)>>>(A);
}

int k_calc_mip_statics(Handle* h, bool rhie_chow, double dt) {
  if (h->Fi == 0) return CFDL_OK;
  MipCellArgs A;
  A.n_cells = h->prep.touch_end; A.Np = h->Np; A.K = h->K; A.ell_nb = h->ell_nb; A.ell_fs = h->ell_fs; A.ftouch = h->ftouch; A.rho = h->rho;
  A.u = h->fld[CFDL_F_U]; A.v = h->fld[CFDL_F_V]; A.w = h->fld[CFDL_F_W];
  A.u0 = h->fld[CFDL_F_U0]; A.v0 = h->fld[CFDL_F_V0]; A.w0 = h->fld[CFDL_F_W0];
  A.p = h->fld[CFDL_F_P]; A.gp = h->fld[CFDL_F_GP]; A.d = h->fld[CFDL_F_D]; A.mip0 = h->fld[CFDL_F_MIP0];
  A.mip = h->fld[CFDL_F_MIP]; A.dt = dt; A.rhie_chow = rhie_chow ? 1 : 0;
  A.S = statics_of(h);
  if (h->mip_hoist) launch_mip<mip_cells_kernel<4, true>, mip_cells_kernel<6, true>>(h, A);
  else launch_mip<mip_cells_kernel<4, false>, mip_cells_kernel<6, false>>(h, A);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- face part of update_uvwp on statics (mod_uvwp.f90:394-415) ----------------------------------
// (the quotient by dr.n through the stored reciprocal: quot<true>, same bits as the division)
__global__ void __launch_bounds__(TPB) correct_faces_statics_kernel(int Fi, const int32_t* __restrict__ face_a, const int32_t* __restrict__ face_b,
                                                                    const double* __restrict__ rho, const double* __restrict__ dc,
                                                                    const double* __restrict__ pc, const FaceStatics S, double* mip) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < Fi; f += gridDim.x * blockDim.x) {
    const int e = face_a[f], nb = face_b[f];
    const double wt = S.wto[f];
    const double dip = (1.0 - wt) * dc[e] + wt * dc[nb];
    const double rhoip = (rho[e] + rho[nb]) / 2.0;
    const double dmip = quot<true>(rhoip * S.area[f] * dip * (pc[nb] - pc[e]), S.dn[f], S.rdn[f]);
    mip[f] = mip[f] - dmip;
  }
}

int k_correct_faces_statics(Handle* h) {
  if (h->Fi == 0) return CFDL_OK;
  correct_faces_statics_kernel<<<occ_grid<correct_faces_statics_kernel>(h, h->Fi, TPB), TPB, 0, S(h)>>>(h->Fi, h->face_a, h->face_b, h->rho, h->fld[CFDL_F_DC],
                                                                                                  h->fld[CFDL_F_PC], statics_of(h), h->fld[CFDL_F_MIP]);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

}  // namespace cfdl
