// The energy (enthalpy) and passive-scalar equations of the reference next to uvwp (sm_100a): SURVEY 8(f3).
//
//   solve_energy      src/equations/mod_energy.f90:59-80    two gradients, calc_coef_energy (:82-169), solve_gs('e'),
//                                                           calc_temperature (mod_properties.f90:214-222)
//   solve_scalar      src/equations/mod_scalar.f90:56-72    calc_coef_scalar (:74-126), gradient, solve_gs('scalar')
//   their boundary callbacks (lid / dirichlet0 of mod_energy.f90:173-212, dirichlet0 / dirichlet1 of
//   mod_scalar.f90:129-155) run with update_boundaries, their phi0 = phi with update_time (mod_physics.f90:104,110).
//
// The reference constructs the energy equation and never solves it (main.f90:59 is commented out) and never constructs
// the scalar one; both reuse everything the uvwp path has — gather-type cell kernels over the ELL slots in the reference's
// face order (bit-identical sums), the face statics, calc_grad, and the solvers (solve_equation: exact natural-order
// SGS in parity mode, multicolour SGS otherwise).  Partitioned handles: tc / cp carry the ghost cells, the gradients' and the
// solved fields' ghosts are exchanged like those of uvwp; the fields live outside the peer-to-peer slab, so the solve itself
// runs on a slab-resident work array (the side-by-side momentum solve's, idle here) and the result is copied back.
#include "state.h"
#include <vector>
#include "device_math.cuh"

namespace cfdl {

#define TPB 256

// lid (373 K on a CFDL_BC_LID section) / dirichlet0 (273 K elsewhere): t(halo) and phi(halo) = cp(cell) t
__global__ void __launch_bounds__(TPB) energy_bc_kernel(int B, int Nc, const int32_t* __restrict__ halo_cell, const int32_t* __restrict__ halo_bc,
                                                        const int32_t* __restrict__ bc_kind, const double* __restrict__ cp, double* t, double* hh) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < B; j += gridDim.x * blockDim.x) {
    const int bc = halo_bc[j];
    if (bc < 0) continue;
    const double T = bc_kind[bc] == CFDL_BC_LID ? 373.0 : 273.0;
    t[Nc + j] = T;
    hh[Nc + j] = cp[halo_cell[j]] * T;
  }
}

__global__ void __launch_bounds__(TPB) scalar_bc_kernel(int B, int Nc, const int32_t* __restrict__ halo_bc, const double* __restrict__ value, double* s) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < B; j += gridDim.x * blockDim.x) {
    const int bc = halo_bc[j];
    if (bc >= 0) s[Nc + j] = value[bc];
  }
}

struct EnergyArgs {
  int N, Nc, Np;
  const int32_t *ell_nb, *ell_fs, *halo_bc;
  const uint8_t* nfc;
  const double *xc, *yc, *zc, *aip, *vol, *rho, *tc, *cp, *mip, *hh, *h0, *gh, *gt;
  const double *fs_area, *fs_dn, *fs_wto, *fs_wtn, *fs_n[3];
  double *ap, *anb, *b;
  double dt;
};

// calc_coef_energy, mod_energy.f90:82-169.  Cell-cell faces from the statics (area, unit normal and dr.n in the owner's
// orientation: seen from the other cell the vectors are the exact negations, dr.n the same value); boundary faces in halo
// order with the geometry evaluated as the reference does.
template <int K>
__global__ void __launch_bounds__(TPB) coef_energy_kernel(const EnergyArgs A) {
  const int Nc = A.Nc, Np = A.Np;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < A.N; c += gridDim.x * blockDim.x) {
    const int n = A.nfc[c];
    const double tc_e = A.tc[c], cp_e = A.cp[c];
    double ghe[3], gte[3];
    load3(A.gh, c, ghe); load3(A.gt, c, gte);
    double ap = 0.0, sumf = 0.0, sumdefc = 0.0;
    double anbk[K];
    int nbk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      anbk[k] = 0.0; nbk[k] = -1;
      if (k < n) {
        const int nb = A.ell_nb[(size_t)k * Np + c];
        nbk[k] = nb;
        if (nb < Nc) {
          const int fs = A.ell_fs[(size_t)k * Np + c];
          const int f = abs(fs) - 1;
          const bool own = fs > 0;
          const double sg = own ? 1.0 : -1.0;
          const double area = A.fs_area[f];
          const double norm[3] = {sg * A.fs_n[0][f], sg * A.fs_n[1][f], sg * A.fs_n[2][f]};
          const double wt = own ? A.fs_wto[f] : A.fs_wtn[f];
          const double f_in = -sg * A.mip[f];
          const double fnb = fmax(f_in, 0.0);
          sumf = sumf + f_in;
          const double tci = (1.0 - wt) * tc_e + wt * A.tc[nb];
          const double cpi = (1.0 - wt) * cp_e + wt * A.cp[nb];
          const double d = tci / cpi / A.fs_dn[f] * area;
          double ghn[3], gtn[3];
          load3(A.gh, nb, ghn); load3(A.gt, nb, gtn);
          const double w1 = 1.0 - wt;
          const double ghi[3] = {w1 * ghe[0] + wt * ghn[0], w1 * ghe[1] + wt * ghn[1], w1 * ghe[2] + wt * ghn[2]};
          const double gti[3] = {w1 * gte[0] + wt * gtn[0], w1 * gte[1] + wt * gtn[1], w1 * gte[2] + wt * gtn[2]};
          sumdefc = sumdefc + tci * area * (dot3(gti, norm) - dot3(ghi, norm) / cpi);
          anbk[k] = d + fnb;
          ap = ap + d + fnb;
        }
      }
    }
    const double vol = A.vol[c];
    const double ap0 = A.rho[c] * vol / A.dt;
    ap = ap + ap0;
    const double b = ap0 * A.h0[c] + sumf * A.hh[c] + sumdefc;
    int last = -1;  // boundary faces of the cell in halo order (the order of the reference's loop over sections and halos)
    for (int t = 0; t < K; ++t) {
      int best = 0x7fffffff, bk = -1;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (nbk[k] >= Nc && nbk[k] > last && nbk[k] < best) { best = nbk[k]; bk = k; }
      if (bk < 0) break;
      last = best;
      if (A.halo_bc[best - Nc] < 0) continue;
      const int f = A.ell_fs[(size_t)bk * Np + c] - 1;
      double a[3];
      load3(A.aip, f, a);
      const double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      const double norm[3] = {a[0] / area, a[1] / area, a[2] / area};
      const double dr[3] = {A.xc[best] - A.xc[c], A.yc[best] - A.yc[c], A.zc[best] - A.zc[c]};
      const double ds = dot3(dr, norm);
      const double d = tc_e * area / ds / cp_e;  // 'dirichlet' (both energy callbacks)
      ap = ap + d + 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k == bk) anbk[k] = anbk[k] + d + 0.0;
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (k < n) A.anb[(size_t)k * Np + c] = anbk[k];
    A.ap[c] = ap;
    A.b[c] = b;
  }
}

__global__ void __launch_bounds__(TPB) temperature_kernel(int N, const double* __restrict__ hh, const double* __restrict__ cp, double* t) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) t[c] = hh[c] / cp[c];
}

struct ScalarArgs {
  int N, Np;
  const int32_t *ell_nb, *ell_fs;
  const uint8_t* nfc;
  const double *xc, *yc, *zc, *aip, *vol, *s, *s0;
  double *ap, *anb, *b;
  double dt, dcoef, vel[3];
};

// calc_coef_scalar, mod_scalar.f90:74-126: every face of the cell, boundary faces included (dr reaches the halo centre)
template <int K>
__global__ void __launch_bounds__(TPB) coef_scalar_kernel(const ScalarArgs A) {
  const int Np = A.Np;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < A.N; c += gridDim.x * blockDim.x) {
    const int n = A.nfc[c];
    const double rp[3] = {A.xc[c], A.yc[c], A.zc[c]};
    double ap = 0.0, sumf = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k < n) {
        const int nb = A.ell_nb[(size_t)k * Np + c];
        const int fs = A.ell_fs[(size_t)k * Np + c];
        const int f = abs(fs) - 1;
        const double sg = fs > 0 ? 1.0 : -1.0;
        double a[3];
        load3(A.aip, f, a);
        const double area = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        const double norm[3] = {sg * a[0] / area, sg * a[1] / area, sg * a[2] / area};
        const double dr[3] = {A.xc[nb] - rp[0], A.yc[nb] - rp[1], A.zc[nb] - rp[2]};
        const double mnorm[3] = {-norm[0], -norm[1], -norm[2]};
        const double f_in = dot3(A.vel, mnorm) * area;
        const double wnb = f_in > 0.0 ? 1.0 : 0.0;
        const double fnb = wnb * f_in;
        sumf = sumf + f_in;
        const double d = A.dcoef / dot3(dr, dr) * dot3(dr, norm) * area;
        A.anb[(size_t)k * Np + c] = d + fnb;
        ap = ap + d + fnb;
      }
    }
    const double ap0 = A.vol[c] / A.dt;
    ap = ap + ap0;
    A.ap[c] = ap;
    A.b[c] = ap0 * A.s0[c] + sumf * A.s[c];
  }
}

__global__ void __launch_bounds__(TPB) fill_kernel(double* a, double v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = v;
}
__global__ void __launch_bounds__(TPB) mul_cp_kernel(int Nc, int H, const double* __restrict__ t, const double* __restrict__ cp, double* hh) {
  // construct_energy :33 phi = t*cp (cp has one entry per cell; the halo entries are set by the boundary callbacks before use)
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < H; c += gridDim.x * blockDim.x) hh[c] = t[c] * cp[c < Nc ? c : 0];
}

static int ensure(Handle* h, double*& p, size_t n) {
  if (p) return CFDL_OK;
  CFDL_CUDA(cudaMalloc(&p, sizeof(double) * (n + 4)));
  h->allocs.push_back(p);
  CFDL_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * (n + 4), h->stream));
  return CFDL_OK;
}

static int transport_supported(const Handle* h, const char* what) {
  if (h->K > 6) return fail(CFDL_ERR_UNSUPPORTED, "%s: cells with more than 6 faces are not supported", what);
  return CFDL_OK;
}

// solve_equation for a field that lives outside the peer-to-peer slab: the two-colour passes store interface values straight
// into the neighbours' copies of the array they sweep, so that array must be one every rank exports
static int solve_transport(Handle* h, int eq, double* phi, int nit, double* out4) {
  const bool scratch = h->prep.nranks > 1 && h->p2p.connected && h->use_p2p;
  if (!scratch) return solve_equation(h, eq, phi, h->fld[CFDL_F_B], nit, out4, false);
  double* w = h->rb3_work[0];
  const size_t bytes = sizeof(double) * (size_t)h->H;
  CFDL_CUDA(cudaMemcpyAsync(w, phi, bytes, cudaMemcpyDeviceToDevice, h->stream));
  int rc = solve_equation(h, eq, w, h->fld[CFDL_F_B], nit, out4, false);
  if (rc) return rc;
  CFDL_CUDA(cudaMemcpyAsync(phi, w, bytes, cudaMemcpyDeviceToDevice, h->stream));
  return CFDL_OK;
}

// construct_energy (mod_energy.f90:14-48) + the tc, cp of init_properties (mod_properties.f90:88-89; NULL: 5 and 1000)
int k_energy_init(Handle* h, const double* tc_host, const double* cp_host) {
  int rc = transport_supported(h, "cfdl_energy_init");
  if (rc) return rc;
  const size_t H = (size_t)h->H;
  if ((rc = ensure(h, h->tc, (size_t)h->Nc)) || (rc = ensure(h, h->cp, (size_t)h->Nc)) || (rc = ensure(h, h->fld[CFDL_F_T], H)) ||
      (rc = ensure(h, h->fld[CFDL_F_H], H)) || (rc = ensure(h, h->fld[CFDL_F_H0], H)) || (rc = ensure(h, h->fld[CFDL_F_GT], 3 * H)) ||
      (rc = ensure(h, h->fld[CFDL_F_GH], 3 * H)))
    return rc;
  const int g = grid_for(h, h->H, TPB);
  // per-cell arrays in the reference's numbering -> device numbering (owned cells, then the ghosts of a partition)
  auto put = [&](double* dst, const double* host, double dflt) -> int {
    if (!host) { fill_kernel<<<g, TPB, 0, S(h)>>>(dst, dflt, h->Nc); return CFDL_OK; }
    if (h->prep.nranks > 1) {
      std::vector<double> t((size_t)h->Nc);
      for (int32_t c = 0; c < h->Nc; ++c) t[(size_t)c] = host[h->prep.c2o[(size_t)c]];
      CFDL_CUDA(cudaMemcpyAsync(dst, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice, h->stream));
      CFDL_CUDA(cudaStreamSynchronize(h->stream));  // (t goes out of scope)
      return CFDL_OK;
    }
    CFDL_CUDA(cudaMemcpyAsync(h->stage, host, sizeof(double) * (size_t)h->prep.gN, cudaMemcpyHostToDevice, h->stream));
    return k_gather(h, dst, h->stage, h->c2o, h->Nc, 1);
  };
  if ((rc = put(h->tc, tc_host, 5.0)) || (rc = put(h->cp, cp_host, 1000.0))) return rc;
  fill_kernel<<<g, TPB, 0, S(h)>>>(h->fld[CFDL_F_T], 273.0, h->H);
  mul_cp_kernel<<<g, TPB, 0, S(h)>>>(h->Nc, h->H, h->fld[CFDL_F_T], h->cp, h->fld[CFDL_F_H]);
  CFDL_CUDA(cudaMemcpyAsync(h->fld[CFDL_F_H0], h->fld[CFDL_F_H], sizeof(double) * H, cudaMemcpyDeviceToDevice, h->stream));
  CFDL_CUDA(cudaMemsetAsync(h->fld[CFDL_F_GT], 0, sizeof(double) * 3 * H, h->stream));
  CFDL_CUDA(cudaMemsetAsync(h->fld[CFDL_F_GH], 0, sizeof(double) * 3 * H, h->stream));
  CFDL_CUDA(cudaGetLastError());
  h->has_energy = true;
  return CFDL_OK;
}

int k_scalar_init(Handle* h, double dcoef, const double vel[3], const double* bc_value_host) {
  int rc = transport_supported(h, "cfdl_scalar_init");
  if (rc) return rc;
  const size_t H = (size_t)h->H, nbc = h->prep.bc_kind.size();
  if ((rc = ensure(h, h->fld[CFDL_F_S], H)) || (rc = ensure(h, h->fld[CFDL_F_S0], H)) || (rc = ensure(h, h->fld[CFDL_F_GS], 3 * H)) ||
      (rc = ensure(h, h->s_bc, nbc)))
    return rc;
  CFDL_CUDA(cudaMemsetAsync(h->fld[CFDL_F_S], 0, sizeof(double) * H, h->stream));
  CFDL_CUDA(cudaMemsetAsync(h->fld[CFDL_F_S0], 0, sizeof(double) * H, h->stream));
  CFDL_CUDA(cudaMemsetAsync(h->fld[CFDL_F_GS], 0, sizeof(double) * 3 * H, h->stream));
  if (bc_value_host && nbc) {
    CFDL_CUDA(cudaMemcpyAsync(h->s_bc, bc_value_host, sizeof(double) * nbc, cudaMemcpyHostToDevice, h->stream));
    CFDL_CUDA(cudaStreamSynchronize(h->stream));
  }
  h->s_dcoef = dcoef;
  for (int i = 0; i < 3; ++i) h->s_vel[i] = vel[i];
  h->has_scalar = true;
  return CFDL_OK;
}

// the callbacks update_boundaries runs for the two equations (mod_physics.f90:45,47)
int k_transport_boundaries(Handle* h) {
  if (h->B == 0) return CFDL_OK;
  if (h->has_energy)
    energy_bc_kernel<<<grid_for(h, h->B, TPB), TPB, 0, S(h)>>>(h->B, h->Nc, h->halo_cell, h->halo_bc, h->bc_kind, h->cp, h->fld[CFDL_F_T], h->fld[CFDL_F_H]);
  if (h->has_scalar) scalar_bc_kernel<<<grid_for(h, h->B, TPB), TPB, 0, S(h)>>>(h->B, h->Nc, h->halo_bc, h->s_bc, h->fld[CFDL_F_S]);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

int k_transport_update_time(Handle* h) {  // mod_physics.f90:104 (scalar, commented out there), :110 (energy)
  const size_t hb = sizeof(double) * (size_t)h->H;
  if (h->has_energy) CFDL_CUDA(cudaMemcpyAsync(h->fld[CFDL_F_H0], h->fld[CFDL_F_H], hb, cudaMemcpyDeviceToDevice, h->stream));
  if (h->has_scalar) CFDL_CUDA(cudaMemcpyAsync(h->fld[CFDL_F_S0], h->fld[CFDL_F_S], hb, cudaMemcpyDeviceToDevice, h->stream));
  return CFDL_OK;
}

int k_solve_energy(Handle* h, double dt, int nit, double* out4) {  // mod_energy.f90:59-80
  int rc = transport_supported(h, "cfdl_solve_energy");
  if (rc) return rc;
  if (!h->has_energy) return fail(CFDL_ERR_ARG, "cfdl_solve_energy: call cfdl_energy_init first");
  if (!h->use_statics) return fail(CFDL_ERR_UNSUPPORTED, "cfdl_solve_energy needs the face statics (option statics = 1)");
  const bool dist = h->prep.nranks > 1;
  if (dist) {  // ghosts of t and phi (uploads and callbacks touch owned cells and halos only), then of the two gradients
    double* th[2] = {h->fld[CFDL_F_T], h->fld[CFDL_F_H]};
    if ((rc = comm_exchange_multi(h, th, 2, 1))) return rc;
  }
  if ((rc = k_calc_grad(h, h->fld[CFDL_F_T], h->fld[CFDL_F_GT])) || (rc = k_calc_grad(h, h->fld[CFDL_F_H], h->fld[CFDL_F_GH]))) return rc;
  if (dist) {
    double* gg[2] = {h->fld[CFDL_F_GT], h->fld[CFDL_F_GH]};
    if ((rc = comm_exchange_multi(h, gg, 2, 3))) return rc;
  }
  EnergyArgs A;
  A.N = h->N; A.Nc = h->Nc; A.Np = h->Np; A.ell_nb = h->ell_nb; A.ell_fs = h->ell_fs; A.halo_bc = h->halo_bc; A.nfc = h->nfc;
  A.xc = h->xc; A.yc = h->yc; A.zc = h->zc; A.aip = h->aip; A.vol = h->vol; A.rho = h->rho; A.tc = h->tc; A.cp = h->cp;
  A.mip = h->fld[CFDL_F_MIP]; A.hh = h->fld[CFDL_F_H]; A.h0 = h->fld[CFDL_F_H0]; A.gh = h->fld[CFDL_F_GH]; A.gt = h->fld[CFDL_F_GT];
  A.fs_area = h->fs_area; A.fs_dn = h->fs_dn; A.fs_wto = h->fs_wto; A.fs_wtn = h->fs_wtn;
  for (int i = 0; i < 3; ++i) A.fs_n[i] = h->fs_n[i];
  A.ap = h->fld[CFDL_F_AP]; A.anb = h->fld[CFDL_F_ANB]; A.b = h->fld[CFDL_F_B]; A.dt = dt;
  h->pc_sumap_ok = false;  // ap / anb no longer hold the pc matrix
  if (h->K <= 4) coef_energy_kernel<4><<<grid_for(h, h->N, TPB, 4), TPB, 0, S(h)>>>(A);
  else coef_energy_kernel<6><<<grid_for(h, h->N, TPB, 4), TPB, 0, S(h)>>>(A);
  CFDL_CUDA(cudaGetLastError());
  if ((rc = solve_transport(h, CFDL_EQ_E, h->fld[CFDL_F_H], nit, out4))) return rc;
  // (owned cells and ghosts: the solve leaves the ghosts of phi current, cp carries them)
  temperature_kernel<<<grid_for(h, h->Nc, TPB), TPB, 0, S(h)>>>(h->Nc, h->fld[CFDL_F_H], h->cp, h->fld[CFDL_F_T]);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

int k_solve_scalar(Handle* h, double dt, int nit, double* out4) {  // mod_scalar.f90:56-72
  int rc = transport_supported(h, "cfdl_solve_scalar");
  if (rc) return rc;
  if (!h->has_scalar) return fail(CFDL_ERR_ARG, "cfdl_solve_scalar: call cfdl_scalar_init first");
  if (h->prep.nranks > 1 && (rc = comm_exchange(h, h->fld[CFDL_F_S], 1, -1))) return rc;  // ghosts of phi for the gradient
  ScalarArgs A;
  A.N = h->N; A.Np = h->Np; A.ell_nb = h->ell_nb; A.ell_fs = h->ell_fs; A.nfc = h->nfc;
  A.xc = h->xc; A.yc = h->yc; A.zc = h->zc; A.aip = h->aip; A.vol = h->vol; A.s = h->fld[CFDL_F_S]; A.s0 = h->fld[CFDL_F_S0];
  A.ap = h->fld[CFDL_F_AP]; A.anb = h->fld[CFDL_F_ANB]; A.b = h->fld[CFDL_F_B]; A.dt = dt; A.dcoef = h->s_dcoef;
  for (int i = 0; i < 3; ++i) A.vel[i] = h->s_vel[i];
  h->pc_sumap_ok = false;
  if (h->K <= 4) coef_scalar_kernel<4><<<grid_for(h, h->N, TPB, 4), TPB, 0, S(h)>>>(A);
  else coef_scalar_kernel<6><<<grid_for(h, h->N, TPB, 4), TPB, 0, S(h)>>>(A);
  CFDL_CUDA(cudaGetLastError());
  if ((rc = k_calc_grad(h, h->fld[CFDL_F_S], h->fld[CFDL_F_GS]))) return rc;
  if (h->prep.nranks > 1 && (rc = comm_exchange(h, h->fld[CFDL_F_GS], 3, -1))) return rc;
  return solve_transport(h, CFDL_EQ_S, h->fld[CFDL_F_S], nit, out4);
}

}  // namespace cfdl
