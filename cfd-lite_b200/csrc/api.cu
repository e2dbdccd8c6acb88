// C ABI of libcfdl (include/cfdl.h): lifecycle, host<->device field sync, the whole-step path
// (update_boundaries / solve_uvwp / update_time, src/main.f90:50-63) and the per-routine path.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include "state.h"

using namespace cfdl;

namespace {

template <class T>
int dev_upload(Handle* h, T*& dst, const T* src, size_t n) {
  dst = nullptr;
  if (n == 0) return CFDL_OK;
  CFDL_CUDA(cudaMalloc(&dst, sizeof(T) * n));
  h->allocs.push_back(dst);
  CFDL_CUDA(cudaMemcpy(dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
  return CFDL_OK;
}
template <class T>
int dev_zero(Handle* h, T*& dst, size_t n) {
  dst = nullptr;
  if (n == 0) n = 1;
  CFDL_CUDA(cudaMalloc(&dst, sizeof(T) * n));
  h->allocs.push_back(dst);
  CFDL_CUDA(cudaMemset(dst, 0, sizeof(T) * n));
  return CFDL_OK;
}

// the fields of the energy / scalar equations (appended to the enum) fall into the same two categories
inline bool is_cell_scalar(int f) { return f <= CFDL_F_PC || (f >= CFDL_F_T && f <= CFDL_F_S0); }
inline bool is_cell_vector(int f) { return (f >= CFDL_F_GU && f <= CFDL_F_GPC) || (f >= CFDL_F_GT && f <= CFDL_F_GS); }

// device-side length of a field (device numbering)
size_t field_len(const Handle* h, int f) {
  if (is_cell_scalar(f)) return (size_t)h->H;
  if (is_cell_vector(f)) return 3 * (size_t)h->H;
  if (f <= CFDL_F_MIP0) return (size_t)h->F;
  if (f == CFDL_F_ANB) return (size_t)h->K * h->Np;  // device ELL storage
  if (f == CFDL_F_D || f == CFDL_F_DC) return (size_t)h->Nc;  // gathered at neighbours: ghost copies needed
  return (size_t)h->N;
}
// host-side length (reference numbering of the global mesh)
size_t host_len(const Handle* h, int f) {
  const Prep& p = h->prep;
  if (is_cell_scalar(f)) return (size_t)p.gN + p.gB;
  if (is_cell_vector(f)) return 3 * ((size_t)p.gN + p.gB);
  if (f <= CFDL_F_MIP0) return (size_t)p.gF;
  if (f == CFDL_F_ANB) return (size_t)p.gZ;
  return (size_t)p.gN;
}

bool good(cfdl_handle h) { return h != nullptr; }

int use_device(Handle* h) {
  CFDL_CUDA(cudaSetDevice(h->device));
  return CFDL_OK;
}

// host (reference numbering, global mesh) -> device field (owned + ghost cells, local halos/faces)
int upload_field(Handle* h, int f, const double* host) {
  if (f == CFDL_F_AP || f == CFDL_F_ANB) h->pc_sumap_ok = false;  // caller-supplied matrix: the diagonal must be read
  CFDL_CUDA(cudaMemcpyAsync(h->stage, host, sizeof(double) * host_len(h, f), cudaMemcpyHostToDevice, h->stream));
  if (f == CFDL_F_ANB) return k_csr_to_ell(h, h->fld[f], h->stage);
  if (is_cell_scalar(f)) return k_gather(h, h->fld[f], h->stage, h->cellmap, h->H, 1);
  if (is_cell_vector(f)) return k_gather(h, h->fld[f], h->stage, h->cellmap, h->H, 3);
  if (f <= CFDL_F_MIP0) return k_gather(h, h->fld[f], h->stage, h->f2o, h->F, 1);
  return k_gather(h, h->fld[f], h->stage, h->c2o, h->N, 1);
}

// device field -> host.  One rank: the whole host array is written.  Several ranks: only the
// entries this rank owns (cells, its boundary halos, faces whose owner cell it owns) are written.
int download_field(Handle* h, int f, double* host) {
  int rc;
  const Prep& p = h->prep;
  if (p.nranks == 1) {
    if (f == CFDL_F_ANB) rc = k_ell_to_csr(h, h->stage, h->fld[f]);
    else if (is_cell_scalar(f)) rc = k_scatter(h, h->stage, h->fld[f], h->cellmap, h->H, 1);
    else if (is_cell_vector(f)) rc = k_scatter(h, h->stage, h->fld[f], h->cellmap, h->H, 3);
    else if (f <= CFDL_F_MIP0) rc = k_scatter(h, h->stage, h->fld[f], h->f2o, h->F, 1);
    else rc = k_scatter(h, h->stage, h->fld[f], h->c2o, h->N, 1);
    if (rc) return rc;
    CFDL_CUDA(cudaMemcpyAsync(host, h->stage, sizeof(double) * host_len(h, f), cudaMemcpyDeviceToHost, h->stream));
    CFDL_CUDA(cudaStreamSynchronize(h->stream));
    return CFDL_OK;
  }
  std::vector<double> t;
  if (f == CFDL_F_ANB) {
    if ((rc = k_ell_to_csr(h, h->stage, h->fld[f]))) return rc;
    t.resize((size_t)p.gZ);
    CFDL_CUDA(cudaMemcpyAsync(t.data(), h->stage, sizeof(double) * t.size(), cudaMemcpyDeviceToHost, h->stream));
    CFDL_CUDA(cudaStreamSynchronize(h->stream));
    for (int32_t c = 0; c < p.N; ++c)
      for (int32_t i = p.row_ptr[p.c2o[c]]; i < p.row_ptr[p.c2o[c] + 1]; ++i) host[i] = t[i];
    return CFDL_OK;
  }
  t.resize(field_len(h, f));
  CFDL_CUDA(cudaMemcpyAsync(t.data(), h->fld[f], sizeof(double) * t.size(), cudaMemcpyDeviceToHost, h->stream));
  CFDL_CUDA(cudaStreamSynchronize(h->stream));
  if (is_cell_scalar(f) || is_cell_vector(f)) {
    const int nc = is_cell_scalar(f) ? 1 : 3;
    for (int32_t c = 0; c < p.N; ++c) for (int q = 0; q < nc; ++q) host[(size_t)p.c2o[c] * nc + q] = t[(size_t)c * nc + q];
    for (int32_t j = 0; j < p.B; ++j) for (int q = 0; q < nc; ++q) host[((size_t)p.gN + p.h2o[j]) * nc + q] = t[((size_t)p.Nc + j) * nc + q];
  } else if (f <= CFDL_F_MIP0) {
    for (int32_t i = 0; i < p.F; ++i) if (p.fown[i]) host[p.f2o[i]] = t[i];
  } else {
    for (int32_t c = 0; c < p.N; ++c) host[p.c2o[c]] = t[c];
  }
  return CFDL_OK;
}

int rhs_field(int eq) { return eq == CFDL_EQ_U ? CFDL_F_BU : eq == CFDL_EQ_V ? CFDL_F_BV : eq == CFDL_EQ_W ? CFDL_F_BW : CFDL_F_B; }
int phi_field(int eq) { return eq == CFDL_EQ_U ? CFDL_F_U : eq == CFDL_EQ_V ? CFDL_F_V : eq == CFDL_EQ_W ? CFDL_F_W : CFDL_F_PC; }

// solve_uvwp, src/equations/mod_uvwp.f90:95-134.  On several GPUs the ghost-cell copies of the
// fields a later kernel gathers are refreshed right after the kernel that produces them.
// `hook`, when given, is called at the points of the iteration where a host driver's transfers
// can run beside the computation (cfdl_step_host): after the momentum solves (u, v, w are final —
// the reference's velocity correction is disabled, mod_uvwp.f90:387-391), after their gradients
// (gu, gv, gw are final) and right before calc_mip (the first reader of mip0).
enum { STEP_MOMENTUM_DONE = 0, STEP_GRAD_DONE = 1, STEP_BEFORE_MIP = 2 };
struct StepHook { virtual int at(int stage) = 0; virtual ~StepHook() {} };

int solve_uvwp_impl(Handle* h, double dt, int nit, double* hist, StepHook* hook = nullptr) {
  int rc;
  double st[16] = {0};
  if ((rc = k_calc_coef_uvw(h, dt))) return rc;                                              // :111
  {
    double* ddc[2] = {h->fld[CFDL_F_D], h->fld[CFDL_F_DC]};
    if ((rc = comm_exchange_multi(h, ddc, 2, 1))) return rc;
  }
  if ((rc = solve_momentum(h, nit, st))) return rc;                                          // :114-116
  if (hook && (rc = hook->at(STEP_MOMENTUM_DONE))) return rc;
  if ((rc = k_calc_grad3(h))) return rc;                                                     // :118-120
  {
    double* g3[3] = {h->fld[CFDL_F_GU], h->fld[CFDL_F_GV], h->fld[CFDL_F_GW]};
    if ((rc = comm_exchange_multi(h, g3, 3, 3))) return rc;
  }
  if (hook && ((rc = hook->at(STEP_GRAD_DONE)) || (rc = hook->at(STEP_BEFORE_MIP)))) return rc;
  if ((rc = k_calc_mip(h, true, dt))) return rc;                                             // :122
  if ((rc = k_calc_coef_p(h))) return rc;                                                    // :124
  CFDL_CUDA(cudaMemsetAsync(h->fld[CFDL_F_PC], 0, sizeof(double) * (size_t)h->H, h->stream)); // :126 set_a_0
  if ((rc = solve_equation(h, CFDL_EQ_PC, h->fld[CFDL_F_PC], h->fld[CFDL_F_B], nit, st + 12, true))) return rc;  // :127
  if ((rc = k_adjust_pc(h))) return rc;                                                      // :129-130
  if ((rc = k_calc_grad(h, h->fld[CFDL_F_PC], h->fld[CFDL_F_GPC]))) return rc;               // :131
  if ((rc = k_update_uvwp(h))) return rc;                                                    // :132
  if ((rc = comm_exchange(h, h->fld[CFDL_F_GP], 3, -1))) return rc;
  if (hist) std::memcpy(hist, st, sizeof st);
  return CFDL_OK;
}

int create_impl(cfdl_handle* out, int32_t ne, int32_t nf, int32_t nbf, const int32_t* ef2nb_idx, const int32_t* ef2nb_nb,
                const int32_t* ef2nb_fg, const int32_t* s2g, const int32_t* bs, const double* xc, const double* yc,
                const double* zc, const double* aip, const double* rip, const double* vol, const double* rho,
                const double* mu, int32_t nbc, const int32_t* bc_esec, const int32_t* bc_kind, const double* bc_uvw,
                int32_t n_subdomains, const int32_t* g2gf_p, const int32_t* g2gf_idx, const int32_t* cell2rank, int32_t rank,
                int32_t nranks, int32_t device) {
  if (!out) return fail(CFDL_ERR_ARG, "cfdl_create: out is NULL");
  *out = nullptr;
  if (!ef2nb_idx || !ef2nb_nb || !ef2nb_fg || !s2g || (!bs && nbf) || !xc || !yc || !zc || !aip || !rip || !vol || !rho || !mu)
    return fail(CFDL_ERR_ARG, "cfdl_create: NULL mesh array");
  if (n_subdomains < 1) return fail(CFDL_ERR_ARG, "cfdl_create: n_subdomains must be >= 1");
  int ndev = cfdl_device_count();
  if (ndev < 1) return fail(CFDL_ERR_CUDA, "cfdl_create: no CUDA device is usable (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(CFDL_ERR_ARG, "cfdl_create: device %d of %d", device, ndev);
  cfdl_handle_s* h = new (std::nothrow) cfdl_handle_s;
  if (!h) return fail(CFDL_ERR_INTERNAL, "out of host memory");
  h->device = device;
  int rc = prepare(h->prep, ne, nf, nbf, ef2nb_idx, ef2nb_nb, ef2nb_fg, s2g, bs, xc, yc, zc, nbc, bc_esec, bc_kind, bc_uvw,
                   n_subdomains, g2gf_p, g2gf_idx, /*reorder auto*/ 2, cell2rank, rank, nranks);
  if (rc) { delete h; return rc; }
  GeomSource G;  // geometry straight from the caller's (global, reference-numbered) arrays
  G.cell_xyz = [=](int32_t g, double* o) { o[0] = xc[g]; o[1] = yc[g]; o[2] = zc[g]; };
  G.vol = [=](int32_t g) { return vol[g]; };
  G.rho = [=](int32_t g) { return rho[g]; };
  G.mu = [=](int32_t g) { return mu[g]; };
  G.face = [=](int32_t f, double* a, double* r) {
    for (int q = 0; q < 3; ++q) { a[q] = aip[3 * (size_t)f + q]; r[q] = rip[3 * (size_t)f + q]; }
  };
  return create_from_prep(h, G, out);
}

}  // namespace

namespace cfdl {

// Device side of cfdl_create*: everything after the host-side mesh preparation.
int create_from_prep(cfdl_handle_s* h, const GeomSource& G, cfdl_handle* out) {
  int rc;
  const int device = h->device, nranks = h->prep.nranks;
  auto bail = [&](int code) { cfdl_destroy(h); return code; };
  if ((rc = use_device(h))) return bail(rc);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(fail(CFDL_ERR_CUDA, "cudaGetDeviceProperties failed"));
  if (prop.major < 10) return bail(fail(CFDL_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor));
  h->num_sms = prop.multiProcessorCount;
  h->l2_bytes = (size_t)std::max(0, prop.l2CacheSize);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(CFDL_ERR_CUDA, "cudaStreamCreate failed"));
  const Prep& p = h->prep;
  h->N = p.N; h->G = p.G; h->Nc = p.Nc; h->F = p.F; h->B = p.B; h->H = p.H; h->Z = p.Z; h->K = p.K; h->Np = p.Np; h->Fi = p.Fi;
  h->ne_global = p.gN;
  h->nnbr = (int)p.nbr_rank.size();
  if (nranks > 1) h->solver_mode = CFDL_SOLVER_MCSGS;
  const int32_t N = p.N, Nc = p.Nc, F = p.F, B = p.B, H = p.H;
#define UP(dst, vec) if ((rc = dev_upload(h, dst, (vec).data(), (vec).size()))) return bail(rc)
  UP(h->ell_nb, p.ell_nb); UP(h->ell_fs, p.ell_fs); UP(h->nfc, p.nfc); UP(h->ftouch, p.ftouch); UP(h->face_a, p.face_a); UP(h->face_b, p.face_b);
  UP(h->halo_cell, p.halo_cell); UP(h->halo_face, p.halo_face); UP(h->halo_bc, p.halo_bc); UP(h->halo_slot, p.halo_slot);
  UP(h->bc_kind, p.bc_kind); UP(h->bc_uvw, p.bc_uvw); UP(h->c2o, p.c2o); UP(h->f2o, p.f2o); UP(h->row_ptr, p.row_ptr);
  UP(h->loc_order, p.loc_order); UP(h->ell_nb16, p.nb16);
  UP(h->send_cells, p.send_cells); UP(h->tgt_ptr, p.tgt_ptr); UP(h->tgt_nbr, p.tgt_nbr); UP(h->tgt_pos, p.tgt_pos);
#undef UP
  {  // global index of every device cell | halo (host transfers), geometry in device numbering
    std::vector<int32_t> cm((size_t)H);
    for (int32_t c = 0; c < Nc; ++c) cm[c] = p.c2o[c];
    for (int32_t j = 0; j < B; ++j) cm[Nc + j] = p.gN + p.h2o[j];
    if ((rc = dev_upload(h, h->cellmap, cm.data(), cm.size()))) return bail(rc);
    std::vector<double> tx((size_t)H), ty((size_t)H), tz((size_t)H), t((size_t)std::max(Nc, 1)), ta(3 * (size_t)F), tr(3 * (size_t)F);
    for (int32_t i = 0; i < H; ++i) { double o[3]; G.cell_xyz(cm[i], o); tx[i] = o[0]; ty[i] = o[1]; tz[i] = o[2]; }
    if ((rc = dev_upload(h, h->xc, tx.data(), (size_t)H)) || (rc = dev_upload(h, h->yc, ty.data(), (size_t)H)) ||
        (rc = dev_upload(h, h->zc, tz.data(), (size_t)H)))
      return bail(rc);
    for (int32_t i = 0; i < N; ++i) t[i] = G.vol(cm[i]);
    if ((rc = dev_upload(h, h->vol, t.data(), (size_t)N))) return bail(rc);
    for (int32_t i = 0; i < Nc; ++i) t[i] = G.rho(cm[i]);
    if ((rc = dev_upload(h, h->rho, t.data(), (size_t)Nc))) return bail(rc);
    for (int32_t i = 0; i < Nc; ++i) t[i] = G.mu(cm[i]);
    if ((rc = dev_upload(h, h->mu, t.data(), (size_t)Nc))) return bail(rc);
    for (int32_t f = 0; f < F; ++f) G.face(p.f2o[f], &ta[3 * (size_t)f], &tr[3 * (size_t)f]);
    if ((rc = dev_upload(h, h->aip, ta.data(), 3 * (size_t)F)) || (rc = dev_upload(h, h->rip, tr.data(), 3 * (size_t)F))) return bail(rc);
  }
  // several GPUs: u, v, w, pc and the second solver array live in one IPC-exportable slab so
  // that neighbours can write their ghost cells directly (cfdl_comm_ipc_connect)
  if (nranks > 1 && (rc = p2p_alloc_slab(h))) return bail(rc);
  for (int f = 0; f <= CFDL_F_ANB; ++f)  // (the fields of the energy / scalar equations are allocated by their init calls)
    if (!h->fld[f] && (rc = dev_zero(h, h->fld[f], field_len(h, f) + 4))) return bail(rc);
  h->stage_len = std::max(std::max(3 * ((size_t)p.gN + p.gB), (size_t)p.gZ), (size_t)p.gF) + 4;
  if ((rc = dev_zero(h, h->stage, h->stage_len))) return bail(rc);
  h->partial_len = 4096 + 2 * 64 * (N / 8192 + 1);
  if ((rc = dev_zero(h, h->partial, (size_t)h->partial_len))) return bail(rc);
  if ((rc = dev_zero(h, h->ctl, 1)) || (rc = dev_zero(h, h->scal, 512)) || (rc = dev_zero(h, h->barrier, 4))) return bail(rc);
  if (nranks > 1 && (rc = dev_zero(h, h->send_buf, 9 * p.send_cells.size() + 16))) return bail(rc);
  if (cudaMallocHost(&h->ctl_host, sizeof(SolveCtl)) != cudaSuccess || cudaMallocHost(&h->scal_host, sizeof(double) * 512) != cudaSuccess)
    return bail(fail(CFDL_ERR_CUDA, "cudaMallocHost failed"));
  if ((rc = solver_init(h))) return bail(rc);
  {  // static face geometry, evaluated once with the reference's own expressions (kernels_statics.cu)
    double** arrs[18] = {&h->fs_area, &h->fs_ds, &h->fs_dsp, &h->fs_dn, &h->fs_wto, &h->fs_wtn, &h->fs_n[0], &h->fs_n[1], &h->fs_n[2],
                         &h->fs_dr[0], &h->fs_dr[1], &h->fs_dr[2], &h->fs_drp[0], &h->fs_drp[1], &h->fs_drp[2], &h->fs_rds, &h->fs_rdsp, &h->fs_rdn};
    for (double** a : arrs)
      if ((rc = dev_zero(h, *a, (size_t)p.Fi + 4))) return bail(rc);
    if ((rc = k_face_statics(h))) return bail(rc);
  }
  // construct_uvwp: fields zero, mip from calc_mip(.false.), mip0 = mip (mod_uvwp.f90:57-82)
  if ((rc = k_calc_mip(h, false, 0.01))) return bail(rc);
  if (cudaMemcpyAsync(h->fld[CFDL_F_MIP0], h->fld[CFDL_F_MIP], sizeof(double) * (size_t)F, cudaMemcpyDeviceToDevice, h->stream) != cudaSuccess ||
      cudaStreamSynchronize(h->stream) != cudaSuccess)
    return bail(fail(CFDL_ERR_CUDA, "initial calc_mip failed: %s", cudaGetErrorString(cudaGetLastError())));
  *out = h;
  return CFDL_OK;
}

}  // namespace

extern "C" {

int cfdl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int cfdl_create(cfdl_handle* out, int32_t ne, int32_t nf, int32_t nbf, const int32_t* ef2nb_idx, const int32_t* ef2nb_nb,
                const int32_t* ef2nb_fg, const int32_t* s2g, const int32_t* bs, const double* xc, const double* yc,
                const double* zc, const double* aip, const double* rip, const double* vol, const double* rho,
                const double* mu, int32_t nbc, const int32_t* bc_esec, const int32_t* bc_kind, const double* bc_uvw,
                int32_t n_subdomains, const int32_t* g2gf_p, const int32_t* g2gf_idx, int32_t device) {
  return create_impl(out, ne, nf, nbf, ef2nb_idx, ef2nb_nb, ef2nb_fg, s2g, bs, xc, yc, zc, aip, rip, vol, rho, mu, nbc, bc_esec, bc_kind,
                     bc_uvw, n_subdomains, g2gf_p, g2gf_idx, nullptr, 0, 1, device);
}

int cfdl_create_distributed(cfdl_handle* out, int32_t ne, int32_t nf, int32_t nbf, const int32_t* ef2nb_idx, const int32_t* ef2nb_nb,
                            const int32_t* ef2nb_fg, const int32_t* s2g, const int32_t* bs, const double* xc, const double* yc,
                            const double* zc, const double* aip, const double* rip, const double* vol, const double* rho,
                            const double* mu, int32_t nbc, const int32_t* bc_esec, const int32_t* bc_kind, const double* bc_uvw,
                            const int32_t* cell2rank, int32_t rank, int32_t nranks, int32_t device) {
  if (nranks > 1 && !cell2rank) return fail(CFDL_ERR_ARG, "cfdl_create_distributed: cell2rank is NULL");
  return create_impl(out, ne, nf, nbf, ef2nb_idx, ef2nb_nb, ef2nb_fg, s2g, bs, xc, yc, zc, aip, rip, vol, rho, mu, nbc, bc_esec, bc_kind,
                     bc_uvw, 1, nullptr, nullptr, cell2rank, rank, nranks, device);
}

int cfdl_destroy(cfdl_handle h) {
  if (!h) return CFDL_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  comm_destroy(h);
  for (cudaEvent_t e : h->prof_ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->timer_ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->xfer_evs) if (e) cudaEventDestroy(e);
  if (h->xfer_in) { cudaStreamSynchronize(h->xfer_in); cudaStreamDestroy(h->xfer_in); }
  if (h->xfer_out) { cudaStreamSynchronize(h->xfer_out); cudaStreamDestroy(h->xfer_out); }
  for (void* p : h->allocs) cudaFree(p);
  if (h->ctl_host) cudaFreeHost(h->ctl_host);
  if (h->ctl3_host) cudaFreeHost(h->ctl3_host);
  if (h->scal_host) cudaFreeHost(h->scal_host);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CFDL_OK;
}

int cfdl_set_option(cfdl_handle h, const char* key, double value) {
  if (!good(h) || !key) return fail(CFDL_ERR_ARG, "cfdl_set_option: bad handle/key");
  if (!std::strcmp(key, "solver")) {
    int m = (int)value;
    if (m < CFDL_SOLVER_PARITY || m > CFDL_SOLVER_PCG) return fail(CFDL_ERR_ARG, "cfdl_set_option: solver mode %d", m);
    if (m == CFDL_SOLVER_PARITY && h->prep.nranks > 1)
      return fail(CFDL_ERR_UNSUPPORTED, "cfdl_set_option: the exact natural-order solver needs the whole mesh on one GPU");
    h->solver_mode = m;
    return CFDL_OK;
  }
  if (!std::strcmp(key, "fused")) { h->fused_rb = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "p2p")) { h->use_p2p = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "occ_grids")) { h->occ_grids = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "pdl_rows")) { h->pdl_rows = std::max(0, (int)value); return CFDL_OK; }
  if (!std::strcmp(key, "pdl")) { h->use_pdl = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "rb_idx16")) { h->rb_idx16 = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "mip_hoist")) { h->mip_hoist = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "pcg_precond")) { h->pcg_precond = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "rbq")) { h->rbq = value != 0.0; if (value == 2.0) h->rbq_refused = 0; return CFDL_OK; }
  // rbq_counter: the persistent pc solve hands its chunks out from a counter: -1 = only where one chunk per CTA would be too long (default), 0 = never
  // (such meshes run pass by pass), 1 = always.  rbq_l2_fraction: largest share of the L2 the value arrays may take for the persistent form.
  // (partitioned handles: set both before cfdl_comm_ipc_handle, which fixes the chunks)
  if (!std::strcmp(key, "rbq_counter")) { h->rbq_counter = value < 0 ? -1 : (value != 0.0); h->rbq_refused = 0; return CFDL_OK; }
  if (!std::strcmp(key, "rbq_l2_fraction")) { h->rbq_l2_fraction = value; h->rbq_refused = 0; return CFDL_OK; }
  if (!std::strcmp(key, "rbq_prefetch")) { h->rbq_prefetch = std::max(-1, std::min(4, (int)value)); return CFDL_OK; }
  if (!std::strcmp(key, "rbq_lmax")) { h->rbq_lmax = std::max(0, (int)value); h->rbq_refused = 0; return CFDL_OK; }
  if (!std::strcmp(key, "rbq_lbig")) { h->rbq_lbig = std::max(0, (int)value); return CFDL_OK; }
  if (!std::strcmp(key, "rbq_cap")) { h->rbq_cap = std::max(0, (int)value); return CFDL_OK; }
  if (!std::strcmp(key, "rbq_ctas")) { h->rbq_ctas_per_sm = std::max(0, (int)value); return CFDL_OK; }
  if (!std::strcmp(key, "pc_sumap")) { h->pc_sumap = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "uvw_fused")) { h->uvw_fused = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "grad_variant")) { h->grad_variant = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "statics")) { h->use_statics = value != 0.0; return CFDL_OK; }
  if (!std::strcmp(key, "ctas_per_sm")) { h->tune_ctas = std::max(1, (int)value); return CFDL_OK; }
  if (!std::strcmp(key, "profile")) {
    int rc = prof_collect(h);
    h->profile = (value == 2.0) ? 2 : (value != 0.0);
    return rc;
  }
  if (!std::strcmp(key, "reset_counters")) {
    int rc = prof_collect(h);
    h->launches = 0;
    for (int k = 0; k < PROF_KINDS; ++k) { h->prof_ms[k] = 0.0; h->prof_n[k] = 0; }
    return rc;
  }
  return fail(CFDL_ERR_ARG, "cfdl_set_option: unknown key '%s'", key);
}

int cfdl_get_info(cfdl_handle h, const char* key, double* value) {
  if (!good(h) || !key || !value) return fail(CFDL_ERR_ARG, "cfdl_get_info: bad argument");
  if (!std::strcmp(key, "ncolors")) *value = h->prep.ncolors;
  else if (!std::strcmp(key, "morton")) *value = h->prep.morton ? 1 : 0;
  else if (!std::strcmp(key, "nlevels_natural")) *value = h->prep.natural.nlevels;
  else if (!std::strcmp(key, "nlevels_blocks")) *value = h->prep.blocks.nlevels;
  else if (!std::strcmp(key, "ell_width")) *value = h->K;
  else if (!std::strcmp(key, "num_sms")) *value = h->num_sms;
  else if (!std::strcmp(key, "solver")) *value = h->solver_mode;
  else if (!std::strcmp(key, "launches")) *value = (double)h->launches;
  else if (!std::strcmp(key, "pc_sumap")) *value = h->pc_sumap;
  else if (!std::strcmp(key, "color_dist")) *value = h->prep.color_dist;
  else if (!std::strcmp(key, "rbq_active")) *value = (h->rbq && !h->rbq_refused && h->rbq_occ > 0) ? 1 : 0;
  else if (!std::strcmp(key, "rbq_refused")) *value = h->rbq_refused;
  else if (!std::strcmp(key, "rbq_chunks")) *value = h->rbq_last_chunks;  // chunks per colour / CTAs / rows per chunk of the last persistent pc solve
  else if (!std::strcmp(key, "rbq_grid")) *value = h->rbq_last_grid;      // (rbq_grid < rbq_chunks: chunks dealt round-robin to the CTAs)
  else if (!std::strcmp(key, "rbq_chunk_rows")) *value = h->rbq_last_L;
  else if (!std::strcmp(key, "rbq_dist")) *value = h->rbq_dist_state;  // partitioned: 1 = persistent pc solve with chunk-to-chunk synchronisation over NVLink in use, -1 = refused (lists too long / not two colours), 0 = not decided yet
  else if (!std::strcmp(key, "rbq_occ")) *value = h->rbq_occ;
  else if (!std::strcmp(key, "rb_idx16")) *value = (h->rb_idx16 && h->ell_nb16) ? 1 : 0;
  else if (!std::strcmp(key, "rbq_extra_passes")) *value = (double)h->prof_extra_passes;
  else if (!std::strcmp(key, "owned_cells")) *value = h->N;
  else if (!std::strcmp(key, "ghost_cells")) *value = h->G;
  else if (!std::strcmp(key, "local_halos")) *value = h->B;
  else if (!std::strcmp(key, "local_faces")) *value = h->F;
  else if (!std::strcmp(key, "neighbour_ranks")) *value = h->nnbr;
  else if (!std::strncmp(key, "prof_ms_", 8) || !std::strncmp(key, "prof_n_", 7)) {
    static const char* names[PROF_KINDS] = {"sgs", "residual", "coef_uvw", "coef_p", "mip", "grad", "levels", "pcg", "sgs3", "residual3", "grad1"};
    const bool is_ms = key[5] == 'm';
    const char* nm = key + (is_ms ? 8 : 7);
    int rc = prof_collect(h);
    if (rc) return rc;
    for (int k = 0; k < PROF_KINDS; ++k)
      if (!std::strcmp(nm, names[k])) { *value = is_ms ? h->prof_ms[k] : (double)h->prof_n[k]; return CFDL_OK; }
    return fail(CFDL_ERR_ARG, "cfdl_get_info: unknown kernel kind '%s'", nm);
  }
  else return fail(CFDL_ERR_ARG, "cfdl_get_info: unknown key '%s'", key);
  return CFDL_OK;
}

int cfdl_get_cell_order(cfdl_handle h, int32_t* c2o, int32_t* color_ptr) {
  if (!good(h)) return fail(CFDL_ERR_ARG, "cfdl_get_cell_order: NULL handle");
  if (c2o) for (int32_t i = 0; i < h->N; ++i) c2o[i] = h->prep.c2o[i] + 1;
  if (color_ptr) for (int c = 0; c <= h->prep.ncolors; ++c) color_ptr[c] = h->prep.color_ptr[c];
  return CFDL_OK;
}

#define ENTER(h)                                                       \
  if (!good(h)) return fail(CFDL_ERR_ARG, "%s: NULL handle", __func__); \
  { int rc__ = use_device(h); if (rc__) return rc__; }
#define FINISH(h)                                                      \
  do {                                                                 \
    CFDL_CUDA(cudaStreamSynchronize((h)->stream));                     \
    CFDL_CUDA(cudaGetLastError());                                     \
    return p2p_check(h); /* partitioned: CFDL_ERR_COMM if a wait on another rank timed out */ \
  } while (0)

int cfdl_upload_field(cfdl_handle h, int field, const double* host) {
  ENTER(h);
  if (field < 0 || field >= CFDL_F_COUNT || !host) return fail(CFDL_ERR_ARG, "cfdl_upload_field: bad field/pointer");
  if (!h->fld[field]) return fail(CFDL_ERR_ARG, "cfdl_upload_field: field %d belongs to the energy / scalar equation: call cfdl_energy_init / cfdl_scalar_init first", field);
  int rc = upload_field(h, field, host);
  if (rc) return rc;
  FINISH(h);
}
int cfdl_download_field(cfdl_handle h, int field, double* host) {
  ENTER(h);
  if (field < 0 || field >= CFDL_F_COUNT || !host) return fail(CFDL_ERR_ARG, "cfdl_download_field: bad field/pointer");
  if (!h->fld[field]) return fail(CFDL_ERR_ARG, "cfdl_download_field: field %d belongs to the energy / scalar equation: call cfdl_energy_init / cfdl_scalar_init first", field);
  return download_field(h, field, host);
}
int cfdl_field_local_size(cfdl_handle h, int field, int64_t* n) {
  if (!good(h) || field < 0 || field >= CFDL_F_COUNT || !n) return fail(CFDL_ERR_ARG, "cfdl_field_local_size: bad argument");
  *n = (int64_t)field_len(h, field);
  return CFDL_OK;
}
int cfdl_upload_field_local(cfdl_handle h, int field, const double* host) {
  ENTER(h);
  if (field < 0 || field >= CFDL_F_COUNT || !host) return fail(CFDL_ERR_ARG, "cfdl_upload_field_local: bad field/pointer");
  if (!h->fld[field]) return fail(CFDL_ERR_ARG, "cfdl_upload_field_local: field %d belongs to the energy / scalar equation: call cfdl_energy_init / cfdl_scalar_init first", field);
  if (field == CFDL_F_AP || field == CFDL_F_ANB) h->pc_sumap_ok = false;
  CFDL_CUDA(cudaMemcpyAsync(h->fld[field], host, sizeof(double) * field_len(h, field), cudaMemcpyHostToDevice, h->stream));
  FINISH(h);
}
int cfdl_download_field_local(cfdl_handle h, int field, double* host) {
  ENTER(h);
  if (field < 0 || field >= CFDL_F_COUNT || !host) return fail(CFDL_ERR_ARG, "cfdl_download_field_local: bad field/pointer");
  if (!h->fld[field]) return fail(CFDL_ERR_ARG, "cfdl_download_field_local: field %d belongs to the energy / scalar equation: call cfdl_energy_init / cfdl_scalar_init first", field);
  CFDL_CUDA(cudaMemcpyAsync(host, h->fld[field], sizeof(double) * field_len(h, field), cudaMemcpyDeviceToHost, h->stream));
  FINISH(h);
}

int cfdl_update_boundaries(cfdl_handle h) { ENTER(h); int rc = k_update_boundaries(h); if (rc) return rc; FINISH(h); }
int cfdl_update_time(cfdl_handle h) { ENTER(h); int rc = k_update_time(h); if (rc) return rc; FINISH(h); }
int cfdl_solve_uvwp(cfdl_handle h, double dt, int32_t nit, double* hist) {
  ENTER(h);
  int rc = solve_uvwp_impl(h, dt, nit, hist);
  if (rc) return rc;
  FINISH(h);
}
int cfdl_run(cfdl_handle h, double dt, int32_t nit, int32_t ntstep, int32_t ncoef, double* hist) {
  ENTER(h);
  int rc;
  for (int ts = 0; ts < ntstep; ++ts) {
    for (int ic = 0; ic < ncoef; ++ic) {
      if ((rc = k_update_boundaries(h))) return rc;
      if ((rc = solve_uvwp_impl(h, dt, nit, hist ? hist + 16 * ((size_t)ts * ncoef + ic) : nullptr))) return rc;
    }
    if ((rc = k_update_time(h))) return rc;
  }
  FINISH(h);
}

// energy and scalar equations (kernels_transport.cu)
int cfdl_energy_init(cfdl_handle h, const double* tc, const double* cp) { ENTER(h); int rc = k_energy_init(h, tc, cp); if (rc) return rc; FINISH(h); }
int cfdl_solve_energy(cfdl_handle h, double dt, int32_t nit, double* out4) { ENTER(h); int rc = k_solve_energy(h, dt, nit, out4); if (rc) return rc; FINISH(h); }
int cfdl_scalar_init(cfdl_handle h, double dcoef, const double* vel, const double* bc_value) {
  ENTER(h);
  if (!vel) return fail(CFDL_ERR_ARG, "cfdl_scalar_init: vel is NULL");
  int rc = k_scalar_init(h, dcoef, vel, bc_value);
  if (rc) return rc;
  FINISH(h);
}
int cfdl_solve_scalar(cfdl_handle h, double dt, int32_t nit, double* out4) { ENTER(h); int rc = k_solve_scalar(h, dt, nit, out4); if (rc) return rc; FINISH(h); }

int cfdl_calc_coef_uvw(cfdl_handle h, double dt) {
  ENTER(h);
  int rc = k_calc_coef_uvw(h, dt);
  if (rc || (rc = comm_exchange(h, h->fld[CFDL_F_D], 1, -1)) || (rc = comm_exchange(h, h->fld[CFDL_F_DC], 1, -1))) return rc;
  FINISH(h);
}
int cfdl_calc_mip(cfdl_handle h, int32_t lrc, double dt) { ENTER(h); int rc = k_calc_mip(h, lrc != 0, dt); if (rc) return rc; FINISH(h); }
int cfdl_calc_coef_p(cfdl_handle h) { ENTER(h); int rc = k_calc_coef_p(h); if (rc) return rc; FINISH(h); }
int cfdl_adjust_pc(cfdl_handle h) { ENTER(h); int rc = k_adjust_pc(h); if (rc) return rc; FINISH(h); }
int cfdl_update_uvwp(cfdl_handle h) {
  ENTER(h);
  int rc = k_update_uvwp(h);
  if (rc || (rc = comm_exchange(h, h->fld[CFDL_F_GP], 3, -1))) return rc;
  FINISH(h);
}
int cfdl_calc_grad(cfdl_handle h, int phi_f, int grad_f) {
  ENTER(h);
  if (phi_f < CFDL_F_U || phi_f > CFDL_F_PC || grad_f < CFDL_F_GU || grad_f > CFDL_F_GPC) return fail(CFDL_ERR_ARG, "cfdl_calc_grad: bad field ids");
  int rc = k_calc_grad(h, h->fld[phi_f], h->fld[grad_f]);
  if (rc || (rc = comm_exchange(h, h->fld[grad_f], 3, -1))) return rc;
  FINISH(h);
}
int cfdl_solve_eq(cfdl_handle h, int eq, int32_t nit, double* out4) {
  ENTER(h);
  if (eq < CFDL_EQ_U || eq > CFDL_EQ_PC) return fail(CFDL_ERR_ARG, "cfdl_solve_eq: bad equation id");
  int rc = solve_equation(h, eq, h->fld[phi_field(eq)], h->fld[rhs_field(eq)], nit, out4, eq == CFDL_EQ_PC);
  if (rc) return rc;
  FINISH(h);
}

// ---- stand-alone drop-ins with host arrays ------------------------------------------------------
int cfdl_host_calc_grad(cfdl_handle h, const double* phi, double* grad) {
  ENTER(h);
  if (!phi || !grad) return fail(CFDL_ERR_ARG, "cfdl_host_calc_grad: NULL array");
  int rc;
  if ((rc = upload_field(h, CFDL_F_PC, phi))) return rc;
  if ((rc = k_calc_grad(h, h->fld[CFDL_F_PC], h->fld[CFDL_F_GPC]))) return rc;
  return download_field(h, CFDL_F_GPC, grad);
}

static int host_solve_common(cfdl_handle h, int eq, double* phi, const double* ap, const double* anb, const double* b, int32_t nit,
                             double* out4, bool dispatch) {
  if (!phi || !ap || !anb || !b) return fail(CFDL_ERR_ARG, "host solve: NULL array");
  if (eq < CFDL_EQ_U || eq > CFDL_EQ_PC) return fail(CFDL_ERR_ARG, "host solve: bad equation id");
  int rc;
  if ((rc = upload_field(h, CFDL_F_AP, ap)) || (rc = upload_field(h, CFDL_F_ANB, anb)) || (rc = upload_field(h, CFDL_F_B, b)) ||
      (rc = upload_field(h, CFDL_F_PC, phi)))
    return rc;
  if ((rc = solve_equation(h, eq, h->fld[CFDL_F_PC], h->fld[CFDL_F_B], nit, out4, dispatch))) return rc;
  return download_field(h, CFDL_F_PC, phi);
}
int cfdl_host_solve_gs(cfdl_handle h, int eq, double* phi, const double* ap, const double* anb, const double* b, int32_t nit, double* out4) {
  ENTER(h);
  return host_solve_common(h, eq, phi, ap, anb, b, nit, out4, false);
}
int cfdl_host_solve(cfdl_handle h, int eq, double* phi, const double* ap, const double* anb, const double* b, int32_t nit, double* out4) {
  ENTER(h);
  return host_solve_common(h, eq, phi, ap, anb, b, nit, out4, true);
}

int cfdl_host_calc_residual(cfdl_handle h, const double* phi, const double* ap, const double* anb, const double* b, double* res,
                            double* res_max) {
  ENTER(h);
  if (!phi || !ap || !anb || !b || !res || !res_max) return fail(CFDL_ERR_ARG, "cfdl_host_calc_residual: NULL array");
  if (h->prep.nranks > 1) return fail(CFDL_ERR_UNSUPPORTED, "cfdl_host_calc_residual: single-GPU drop-in");
  int rc;
  if ((rc = upload_field(h, CFDL_F_AP, ap)) || (rc = upload_field(h, CFDL_F_ANB, anb)) || (rc = upload_field(h, CFDL_F_B, b)) ||
      (rc = upload_field(h, CFDL_F_PC, phi)))
    return rc;
  return residual_plain(h, h->fld[CFDL_F_PC], h->fld[CFDL_F_B], true, res, res_max);
}

}  // extern "C"

// ---- the assembly routines with host arrays (flattened derived types of the reference) ----------
namespace {
struct In { int field; const double* host; };
struct Out { int field; double* host; };
template <size_t NI, size_t NO, typename Run>
int host_routine(cfdl_handle h, const char* who, const In (&ins)[NI], const Out (&outs)[NO], Run run) {
  if (h->prep.nranks > 1) return fail(CFDL_ERR_UNSUPPORTED, "%s: single-GPU drop-in (use the per-routine calls on a partitioned handle)", who);
  for (const In& i : ins) if (!i.host) return fail(CFDL_ERR_ARG, "%s: NULL input array", who);
  for (const Out& o : outs) if (!o.host) return fail(CFDL_ERR_ARG, "%s: NULL output array", who);
  int rc;
  for (const In& i : ins) if ((rc = upload_field(h, i.field, i.host))) return rc;
  if ((rc = run())) return rc;
  for (const Out& o : outs) if ((rc = download_field(h, o.field, o.host))) return rc;
  return CFDL_OK;
}
}  // namespace

extern "C" {

int cfdl_host_calc_coef_uvw(cfdl_handle h, double dt, const double* u, const double* v, const double* w, const double* u0,
                            const double* v0, const double* w0, const double* gu, const double* gv, const double* gw, const double* gp,
                            const double* mip, double* ap, double* anb, double* bu, double* bv, double* bw, double* d, double* dc) {
  ENTER(h);
  const In ins[] = {{CFDL_F_U, u}, {CFDL_F_V, v}, {CFDL_F_W, w}, {CFDL_F_U0, u0}, {CFDL_F_V0, v0}, {CFDL_F_W0, w0},
                    {CFDL_F_GU, gu}, {CFDL_F_GV, gv}, {CFDL_F_GW, gw}, {CFDL_F_GP, gp}, {CFDL_F_MIP, mip}};
  const Out outs[] = {{CFDL_F_AP, ap}, {CFDL_F_ANB, anb}, {CFDL_F_BU, bu}, {CFDL_F_BV, bv}, {CFDL_F_BW, bw}, {CFDL_F_D, d}, {CFDL_F_DC, dc}};
  return host_routine(h, "cfdl_host_calc_coef_uvw", ins, outs, [&] { return k_calc_coef_uvw(h, dt); });
}

int cfdl_host_calc_mip(cfdl_handle h, int32_t l_rhie_chow, double dt, const double* u, const double* v, const double* w, const double* u0,
                       const double* v0, const double* w0, const double* p, const double* gp, const double* d, const double* mip0,
                       double* mip) {
  ENTER(h);
  // mip is in/out: only cell-cell faces are written (mod_uvwp.f90:456), boundary entries pass through
  const In ins[] = {{CFDL_F_U, u}, {CFDL_F_V, v}, {CFDL_F_W, w}, {CFDL_F_U0, u0}, {CFDL_F_V0, v0}, {CFDL_F_W0, w0},
                    {CFDL_F_P, p}, {CFDL_F_GP, gp}, {CFDL_F_D, d}, {CFDL_F_MIP0, mip0}, {CFDL_F_MIP, mip}};
  const Out outs[] = {{CFDL_F_MIP, mip}};
  return host_routine(h, "cfdl_host_calc_mip", ins, outs, [&] { return k_calc_mip(h, l_rhie_chow != 0, dt); });
}

int cfdl_host_calc_coef_p(cfdl_handle h, const double* dc, const double* mip, double* ap, double* anb, double* b) {
  ENTER(h);
  const In ins[] = {{CFDL_F_DC, dc}, {CFDL_F_MIP, mip}};
  const Out outs[] = {{CFDL_F_AP, ap}, {CFDL_F_ANB, anb}, {CFDL_F_B, b}};
  return host_routine(h, "cfdl_host_calc_coef_p", ins, outs, [&] { return k_calc_coef_p(h); });
}

int cfdl_host_adjust_pc(cfdl_handle h, double* pc) {
  ENTER(h);
  const In ins[] = {{CFDL_F_PC, pc}};
  const Out outs[] = {{CFDL_F_PC, pc}};
  return host_routine(h, "cfdl_host_adjust_pc", ins, outs, [&] { return k_adjust_pc(h); });
}

int cfdl_host_update_uvwp(cfdl_handle h, const double* pc, const double* gpc, const double* dc, double* p, double* gp, double* mip) {
  ENTER(h);
  const In ins[] = {{CFDL_F_PC, pc}, {CFDL_F_GPC, gpc}, {CFDL_F_DC, dc}, {CFDL_F_P, p}, {CFDL_F_GP, gp}, {CFDL_F_MIP, mip}};
  const Out outs[] = {{CFDL_F_P, p}, {CFDL_F_GP, gp}, {CFDL_F_MIP, mip}};
  return host_routine(h, "cfdl_host_update_uvwp", ins, outs, [&] { return k_update_uvwp(h); });
}

}  // extern "C"

// ---- one SIMPLE iteration with host arrays, transfers overlapped with the computation -----------
namespace {

// host <-> device copy of one field in the reference numbering through a staging buffer of its
// own; `map`/`ncomp`/`n` describe the permutation (device index i <-> host index map[i])
struct FieldMap { const int32_t* map; int64_t n; int ncomp; };
FieldMap field_map(const Handle* h, int f) {
  if (is_cell_scalar(f)) return {h->cellmap, h->H, 1};
  if (is_cell_vector(f)) return {h->cellmap, h->H, 3};
  if (f <= CFDL_F_MIP0) return {h->f2o, h->F, 1};
  return {h->c2o, h->N, 1};
}

// streams, events and (reference-numbering transfers) a staging area with one slot per field of the call
int ensure_xfer(Handle* h, size_t stage_doubles) {
  if (!h->xfer_in && cudaStreamCreateWithFlags(&h->xfer_in, cudaStreamNonBlocking) != cudaSuccess) return fail(CFDL_ERR_CUDA, "cudaStreamCreate failed");
  if (!h->xfer_out && cudaStreamCreateWithFlags(&h->xfer_out, cudaStreamNonBlocking) != cudaSuccess) return fail(CFDL_ERR_CUDA, "cudaStreamCreate failed");
  if (h->xfer_evs.empty()) {
    h->xfer_evs.assign(2 * CFDL_F_COUNT, nullptr);
    for (cudaEvent_t& e : h->xfer_evs)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(CFDL_ERR_CUDA, "cudaEventCreate failed");
  }
  if (stage_doubles > h->xstage_len) {  // grows with the largest field list seen (an outgrown area is freed with the handle)
    h->xstage = nullptr;
    h->xstage_len = 0;
    int rc = dev_zero(h, h->xstage, stage_doubles);
    if (rc) return rc;
    h->xstage_len = stage_doubles;
  }
  return CFDL_OK;
}

// The transfers of one cfdl_step_host call.  Reference numbering: every field of the call has a slot of
// its own in the staging area; uploads are queued back to back on the upload stream (slot <- host) and
// each is followed on the compute stream by its permutation into the device numbering, so the copy engine
// never waits for a permutation kernel; downloads are the mirror image on the download stream.
// Partition-local arrays need no permutation and move straight between the host and the field.
struct HostStep : StepHook {
  Handle* h = nullptr;
  bool local = false;
  size_t off[CFDL_F_COUNT];
  const double* in[CFDL_F_COUNT];
  double* out[CFDL_F_COUNT];
  HostStep() { for (int f = 0; f < CFDL_F_COUNT; ++f) { off[f] = 0; in[f] = nullptr; out[f] = nullptr; } }
  cudaEvent_t ev_in(int f) const { return h->xfer_evs[f]; }
  cudaEvent_t ev_out(int f) const { return h->xfer_evs[CFDL_F_COUNT + f]; }
  double* slot(int f) const { return h->xstage + off[f]; }
  size_t count(int f) const { return local ? field_len(h, f) : host_len(h, f); }
  static bool staged(int f) { return f != CFDL_F_ANB; }  // anb changes layout (CSR <-> ELL): the plain upload / download path

  // host -> staging slot (or straight into the field) on the upload stream
  int queue_upload(int f) {
    if (!local && !staged(f)) return CFDL_OK;
    CFDL_CUDA(cudaMemcpyAsync(local ? h->fld[f] : slot(f), in[f], sizeof(double) * count(f), cudaMemcpyHostToDevice, h->xfer_in));
    CFDL_CUDA(cudaEventRecord(ev_in(f), h->xfer_in));
    return CFDL_OK;
  }
  // compute stream: wait for the field's upload, bring it into the device numbering
  int land(int f) {
    if (f == CFDL_F_AP || f == CFDL_F_ANB) h->pc_sumap_ok = false;
    if (!local && !staged(f)) return upload_field(h, f, in[f]);
    CFDL_CUDA(cudaStreamWaitEvent(h->stream, ev_in(f), 0));
    if (local) return CFDL_OK;
    const FieldMap m = field_map(h, f);
    return k_gather(h, h->fld[f], slot(f), m.map, m.n, m.ncomp);
  }
  // compute stream: field is final -> permute into its slot; download stream: copy out once that is done
  int send(int f) {
    if (!out[f]) return CFDL_OK;
    if (!local && !staged(f)) return download_field(h, f, out[f]);
    if (!local) {
      const FieldMap m = field_map(h, f);
      int rc = k_scatter(h, slot(f), h->fld[f], m.map, m.n, m.ncomp);
      if (rc) return rc;
    }
    CFDL_CUDA(cudaEventRecord(ev_out(f), h->stream));
    CFDL_CUDA(cudaStreamWaitEvent(h->xfer_out, ev_out(f), 0));
    CFDL_CUDA(cudaMemcpyAsync(out[f], local ? h->fld[f] : slot(f), sizeof(double) * count(f), cudaMemcpyDeviceToHost, h->xfer_out));
    return CFDL_OK;
  }
  static bool early_out(int f) { return (f >= CFDL_F_U && f <= CFDL_F_W) || (f >= CFDL_F_GU && f <= CFDL_F_GW); }
  int at(int stage) override {
    int rc = CFDL_OK;
    if (stage == STEP_MOMENTUM_DONE) for (int f = CFDL_F_U; f <= CFDL_F_W && !rc; ++f) rc = send(f);
    if (stage == STEP_GRAD_DONE) for (int f = CFDL_F_GU; f <= CFDL_F_GW && !rc; ++f) rc = send(f);
    if (stage == STEP_BEFORE_MIP && in[CFDL_F_MIP0]) rc = land(CFDL_F_MIP0);
    return rc;
  }
};

}  // namespace

extern "C" int cfdl_step_host(cfdl_handle h, double dt, int32_t nit, int32_t apply_bcs, int32_t local_numbering, int32_t n_in,
                              const int32_t* in_fields, const double* const* in_ptrs, int32_t n_out, const int32_t* out_fields,
                              double* const* out_ptrs, double* hist) {
  ENTER(h);
  if (n_in < 0 || n_out < 0 || (n_in && (!in_fields || !in_ptrs)) || (n_out && (!out_fields || !out_ptrs)))
    return fail(CFDL_ERR_ARG, "cfdl_step_host: bad field lists");
  HostStep st;
  st.h = h;
  st.local = local_numbering != 0;
  if (!st.local && h->prep.nranks > 1)
    return fail(CFDL_ERR_UNSUPPORTED, "cfdl_step_host: on a partitioned handle pass partition-local arrays (local_numbering = 1)");
  for (int i = 0; i < n_in; ++i) {
    if (in_fields[i] < 0 || in_fields[i] >= CFDL_F_COUNT || !in_ptrs[i] || st.in[in_fields[i]]) return fail(CFDL_ERR_ARG, "cfdl_step_host: bad input %d", i);
    st.in[in_fields[i]] = in_ptrs[i];
  }
  for (int i = 0; i < n_out; ++i) {
    if (out_fields[i] < 0 || out_fields[i] >= CFDL_F_COUNT || !out_ptrs[i] || st.out[out_fields[i]]) return fail(CFDL_ERR_ARG, "cfdl_step_host: bad output %d", i);
    st.out[out_fields[i]] = out_ptrs[i];
  }
  size_t need = 0;
  for (int f = 0; f < CFDL_F_COUNT && !st.local; ++f)
    if ((st.in[f] || st.out[f]) && HostStep::staged(f)) { st.off[f] = need; need += (host_len(h, f) + 31) / 32 * 32; }
  int rc;
  if ((rc = ensure_xfer(h, need))) return rc;
  // uploads in the caller's order, mip0 (first read by calc_mip) last: it travels while the momentum
  // equations are assembled and solved
  for (int i = 0; i < n_in && !rc; ++i) if (in_fields[i] != CFDL_F_MIP0) rc = st.queue_upload(in_fields[i]);
  if (!rc && st.in[CFDL_F_MIP0]) rc = st.queue_upload(CFDL_F_MIP0);
  for (int i = 0; i < n_in && !rc; ++i) if (in_fields[i] != CFDL_F_MIP0) rc = st.land(in_fields[i]);
  if (!rc && apply_bcs) rc = k_update_boundaries(h);
  double hs[16];
  if (!rc) rc = solve_uvwp_impl(h, dt, nit, hs, &st);
  // what the iteration finishes last (p, gp, gpc, mip, ...) goes out now; u,v,w,gu,gv,gw left during the pc solve
  for (int i = 0; i < n_out && !rc; ++i) if (!HostStep::early_out(out_fields[i])) rc = st.send(out_fields[i]);
  // all three streams are drained before the call returns, also on failure
  cudaError_t e1 = cudaStreamSynchronize(h->xfer_in), e2 = cudaStreamSynchronize(h->xfer_out), e3 = cudaStreamSynchronize(h->stream);
  if (rc) return rc;
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) return fail(CFDL_ERR_CUDA, "cfdl_step_host: %s", cudaGetErrorString(cudaGetLastError()));
  if (hist) std::memcpy(hist, hs, sizeof hs);
  return CFDL_OK;
}

// ---- checkpoint / restart ---------------------------------------------------------------------------
// The reference has no restart.  The state that carries over from one SIMPLE iteration to
// the next is u,v,w,p,u0,v0,w0,gu,gv,gw,gp,mip,mip0 (pc, gpc, d, dc, the matrix and the right-hand
// sides are rebuilt by every solve_uvwp): written in the reference numbering on one GPU (the file can
// be read back by any single-GPU handle of the same mesh), partition-local on several (one file per
// rank, same partition on restart).  Continuing from a checkpoint reproduces the uninterrupted run bit
// for bit (test).
namespace {
const char kCkpMagic[8] = {'C', 'F', 'D', 'L', 'C', 'K', 'P', '2'};
const int kCkpFields[13] = {CFDL_F_U, CFDL_F_V, CFDL_F_W, CFDL_F_P, CFDL_F_U0, CFDL_F_V0, CFDL_F_W0,
                            CFDL_F_GU, CFDL_F_GV, CFDL_F_GW, CFDL_F_GP, CFDL_F_MIP, CFDL_F_MIP0};
struct CkpHeader { char magic[8]; int64_t ne, nbf, nf; int32_t rank, nranks, nfields, pad; uint64_t mesh; };
// Identifies the mesh (and, on a partitioned handle, this rank's part of it) beyond its counts: every face
// contributes a hash of (reference face id, reference ids of its two sides); the sum does not depend on the
// device numbering, so a file written with one reordering mode is accepted by a handle created with another.
uint64_t mesh_fingerprint(const Handle* h) {
  const Prep& P = h->prep;
  auto orig = [&](int32_t d) -> uint64_t { return d < P.Nc ? (uint64_t)P.c2o[d] : (uint64_t)P.gN + (uint64_t)P.h2o[d - P.Nc]; };
  auto mix = [](uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; };
  uint64_t s = mix((uint64_t)P.gN * 31u + (uint64_t)P.rank) ^ mix((uint64_t)P.nranks);
  for (int32_t f = 0; f < P.F; ++f)
    s += mix(mix((uint64_t)P.f2o[f] + 1) ^ (mix(orig(P.face_a[f]) + 0x9e3779b97f4a7c15ULL) + 3 * mix(orig(P.face_b[f]) + 0x7f4a7c15ULL)));
  return s;
}
std::string ckp_path(const Handle* h, const char* path) {
  std::string p(path);
  if (h->prep.nranks > 1) p += ".r" + std::to_string(h->prep.rank) + "of" + std::to_string(h->prep.nranks);
  return p;
}
}  // namespace

extern "C" int cfdl_checkpoint_write(cfdl_handle h, const char* path) {
  ENTER(h);
  if (!path) return fail(CFDL_ERR_ARG, "cfdl_checkpoint_write: NULL path");
  const bool local = h->prep.nranks > 1;
  const std::string p = ckp_path(h, path);
  std::FILE* f = std::fopen(p.c_str(), "wb");
  if (!f) return fail(CFDL_ERR_ARG, "cfdl_checkpoint_write: cannot open %s", p.c_str());
  CkpHeader hd;
  std::memset(&hd, 0, sizeof hd);
  std::memcpy(hd.magic, kCkpMagic, 8);
  hd.ne = h->prep.gN; hd.nbf = h->prep.gB; hd.nf = h->prep.gF; hd.rank = h->prep.rank; hd.nranks = h->prep.nranks; hd.nfields = 13; hd.mesh = mesh_fingerprint(h);
  bool ok = std::fwrite(&hd, sizeof hd, 1, f) == 1;
  std::vector<double> buf;
  int rc = CFDL_OK;
  for (int i = 0; i < 13 && ok && !rc; ++i) {
    const int fld = kCkpFields[i];
    const int64_t n = (int64_t)(local ? field_len(h, fld) : host_len(h, fld));
    buf.resize((size_t)n);
    if (local) {
      if (cudaMemcpyAsync(buf.data(), h->fld[fld], sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
          cudaStreamSynchronize(h->stream) != cudaSuccess)
        rc = fail(CFDL_ERR_CUDA, "cfdl_checkpoint_write: download failed");
    } else {
      rc = download_field(h, fld, buf.data());
    }
    const int32_t id = fld;
    ok = !rc && std::fwrite(&id, 4, 1, f) == 1 && std::fwrite(&n, 8, 1, f) == 1 && std::fwrite(buf.data(), 8, (size_t)n, f) == (size_t)n;
  }
  ok = (std::fclose(f) == 0) && ok;
  if (rc) return rc;
  if (!ok) return fail(CFDL_ERR_INTERNAL, "cfdl_checkpoint_write: short write to %s", p.c_str());
  return CFDL_OK;
}

extern "C" int cfdl_checkpoint_read(cfdl_handle h, const char* path) {
  ENTER(h);
  if (!path) return fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: NULL path");
  const bool local = h->prep.nranks > 1;
  const std::string p = ckp_path(h, path);
  std::FILE* f = std::fopen(p.c_str(), "rb");
  if (!f) return fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: cannot open %s", p.c_str());
  CkpHeader hd;
  int rc = CFDL_OK;
  if (std::fread(&hd, sizeof hd, 1, f) != 1 || std::memcmp(hd.magic, kCkpMagic, 8) != 0) rc = fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: %s is not a checkpoint", p.c_str());
  else if (hd.ne != h->prep.gN || hd.nbf != h->prep.gB || hd.nf != h->prep.gF || hd.rank != h->prep.rank || hd.nranks != h->prep.nranks ||
           hd.mesh != mesh_fingerprint(h))
    rc = fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: %s belongs to another mesh or partition (ne %lld nbf %lld nf %lld, rank %d of %d, mesh id %016llx)", p.c_str(),
              (long long)hd.ne, (long long)hd.nbf, (long long)hd.nf, hd.rank, hd.nranks, (unsigned long long)hd.mesh);
  else if (hd.nfields != 13)
    rc = fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: %s holds %d fields, a checkpoint has 13", p.c_str(), hd.nfields);
  // the whole file is read and checked before the first upload: a short or damaged file leaves the handle as it was
  std::vector<std::vector<double>> buf(13);
  for (int i = 0; i < 13 && !rc; ++i) {
    int32_t id = -1;
    int64_t n = -1;
    if (std::fread(&id, 4, 1, f) != 1 || std::fread(&n, 8, 1, f) != 1) { rc = fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: file ends before record %d", i); break; }
    if (id != kCkpFields[i]) { rc = fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: record %d holds field %d, expected %d", i, id, kCkpFields[i]); break; }
    const int64_t want = (int64_t)(local ? field_len(h, id) : host_len(h, id));
    if (n != want) { rc = fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: field %d has %lld entries, expected %lld", id, (long long)n, (long long)want); break; }
    buf[i].resize((size_t)n);
    if (std::fread(buf[i].data(), 8, (size_t)n, f) != (size_t)n) { rc = fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: file ends inside field %d", id); break; }
  }
  if (!rc && std::fgetc(f) != EOF) rc = fail(CFDL_ERR_ARG, "cfdl_checkpoint_read: %s has data after the last field", p.c_str());
  std::fclose(f);
  for (int i = 0; i < 13 && !rc; ++i) {
    const int id = kCkpFields[i];
    if (local) {
      if (cudaMemcpyAsync(h->fld[id], buf[i].data(), sizeof(double) * buf[i].size(), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
        rc = fail(CFDL_ERR_CUDA, "cfdl_checkpoint_read: upload failed");
    } else {
      rc = upload_field(h, id, buf[i].data());
    }
  }
  if (cudaStreamSynchronize(h->stream) != cudaSuccess && !rc) rc = fail(CFDL_ERR_CUDA, "cfdl_checkpoint_read: upload failed");
  return rc;
}

#include "vtu.inc"
