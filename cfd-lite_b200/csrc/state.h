// Device-resident state of one mesh (one GPU): SoA cell/face arrays in the device numbering
// chosen by prep.cpp, fields of uvwp_t/phys_t (src/equations/mod_uvwp.f90:6-16,
// src/modules/mod_physics.f90:30), and solver work space.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <functional>
#include <vector>
#include "cfdl_common.h"
#include "prep.h"

namespace cfdl {

#define CFDL_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t err__ = (call);                                                                  \
    if (err__ != cudaSuccess)                                                                    \
      return ::cfdl::fail(CFDL_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
  } while (0)

// per-solve control block, lives in device memory; kernels of an iteration return at once
// when done != 0 so a batch of iterations can be enqueued without a host round trip
struct SolveCtl {
  int it, done, nit, pad;
  double res_i, res_f, res_max, res_target;
  unsigned int ticket, pad2;
};

struct DevSchedule {
  int nlevels = 0, nblocks = 1, nlag = 0;
  int32_t *lvl_ptr = nullptr, *s2c = nullptr, *nbs = nullptr, *bpos = nullptr, *lag_src = nullptr;
  std::vector<int32_t> blk_ptr;
  int max_level_cells = 0;
};

// Peer-to-peer ghost exchange over NVLink (multi-GPU, one process per GPU).  The solution
// arrays whose ghost cells neighbours write into live in one cudaMalloc'ed slab that is
// exported with cudaIpcGetMemHandle; the slab starts with this header so that a peer can find
// the arrays, the flag words and the reduction slots of its owner.
struct P2PHeader {
  long long magic;
  int rank, nranks, N, Nc, H, ncolors, nnbr, pad;
  long long off_field[7];  // byte offsets of u, v, w, pc, the second solver array, and the second arrays of v and w (side-by-side momentum passes)
  long long off_flags;     // unsigned long long flags[64]: flags[r] = last push sequence completed by rank r
  long long off_red_val;   // double red_val[2][64][2]: (sum r^2, max) of rank r, two alternating sets
  long long off_red_seq;   // unsigned long long red_seq[2][64]
  long long off_red3_val;  // double red3_val[2][64][6]: (sum r^2, max) of u, v, w of rank r (side-by-side momentum passes)
  long long off_red3_seq;  // unsigned long long red3_seq[2][64]
  long long off_stage[2];  // double stage[G][3]: landing zones for staged ghost exchanges, two alternating
  long long off_xflag;     // unsigned long long xflag[64]: xflag[r] = last staged exchange rank r has delivered
  long long off_mail_val;  // double mail_val[2][64][2]: small all-to-all mailbox (all-reduce / broadcast), alternating sets
  long long off_mail_seq;  // unsigned long long mail_seq[2][64]
  int nbr_rank[64];
  int recv_ptr[64 * 4 + 1];
  // the pc solve as one persistent launch per rank with chunk-to-chunk synchronisation over NVLink (kernels_rbq.inc, two colours)
  int rbq_ok;              // 1: this rank can take part (two colours, launch geometry fits)
  int nred;                // cells of the first colour
  int color_if[2];         // leading interface cells per colour
  int rbq_L, rbq_Gc;       // rows per interior chunk, chunks per colour
  int rbq_grid, rbq_pad;   // CTAs of the launch (= chunks on small meshes; fewer: chunks dealt round-robin)
  int rbq_ifc;             // chunks 0..rbq_ifc-1 hold the interface rows: they publish their progress to every neighbour
  int rbq_Ls;              // rows per interface chunk (shorter, so that their pushes and flags travel while the interior chunks still work)
  long long off_rbq_r2;    // double2 r2[nred + 2 + G]: red {newest, mid}; ghost g at nred + 2 + g
  long long off_rbq_b[2];  // double b[N - nred + 2 + G] x 2: the two black buffers; ghost g at (N - nred) + 2 + g
  long long off_rbq_prog;  // unsigned long long prog[max(RBQ_PROG_STRIDE, rbq_Gc)]: progress words of the own chunks
  long long off_rbq_llr;   // {double value, u64 tag} x 2 per ghost: newest and mid value of a red ghost, each tagged with the pass that wrote it
  long long off_rbq_llb[2];// {double value, u64 tag} per ghost and black buffer
  long long off_mailv_val; // double mailv_val[2][64][MAILV_LEN]: vector all-reduce mailbox, alternating sets
  long long off_mailv_seq; // unsigned long long mailv_seq[2][64]
};
enum { RBQ_PROG_STRIDE = 1024, MAILV_LEN = 2056 };
struct P2P {
  bool connected = false;
  char* slab = nullptr;
  size_t slab_bytes = 0;
  P2PHeader hdr;                        // this rank's header (host copy)
  std::vector<char*> peer_base;         // per rank; own slab for self
  std::vector<P2PHeader> peer_hdr;
  std::vector<void*> opened;            // cudaIpcOpenMemHandle results to close
  unsigned long long epoch = 0;
  unsigned long long xseq = 0, mseq = 0;  // staged exchanges / mailbox rounds done (every rank performs the same sequence)
  unsigned int* ticket = nullptr;
  unsigned int* xticket = nullptr;
  unsigned long long vseq = 0;          // vector mailbox rounds done
  int* err = nullptr;                   // device word: a peer-to-peer wait ran into its time limit (checked at the host syncs)
};
enum { P2P_U = 0, P2P_V = 1, P2P_W = 2, P2P_PC = 3, P2P_WORK = 4, P2P_WORK_V = 5, P2P_WORK_W = 6 };

struct Handle {
  Prep prep;
  P2P p2p;
  int device = 0, num_sms = 0;
  cudaStream_t stream = nullptr;
  int32_t N = 0, G = 0, Nc = 0, F = 0, B = 0, H = 0, Z = 0, K = 0, Np = 0, Fi = 0;  // owned, ghost, owned+ghost cells; local faces, halos
  int solver_mode = CFDL_SOLVER_PARITY;
  std::vector<void*> allocs;  // everything cudaMalloc'ed, freed in destroy
  // mesh
  int32_t *ell_nb = nullptr, *ell_fs = nullptr, *face_a = nullptr, *face_b = nullptr;
  int32_t *halo_cell = nullptr, *halo_face = nullptr, *halo_bc = nullptr, *bc_kind = nullptr;
  int32_t *c2o = nullptr, *cellmap = nullptr, *f2o = nullptr, *row_ptr = nullptr;  // cellmap: device cell|halo -> global index (size H)
  uint8_t *nfc = nullptr, *halo_slot = nullptr, *ftouch = nullptr;
  int16_t* ell_nb16 = nullptr;   // prep.nb16 on the device (nullptr when the offsets do not fit or the mesh is partitioned)
  int32_t* loc_order = nullptr;  // owned cells in base (spatial) order: the "locality order" of the assembly kernel variants
  double *bc_uvw = nullptr, *xc = nullptr, *yc = nullptr, *zc = nullptr, *aip = nullptr, *rip = nullptr;
  double *vol = nullptr, *rho = nullptr, *mu = nullptr;
  // fields, indexed by CFDL_F_*
  double* fld[CFDL_F_COUNT] = {nullptr};
  // energy and scalar equations (kernels_transport.cu; fields allocated by cfdl_energy_init / cfdl_scalar_init)
  bool has_energy = false, has_scalar = false;
  double *tc = nullptr, *cp = nullptr;  // thermal conductivity, heat capacity per cell (Nc)
  double* s_bc = nullptr;               // scalar: Dirichlet value per boundary section
  double s_dcoef = 1.0, s_vel[3] = {0.0, 0.0, -100.0};
  // staging + solver work
  double* stage = nullptr;      // max(3H, Z, F) doubles for permuted host transfers
  size_t stage_len = 0;
  double* partial = nullptr;    // per-CTA reduction partials
  int partial_len = 0;
  SolveCtl* ctl = nullptr;      // device
  SolveCtl* ctl_host = nullptr; // pinned
  double* scal = nullptr;       // small device scalars (pref, ...)
  double* scal_host = nullptr;  // pinned mirror / readback area (64 doubles)
  unsigned int* barrier = nullptr;
  // parity (level-scheduled) solver work: sweep-space copies
  DevSchedule natural, blocks;
  double *ap_s = nullptr, *b_s = nullptr, *anb_s = nullptr, *phi_s = nullptr, *rr = nullptr, *rsig = nullptr;
  int coop_ctas = 0;
  double* rb_work = nullptr;   // second value array of the fused two-colour solver (H doubles)
  double *pcg_r = nullptr, *pcg_q = nullptr, *pcg_p = nullptr;  // conjugate-gradient work vectors (first use)
  double* pcg_z = nullptr;     // preconditioned residual (multicolour SSOR preconditioner)
  int pcg_precond = 1;         // 1: multicolour SSOR on two-colour single-GPU meshes, 0: Jacobi
  // the three momentum solves side by side (kernels_rb3.inc): second value arrays of v and w (u uses rb_work),
  // one control block per equation
  double* rb3_work[2] = {nullptr, nullptr};
  SolveCtl* ctl3 = nullptr;       // device, 3 blocks
  SolveCtl* ctl3_host = nullptr;  // pinned
  // the pc solve as one persistent launch with neighbour-only synchronisation (kernels_rbq.inc; one GPU, two colours)
  int rbq = 1;                    // 0: always pass by pass
  int rbq_refused = 0;            // the cooperative launch was refused or a wait timed out: pass by pass from then on
  int rbq_occ = 0, rbq_occ_for = -1;  // co-resident CTAs per SM of the rbq_kernel instantiation in use (occupancy calculator)
  int rbq_ctas_per_sm = 0;        // > 0: use fewer CTAs per SM than that
  double* rbq_mem = nullptr;      // red pairs, black buffers, per-pass partials, progress words
  size_t rbq_len = 0;
  unsigned long long* rbq_prog = nullptr;  // one GPU: progress words of the chunks + error word (fixed place, never reset)
  unsigned int* rbq_ticket = nullptr;  // large meshes: the counter the CTAs draw their next (pass, chunk) from
  size_t rbq_prog_len = 0;                 // words allocated (the error word sits behind them)
  int rbq_lmax = 0, rbq_lbig = 0, rbq_cap = 0;  // options (0: defaults / environment): largest one-chunk-per-CTA size, chunk size of the round-robin form, CTA limit
  int rbq_last_chunks = 0, rbq_last_grid = 0, rbq_last_L = 0;  // geometry of the last persistent pc solve (info)
  int rbq_prefetch = -1;                   // option: L2 prefetch of the next trips' constants in the persistent pc solve (trips ahead; 0 = off, -1 = by size)
  int rbq_counter = -1;                    // option: chunks handed out from a counter: -1 where one chunk per CTA would be too long, 0 never, 1 always
  double rbq_l2_fraction = 0.0;            // option: share of the L2 the value arrays may take for the persistent pc solve (0: default / environment)
  size_t l2_bytes = 0;
  // partitioned meshes (peer-to-peer mode): value arrays and progress words live in the exported slab; per chunk the progress
  // words to wait for (own chunks and the neighbours' interface chunks), the interface rows to push after a pass
  int rbq_dist_state = 0;         // 0: not set up yet, 1: ready, -1: refused (on every rank alike)
  int32_t* rbq_nbq = nullptr;     // K x Np: neighbour position in the other colour's value array (ghosts appended)
  int32_t *rbq_wait_ptr = nullptr, *rbq_wait_idx = nullptr;
  int32_t *rbq_land_ptr[2] = {nullptr, nullptr}, *rbq_land_g[2] = {nullptr, nullptr};
  int32_t *rbq_push_ptr[2] = {nullptr, nullptr}, *rbq_push_src[2] = {nullptr, nullptr}, *rbq_push_nbr[2] = {nullptr, nullptr}, *rbq_push_dst[2] = {nullptr, nullptr};
  unsigned long long rbq_epoch = 0;  // launches done (every rank performs the same sequence): progress words only ever grow
  int64_t prof_extra_passes = 0;  // passes executed a second time because the stopping rule fired inside a block
  int rb_idx16 = 1;               // pc passes read 16-bit neighbour offsets instead of 32-bit ids where prep.nb16 exists
  int pc_sumap = 1;               // fused pc passes rebuild ap as the slot-order sum of anb instead of reading it
  bool pc_sumap_ok = false;       // true while ap/anb on the device are what calc_coef_p wrote (cleared by any other writer)
  int uvw_fused = 1;              // 1: u, v, w side by side, 0: one after the other as the reference does (same bits)
  int last_passes[8] = {1, 1, 1, 1, 1, 1, 1, 1};  // per equation: passes the previous solve needed (first-batch size estimate)
  int fused_rb = 1;            // 0: always use one launch per colour + residual pass
  int use_p2p = 1;             // 0: NCCL send/recv even when peer slabs are connected
  int occ_grids = 1;           // assembly kernels: grid = resident CTAs (occupancy API) instead of 8 per SM
  int pdl_rows = 1;            // rows per thread prefetched to L2 ahead of the dependency wait
  int use_pdl = 1;             // fused passes: programmatic dependent launch (a pass loads its first matrix rows while the previous one drains)
  // per cell-cell face, owner orientation, computed once with the very expressions the reference
  // evaluates every iteration (area, unit normal, |dr|, projected |dr_p|, dr.n, both distance
  // weights, dr, dr_p): the assembly kernels then need 10 instead of 19 FP64 div/sqrt per face
  double *fs_area = nullptr, *fs_ds = nullptr, *fs_dsp = nullptr, *fs_dn = nullptr, *fs_wto = nullptr, *fs_wtn = nullptr;
  double *fs_rds = nullptr, *fs_rdsp = nullptr, *fs_rdn = nullptr;  // RN(1/ds), RN(1/dsp), RN(1/dn)
  // least-squares gradient statics (filled on first use by grad_variant 1): inverse LSQ matrix per cell
  // (9 x Np, entry-major) and the weight 1/|dr|^2 per slot (K x Np)
  double *lsq_binv = nullptr, *lsq_w = nullptr;
  int grad_variant = 1;        // calc_grad: 1 = on the LSQ statics (inverse matrix and weights precomputed), 0 = the reference's form
  double *fs_n[3] = {nullptr, nullptr, nullptr}, *fs_dr[3] = {nullptr, nullptr, nullptr}, *fs_drp[3] = {nullptr, nullptr, nullptr};
  int use_statics = 1;         // 0: recompute face geometry in every kernel (the reference's way)
  int mip_hoist = 0;           // calc_mip: connectivity of all slots loaded up front (two CTAs per SM) / per face (three)
  int tune_ctas = 8;           // CTAs per SM for the solver passes (grid = min(need, num_sms * tune_ctas))
  // cfdl_step_host: transfer streams that run beside the compute stream, and a staging area of
  // its own (reference-numbered copies of the late input and the early outputs)
  cudaStream_t xfer_in = nullptr, xfer_out = nullptr;
  std::vector<cudaEvent_t> xfer_evs;  // per field: upload landed in its slot | field permuted into its slot for download
  double* xstage = nullptr;
  size_t xstage_len = 0;
  // multi-GPU (one process per GPU): NCCL communicator and interface buffers
  void* comm = nullptr;            // ncclComm_t
  int nnbr = 0;
  int32_t* send_cells = nullptr;   // device copy of prep.send_cells
  int32_t *tgt_ptr = nullptr, *tgt_nbr = nullptr, *tgt_pos = nullptr;  // device copies of prep.tgt_*
  double* send_buf = nullptr;      // 9 doubles per send cell
  double* recv_buf = nullptr;      // 9 doubles per ghost (multi-component receives)
  int64_t ne_global = 0;
  // instrumentation
  int64_t launches = 0;         // kernels launched since the last reset ("gpu_launches")
  int profile = 0;              // 1: bracket every launch of the profiled kernel kinds with CUDA events
  std::vector<cudaEvent_t> prof_ev;  // pairs (begin, end)
  std::vector<int> prof_kind, prof_cnt;  // per pair: kernel kind, launches it covers
  size_t prof_used = 0;
  double prof_ms[12] = {0};
  int64_t prof_n[12] = {0};
  cudaEvent_t timer_ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

// every launch site takes its stream from S(h), which also counts the launch
inline cudaStream_t S(Handle* h) { h->launches++; return h->stream; }

// profiled kernel kinds (per-launch CUDA-event timing when h->profile is on)
// PROF_RESIDUAL: residual_kernel (one right-hand side); PROF_RESIDUAL3: residual3_kernel (u, v, w) incl. its cross-rank reduction;
// PROF_GRAD: the fused u,v,w gradient pass; PROF_GRAD1: the single-field gradient (gpc)
enum { PROF_SGS_SWEEP = 0, PROF_RESIDUAL = 1, PROF_COEF_UVW = 2, PROF_COEF_P = 3, PROF_MIP = 4, PROF_GRAD = 5, PROF_LEVELS = 6, PROF_PCG = 7, PROF_SGS3 = 8,
       PROF_RESIDUAL3 = 9, PROF_GRAD1 = 10, PROF_KINDS = 11 };
int prof_begin(Handle* h, int kind, int level = 1);
int prof_end(Handle* h, int count = 1, int level = 1);
int prof_collect(Handle* h);  // after a stream sync: fold finished event pairs into prof_ms / prof_n
// launch geometry helper: grids are sized as a multiple of the SM count (B200: 148)
inline int grid_for(const Handle* h, int64_t n, int threads, int ctas_per_sm = 8) {
  int64_t need = (n + threads - 1) / threads;
  int64_t cap = (int64_t)h->num_sms * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// grid for a grid-stride kernel: exactly the CTAs that are resident at once (one wave), as the
// occupancy calculator reports for this kernel — a fixed 8 per SM leaves register-heavy assembly
// kernels (4-5 resident CTAs) with a second, partly filled wave
template <auto Kern>
inline int occ_grid(const Handle* h, int64_t n, int threads) {
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, Kern, threads, 0) != cudaSuccess || o < 1) o = 8;
    occ = o;
  }
  return grid_for(h, n, threads, h->occ_grids ? occ : 8);
}

// ---- kernels_assembly.cu
int k_face_statics(Handle* h);  // fills fs_* (arrays must be allocated, Fi doubles each)
// kernels_statics.cu: the same routines on the precomputed face statics
int k_calc_coef_uvw_statics(Handle* h, double dt);
int k_calc_coef_p_statics(Handle* h);
int k_calc_mip_statics(Handle* h, bool rhie_chow, double dt);
int k_correct_faces_statics(Handle* h);
int k_update_boundaries(Handle* h);
int k_calc_coef_uvw(Handle* h, double dt);
int k_calc_mip(Handle* h, bool rhie_chow, double dt);
int k_calc_coef_p(Handle* h);
int k_adjust_pc(Handle* h);
int k_update_uvwp(Handle* h);
int k_calc_grad(Handle* h, const double* phi, double* grad);
int k_calc_grad3(Handle* h);  // gu,gv,gw from u,v,w in one pass
int k_update_time(Handle* h);
// kernels_transport.cu: energy (mod_energy.f90) and scalar (mod_scalar.f90) equations
int k_energy_init(Handle* h, const double* tc_host, const double* cp_host);
int k_scalar_init(Handle* h, double dcoef, const double vel[3], const double* bc_value_host);
int k_transport_boundaries(Handle* h);
int k_transport_update_time(Handle* h);
int k_solve_energy(Handle* h, double dt, int nit, double* out4);
int k_solve_scalar(Handle* h, double dt, int nit, double* out4);
int k_gather(Handle* h, double* dst, const double* src, const int32_t* map, int64_t n, int ncomp);
int k_scatter(Handle* h, double* dst, const double* src, const int32_t* map, int64_t n, int ncomp);
int k_csr_to_ell(Handle* h, double* ell, const double* csr);
int k_ell_to_csr(Handle* h, double* csr, const double* ell);
// ---- kernels_solver.cu
int solver_init(Handle* h);
// dispatch=false: solve_gs (mod_solver.f90:255); dispatch=true: solve() (:329), i.e. the block
// solver when the handle has n_subdomains>1
int solve_equation(Handle* h, int eq, double* phi, const double* rhs, int nit, double* out4, bool dispatch);
int solve_momentum(Handle* h, int nit, double* out12);  // u, v, w (side by side where possible)

int residual_plain(Handle* h, const double* phi, const double* rhs, bool signed_max, double* res, double* res_max);
// ---- comm.cu (no-ops on a single rank)
// refresh the ghost-cell entries of a device field from their owner ranks; ncomp doubles per
// cell (AoS); color >= 0 restricts the exchange to ghosts/sends of that colour
int comm_exchange(Handle* h, double* field, int ncomp, int color);
int comm_exchange_multi(Handle* h, double* const* fields, int nf, int ncomp);  // all colours, up to three fields of equal width at once
int comm_allreduce_sum_max(Handle* h, double* dev2);  // dev2[0] summed, dev2[1] maxed over ranks
// peer-to-peer path (after cfdl_comm_ipc_connect): write the colour-`color` interface values of
// up to two arrays straight into the neighbours' ghost cells, then publish sequence number
// `seq` in their flag words; with reduce_parity >= 0 also combine (sum r^2, max) in local_sm
// over all ranks through the peers' reduction slots and advance ctl like finalize_residual_kernel
int p2p_alloc_slab(Handle* h);  // carves u,v,w,pc and rb_work out of one exportable allocation
int p2p_push(Handle* h, int color, const double* a, const double* b, unsigned long long seq, int reduce_parity, const double* local_sm);
struct P2PWait { const unsigned long long* flags; unsigned long long expect; int n; int r[8]; int* err; };
P2PWait p2p_wait_args(Handle* h, unsigned long long expect);
// remote-store description of one launch (kernel parameter; tptr == nullptr: no remote stores):
// cell c mirrors into ghost slot d0[i] + tpos[e] of neighbour i = tnbr[e], e in [tptr[c], tptr[c+1])
struct P2PStore {
  const int32_t *tptr, *tnbr, *tpos;
  double* dst_a[8];
  double* dst_b[8];
  int d0[8];
  unsigned long long* peer_flag[8];
  int nnbr;
  unsigned long long seq;
  unsigned int* ticket;
};
// cross-rank reduction of (sum r^2, max) through peer memory (on == 0: single rank / NCCL mode)
struct P2PReduce {
  int on, nranks, rank, parity;
  unsigned long long seq;
  double ne_global;
  double* peer_val[64];
  unsigned long long* peer_seq[64];
  const double* my_val;
  const unsigned long long* my_seq;
  int* err;
};
// Every wait on a word another GPU writes is time-limited (~10 s of SM clock, far beyond any legitimate wait): the
// waiter that gives up raises the handle's error word, every later wait returns at once, the kernels run to their end
// and the next API call that synchronises reports CFDL_ERR_COMM (p2p_check) — a dead rank is an error code, not a hang.
#ifdef CUEMU
#define CFDL_SPIN_LIMIT 4000000ll      // host emulation: 4 s (its clock64 counts microseconds) so that the dead-rank test stays short
#else
#define CFDL_SPIN_LIMIT 20000000000ll
#endif
#ifdef CUEMU
#define CFDL_SPIN_PAUSE CUEMU_SPIN  // host emulation (tests/emul): a waiting CTA yields its core
#else
#define CFDL_SPIN_PAUSE
#endif
#if defined(__CUDACC__) || defined(CUEMU)
__device__ __forceinline__ bool spin_until_ge(const volatile unsigned long long* w, unsigned long long need, int* err) {
  const long long t0 = clock64();
  while (*w < need) {
    CFDL_SPIN_PAUSE;
    if (err && *((volatile const int*)err)) return false;
    if (clock64() - t0 > CFDL_SPIN_LIMIT) { if (err) *((volatile int*)err) = 1; return false; }
  }
  return true;
}
#endif
int p2p_check(Handle* h);  // after a stream sync point: CFDL_ERR_COMM when a peer-to-peer wait has timed out on this handle
// element-wise all-reduce of npairs (sum, max) pairs in device memory over the ranks, summed in rank order (peer-to-peer mailbox)
int comm_allreduce_pairs(Handle* h, double* dev, int npairs);
// launch geometry of the persistent pc solve for a rank with nred first-colour cells of N: rows per chunk, chunks
// (partitioned: if0 / if1 leading interface rows per colour -> Ls rows per interface chunk, ifc such chunks in front)
int rbq_plan(const Handle* h, int nred, int N, int K, bool dist, int if0, int if1, int* L, int* Gc, int* Ls, int* ifc, int* grid);
int p2p_store_args(Handle* h, int color, const double* a, const double* b, unsigned long long seq, P2PStore* out);
int p2p_reduce_args(Handle* h, int parity, unsigned long long seq, P2PReduce* out);
int p2p_reduce3_args(Handle* h, int parity, unsigned long long seq, P2PReduce* out);  // slots of 6 doubles (u, v, w)
int comm_bcast(Handle* h, double* dev, int count, int root);
void comm_destroy(Handle* h);

}  // namespace cfdl

struct cfdl_handle_s : public cfdl::Handle {};

#include <functional>
namespace cfdl {
// where create_from_prep takes geometry from: indices are 0-based ids of the GLOBAL mesh
// (cells 0..gN-1, halo j as gN + j, faces 0..gF-1)
struct GeomSource {
  std::function<void(int32_t, double*)> cell_xyz;
  std::function<double(int32_t)> vol, rho, mu;
  std::function<void(int32_t, double*, double*)> face;  // area vector (owner -> neighbour) and centroid
};
int create_from_prep(cfdl_handle_s* h, const GeomSource& G, cfdl_handle* out);
}  // namespace cfdl
