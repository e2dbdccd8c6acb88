// Multi-GPU plumbing: one process per GPU, one partition per process.  Ghost-cell values move
// by NCCL send/recv over NVLink (grouped per exchange), residual scalars by NCCL all-reduce,
// the reference pressure pc(1) by NCCL broadcast.  NCCL is dlopen'ed (libnccl.so.2 — the copy
// torch already loaded when the host is Python), so the library loads on boxes without NCCL
// and single-GPU use never touches it.  This is the GPU analogue of update_halos
// (src/modules/mod_subdomains.f90:191-212) and of the residual accumulation loop of
// multi_subdomain_solver (src/modules/mod_solver.f90:144-150,172-178).
#include <dlfcn.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "state.h"

namespace cfdl {
namespace {

// minimal NCCL ABI (nccl.h 2.x): opaque comm, 128-byte unique id, enums by value
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { kNcclDouble = 8, kNcclSum = 0, kNcclMax = 2 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib) return CFDL_OK;
  const char* cand[] = {std::getenv("CFDL_NCCL_PATH"), "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* c : cand) {
    if (!c || !*c) continue;
    lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return fail(CFDL_ERR_NCCL, "cannot dlopen libnccl.so.2 (set CFDL_NCCL_PATH): %s", dlerror());
#define SYM(field, name)                                                    \
  *(void**)(&g_nccl.field) = dlsym(lib, name);                              \
  if (!g_nccl.field) return fail(CFDL_ERR_NCCL, "libnccl lacks symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Broadcast, "ncclBroadcast");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = lib;
  return CFDL_OK;
}

#define CFDL_NCCL(call)                                                                      \
  do {                                                                                       \
    int r__ = (call);                                                                        \
    if (r__ != 0) return fail(CFDL_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
  } while (0)

// sendbuf[i*ncomp + q] = field[cells[i]*ncomp + q]
__global__ void __launch_bounds__(256) pack_kernel(double* __restrict__ buf, const double* __restrict__ field,
                                                   const int32_t* __restrict__ cells, int i0, int i1, int ncomp) {
  for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += gridDim.x * blockDim.x)
    for (int q = 0; q < ncomp; ++q) buf[(size_t)i * ncomp + q] = field[(size_t)cells[i] * ncomp + q];
}

}  // namespace

static int p2p_exchange(Handle* h, double* const* fields, int nf, int ncomp, int color = -1);
static int p2p_mail(Handle* h, double* dev, int mode, int root);

int comm_exchange(Handle* h, double* field, int ncomp, int color) {
  if (h->prep.nranks == 1 || h->nnbr == 0) return CFDL_OK;
  if (ncomp <= 3 && h->p2p.connected && h->use_p2p) return p2p_exchange(h, &field, 1, ncomp, color);  // all colours (-1) or one
  if (!h->comm) return fail(CFDL_ERR_NCCL, "this handle is one partition of %d: call cfdl_comm_init before computing", h->prep.nranks);
  if (ncomp < 1 || ncomp > 9) return fail(CFDL_ERR_ARG, "comm_exchange: ncomp %d", ncomp);
  const Prep& p = h->prep;
  const int nc = p.ncolors;
  // pack every needed slice with one launch per neighbour (colour slices of a neighbour are adjacent)
  for (int r = 0; r < h->nnbr; ++r) {
    const int s0 = p.send_ptr[(size_t)r * nc + (color >= 0 ? color : 0)], s1 = p.send_ptr[(size_t)r * nc + (color >= 0 ? color + 1 : nc)];
    if (s1 > s0) pack_kernel<<<grid_for(h, s1 - s0, 256, 2), 256, 0, S(h)>>>(h->send_buf, field, h->send_cells, s0, s1, ncomp);
  }
  CFDL_CUDA(cudaGetLastError());
  CFDL_NCCL(g_nccl.GroupStart());
  for (int r = 0; r < h->nnbr; ++r) {
    const int peer = p.nbr_rank[r];
    const int s0 = p.send_ptr[(size_t)r * nc + (color >= 0 ? color : 0)], s1 = p.send_ptr[(size_t)r * nc + (color >= 0 ? color + 1 : nc)];
    const int g0 = p.recv_ptr[(size_t)r * nc + (color >= 0 ? color : 0)], g1 = p.recv_ptr[(size_t)r * nc + (color >= 0 ? color + 1 : nc)];
    // ghosts of one (neighbour, colour) are contiguous in the field, so receives land in place
    if (s1 > s0) CFDL_NCCL(g_nccl.Send(h->send_buf + (size_t)s0 * ncomp, (size_t)(s1 - s0) * ncomp, kNcclDouble, peer, h->comm, h->stream));
    if (g1 > g0) CFDL_NCCL(g_nccl.Recv(field + ((size_t)h->N + g0) * ncomp, (size_t)(g1 - g0) * ncomp, kNcclDouble, peer, h->comm, h->stream));
  }
  CFDL_NCCL(g_nccl.GroupEnd());
  return CFDL_OK;
}

int comm_allreduce_sum_max(Handle* h, double* dev2) {
  if (h->prep.nranks == 1) return CFDL_OK;
  if (h->p2p.connected && h->use_p2p) return p2p_mail(h, dev2, /*sum,max*/ 0, 0);
  if (!h->comm) return fail(CFDL_ERR_NCCL, "cfdl_comm_init has not been called");
  CFDL_NCCL(g_nccl.GroupStart());
  CFDL_NCCL(g_nccl.AllReduce(dev2, dev2, 1, kNcclDouble, kNcclSum, h->comm, h->stream));
  CFDL_NCCL(g_nccl.AllReduce(dev2 + 1, dev2 + 1, 1, kNcclDouble, kNcclMax, h->comm, h->stream));
  CFDL_NCCL(g_nccl.GroupEnd());
  return CFDL_OK;
}

int comm_bcast(Handle* h, double* dev, int count, int root) {
  if (h->prep.nranks == 1) return CFDL_OK;
  if (count == 1 && h->p2p.connected && h->use_p2p) return p2p_mail(h, dev, /*pick root*/ 1, root);
  if (!h->comm) return fail(CFDL_ERR_NCCL, "cfdl_comm_init has not been called");
  CFDL_NCCL(g_nccl.Broadcast(dev, dev, (size_t)count, kNcclDouble, root, h->comm, h->stream));
  return CFDL_OK;
}

void comm_destroy(Handle* h) {
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr;
  for (void* p : h->p2p.opened) cudaIpcCloseMemHandle(p);
  h->p2p.opened.clear();
  h->p2p.connected = false;
}

// ---------------------------------------------------------------------------------------------
// Peer-to-peer ghost exchange: the sender's kernel stores interface values directly into the
// receiver's ghost cells over NVLink and then raises a flag word in the receiver's memory; the
// receiver's next pass spins on that flag before it gathers.  No host round trip, no NCCL call
// inside a solver iteration; the residual norm is combined through per-rank slots the same way.
namespace {

constexpr long long kMagic = 0x4346444C50325031ll;  // "CFDLP2P1"
inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct PushArgs {
  int nnbr, narr, do_reduce, nranks, rank, parity;
  const int32_t* send_cells;
  const double* src[2];
  double* dst[8][2];
  int s0[8], cnt[8], d0[8];
  unsigned long long* peer_flag[8];
  unsigned long long seq;
  unsigned int* ticket;
  SolveCtl* ctl;
  double ne_global;
  const double* local_sm;
  double* peer_red_val[64];
  unsigned long long* peer_red_seq[64];
  const double* my_red_val;
  const unsigned long long* my_red_seq;
  int* err;
};

__global__ void __launch_bounds__(256) p2p_push_kernel(const __grid_constant__ PushArgs A) {
  if (A.ctl->done) return;
  int total = 0;
  for (int i = 0; i < A.nnbr; ++i) total += A.cnt[i];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int i = 0, k = t;
    while (k >= A.cnt[i]) { k -= A.cnt[i]; ++i; }
    const int cell = A.send_cells[A.s0[i] + k];
    for (int a = 0; a < A.narr; ++a) A.dst[i][a][A.d0[i] + k] = A.src[a][cell];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {  // every CTA's stores are fenced: publish "push `seq` complete" to the neighbours
    __threadfence_system();
    if (threadIdx.x < A.nnbr) *((volatile unsigned long long*)A.peer_flag[threadIdx.x]) = A.seq;
    if (threadIdx.x == 0) *A.ticket = 0;
  }
  if (blockIdx.x != 0 || !A.do_reduce) return;
  // all-reduce of (sum r^2, max) through the peers' slots; summed in rank order on every rank
  const int slot = A.parity * 64;
  if (threadIdx.x < A.nranks) {
    const int r = threadIdx.x;
    A.peer_red_val[r][(slot + A.rank) * 2] = A.local_sm[0];
    A.peer_red_val[r][(slot + A.rank) * 2 + 1] = A.local_sm[1];
    __threadfence_system();
    *((volatile unsigned long long*)&A.peer_red_seq[r][slot + A.rank]) = A.seq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, m = 0.0;
    for (int r = 0; r < A.nranks; ++r) spin_until_ge(&A.my_red_seq[slot + r], A.seq, A.err);
    __threadfence_system();
    for (int r = 0; r < A.nranks; ++r) {
      s += *((volatile const double*)&A.my_red_val[(slot + r) * 2]);
      m = fmax(m, *((volatile const double*)&A.my_red_val[(slot + r) * 2 + 1]));
    }
    const double res = sqrt(s / A.ne_global);
    A.ctl->it += 1;
    A.ctl->res_f = res; A.ctl->res_max = m;
    A.ctl->done = !(A.ctl->it < A.ctl->nit && res > A.ctl->res_target);
  }
}

int field_id(const Handle* h, const double* p) {
  if (p == h->fld[CFDL_F_U]) return P2P_U;
  if (p == h->fld[CFDL_F_V]) return P2P_V;
  if (p == h->fld[CFDL_F_W]) return P2P_W;
  if (p == h->fld[CFDL_F_PC]) return P2P_PC;
  if (p == h->rb_work) return P2P_WORK;
  if (p == h->rb3_work[0]) return P2P_WORK_V;
  if (p == h->rb3_work[1]) return P2P_WORK_W;
  return -1;
}

}  // namespace

int p2p_alloc_slab(Handle* h) {
  P2P& q = h->p2p;
  const Prep& p = h->prep;
  if (p.ncolors > 4 || p.nbr_rank.size() > 8 || p.nranks > 64) return CFDL_OK;  // fall back to NCCL exchanges
  const size_t arr = align256(sizeof(double) * ((size_t)h->H + 32));
  size_t off = 4096;
  P2PHeader& hd = q.hdr;
  std::memset(&hd, 0, sizeof hd);
  hd.magic = kMagic; hd.rank = p.rank; hd.nranks = p.nranks; hd.N = h->N; hd.Nc = h->Nc; hd.H = h->H;
  hd.ncolors = p.ncolors; hd.nnbr = (int)p.nbr_rank.size();
  hd.off_flags = off; off += align256(64 * 8);
  hd.off_red_val = off; off += align256(2 * 64 * 2 * 8);
  hd.off_red_seq = off; off += align256(2 * 64 * 8);
  hd.off_red3_val = off; off += align256(2 * 64 * 6 * 8);
  hd.off_red3_seq = off; off += align256(2 * 64 * 8);
  const size_t off_ticket = off; off += 256;
  hd.off_xflag = off; off += align256(64 * 8);
  hd.off_mail_val = off; off += align256(2 * 64 * 2 * 8);
  hd.off_mail_seq = off; off += align256(2 * 64 * 8);
  const size_t off_err = off; off += 256;
  for (int b = 0; b < 2; ++b) { hd.off_stage[b] = (long long)off; off += align256(sizeof(double) * 9 * ((size_t)h->G + 32)); }
  for (int a = 0; a < 7; ++a) { hd.off_field[a] = (long long)off; off += arr; }
  hd.off_mailv_val = (long long)off; off += align256(sizeof(double) * 2 * 64 * MAILV_LEN);
  hd.off_mailv_seq = (long long)off; off += align256(2 * 64 * 8);
  // the persistent pc solve on a partitioned two-colour mesh (kernels_rbq.inc): its value arrays and progress words are
  // written by the neighbours, so they live in the exported slab; the launch geometry is fixed here so that peers can read it
  hd.rbq_ok = 0;
  if (p.ncolors == 2) {
    const int nred = p.color_ptr[1], nblack = h->N - nred;
    int L = 0, Gc = 0, Ls = 0, ifc = 0, grid = 0;
    if (nred > 0 && nblack > 0 && rbq_plan(h, nred, h->N, h->K, true, p.color_if[0], p.color_if[1], &L, &Gc, &Ls, &ifc, &grid) == CFDL_OK) {
      hd.rbq_ok = 1; hd.nred = nred; hd.color_if[0] = p.color_if[0]; hd.color_if[1] = p.color_if[1];
      hd.rbq_L = L; hd.rbq_Gc = Gc; hd.rbq_Ls = Ls; hd.rbq_ifc = ifc; hd.rbq_grid = grid;
      hd.off_rbq_r2 = (long long)off; off += align256(sizeof(double) * 2 * ((size_t)nred + 2 + h->G + 2));
      for (int b = 0; b < 2; ++b) { hd.off_rbq_b[b] = (long long)off; off += align256(sizeof(double) * ((size_t)nblack + 2 + h->G + 2)); }
      hd.off_rbq_prog = (long long)off; off += align256(sizeof(unsigned long long) * std::max((size_t)RBQ_PROG_STRIDE, (size_t)Gc));
      hd.off_rbq_llr = (long long)off; off += align256(32 * ((size_t)h->G + 2));
      for (int b = 0; b < 2; ++b) { hd.off_rbq_llb[b] = (long long)off; off += align256(16 * ((size_t)h->G + 2)); }
    }
  }
  for (int i = 0; i < hd.nnbr; ++i) hd.nbr_rank[i] = p.nbr_rank[i];
  for (size_t i = 0; i < p.recv_ptr.size(); ++i) hd.recv_ptr[i] = p.recv_ptr[i];
  CFDL_CUDA(cudaMalloc(&q.slab, off));
  h->allocs.push_back(q.slab);
  q.slab_bytes = off;
  CFDL_CUDA(cudaMemset(q.slab, 0, off));
  CFDL_CUDA(cudaMemcpy(q.slab, &hd, sizeof hd, cudaMemcpyHostToDevice));
  h->fld[CFDL_F_U] = (double*)(q.slab + hd.off_field[P2P_U]);
  h->fld[CFDL_F_V] = (double*)(q.slab + hd.off_field[P2P_V]);
  h->fld[CFDL_F_W] = (double*)(q.slab + hd.off_field[P2P_W]);
  h->fld[CFDL_F_PC] = (double*)(q.slab + hd.off_field[P2P_PC]);
  h->rb_work = (double*)(q.slab + hd.off_field[P2P_WORK]);
  h->rb3_work[0] = (double*)(q.slab + hd.off_field[P2P_WORK_V]);
  h->rb3_work[1] = (double*)(q.slab + hd.off_field[P2P_WORK_W]);
  q.ticket = (unsigned int*)(q.slab + off_ticket);
  q.xticket = q.ticket + 8;
  q.err = (int*)(q.slab + off_err);
  return CFDL_OK;
}

P2PWait p2p_wait_args(Handle* h, unsigned long long expect) {
  P2PWait w;
  w.flags = (const unsigned long long*)(h->p2p.slab + h->p2p.hdr.off_flags);
  w.expect = expect;
  w.n = h->nnbr;
  for (int i = 0; i < 8; ++i) w.r[i] = i < h->nnbr ? h->prep.nbr_rank[i] : 0;
  w.err = h->p2p.err;
  return w;
}

int p2p_store_args(Handle* h, int color, const double* a, const double* b, unsigned long long seq, P2PStore* out) {
  P2P& q = h->p2p;
  const Prep& p = h->prep;
  if (!q.connected) return fail(CFDL_ERR_INTERNAL, "p2p_store_args without cfdl_comm_ipc_connect");
  std::memset(out, 0, sizeof *out);
  const int fa = field_id(h, a), fb = b ? field_id(h, b) : 0;
  if (fa < 0 || fb < 0) return fail(CFDL_ERR_INTERNAL, "p2p_store_args: array is not part of the exported slab");
  const int nc = p.ncolors;
  out->tptr = h->tgt_ptr; out->tnbr = h->tgt_nbr; out->tpos = h->tgt_pos;
  out->nnbr = h->nnbr; out->seq = seq; out->ticket = q.ticket;
  for (int i = 0; i < h->nnbr; ++i) {
    const int r = p.nbr_rank[i];
    const P2PHeader& ph = q.peer_hdr[r];
    int me = -1;
    for (int k = 0; k < ph.nnbr; ++k) if (ph.nbr_rank[k] == p.rank) me = k;
    if (me < 0) return fail(CFDL_ERR_INTERNAL, "p2p: rank %d does not list rank %d as a neighbour", r, p.rank);
    const int cnt = p.send_ptr[(size_t)i * nc + color + 1] - p.send_ptr[(size_t)i * nc + color];
    const int expect_cnt = ph.recv_ptr[me * nc + color + 1] - ph.recv_ptr[me * nc + color];
    if (expect_cnt != cnt) return fail(CFDL_ERR_INTERNAL, "p2p: interface size mismatch with rank %d (%d vs %d)", r, cnt, expect_cnt);
    out->d0[i] = ph.N + ph.recv_ptr[me * nc + color];
    out->dst_a[i] = (double*)(q.peer_base[r] + ph.off_field[fa]);
    out->dst_b[i] = b ? (double*)(q.peer_base[r] + ph.off_field[fb]) : nullptr;
    out->peer_flag[i] = (unsigned long long*)(q.peer_base[r] + ph.off_flags) + p.rank;
  }
  return CFDL_OK;
}

int p2p_reduce_args(Handle* h, int parity, unsigned long long seq, P2PReduce* out) {
  P2P& q = h->p2p;
  const Prep& p = h->prep;
  if (!q.connected) return fail(CFDL_ERR_INTERNAL, "p2p_reduce_args without cfdl_comm_ipc_connect");
  std::memset(out, 0, sizeof *out);
  out->on = 1; out->nranks = p.nranks; out->rank = p.rank; out->parity = parity; out->seq = seq; out->ne_global = (double)h->ne_global;
  for (int r = 0; r < p.nranks; ++r) {
    out->peer_val[r] = (double*)(q.peer_base[r] + q.peer_hdr[r].off_red_val);
    out->peer_seq[r] = (unsigned long long*)(q.peer_base[r] + q.peer_hdr[r].off_red_seq);
  }
  out->my_val = (const double*)(q.slab + q.hdr.off_red_val);
  out->my_seq = (const unsigned long long*)(q.slab + q.hdr.off_red_seq);
  out->err = q.err;
  return CFDL_OK;
}

int p2p_reduce3_args(Handle* h, int parity, unsigned long long seq, P2PReduce* out) {
  int rc = p2p_reduce_args(h, parity, seq, out);
  if (rc) return rc;
  P2P& q = h->p2p;
  for (int r = 0; r < h->prep.nranks; ++r) {
    out->peer_val[r] = (double*)(q.peer_base[r] + q.peer_hdr[r].off_red3_val);
    out->peer_seq[r] = (unsigned long long*)(q.peer_base[r] + q.peer_hdr[r].off_red3_seq);
  }
  out->my_val = (const double*)(q.slab + q.hdr.off_red3_val);
  out->my_seq = (const unsigned long long*)(q.slab + q.hdr.off_red3_seq);
  return CFDL_OK;
}

int p2p_push(Handle* h, int color, const double* a, const double* b, unsigned long long seq, int reduce_parity, const double* local_sm) {
  P2P& q = h->p2p;
  const Prep& p = h->prep;
  if (!q.connected) return fail(CFDL_ERR_INTERNAL, "p2p_push without cfdl_comm_ipc_connect");
  PushArgs A;
  std::memset(&A, 0, sizeof A);
  A.nnbr = h->nnbr; A.narr = b ? 2 : 1; A.do_reduce = reduce_parity >= 0; A.nranks = p.nranks; A.rank = p.rank;
  A.parity = reduce_parity >= 0 ? reduce_parity : 0;
  A.send_cells = h->send_cells;
  A.src[0] = a; A.src[1] = b;
  const int fa = field_id(h, a), fb = b ? field_id(h, b) : 0;
  if (fa < 0 || fb < 0) return fail(CFDL_ERR_INTERNAL, "p2p_push: array is not part of the exported slab");
  const int nc = p.ncolors;
  int total = 0;
  for (int i = 0; i < h->nnbr; ++i) {
    const int r = p.nbr_rank[i];
    const P2PHeader& ph = q.peer_hdr[r];
    int me = -1;
    for (int k = 0; k < ph.nnbr; ++k) if (ph.nbr_rank[k] == p.rank) me = k;
    if (me < 0) return fail(CFDL_ERR_INTERNAL, "p2p_push: rank %d does not list rank %d as a neighbour", r, p.rank);
    A.s0[i] = p.send_ptr[(size_t)i * nc + color];
    A.cnt[i] = p.send_ptr[(size_t)i * nc + color + 1] - A.s0[i];
    const int expect_cnt = ph.recv_ptr[me * nc + color + 1] - ph.recv_ptr[me * nc + color];
    if (expect_cnt != A.cnt[i]) return fail(CFDL_ERR_INTERNAL, "p2p_push: interface size mismatch with rank %d (%d vs %d)", r, A.cnt[i], expect_cnt);
    A.d0[i] = ph.N + ph.recv_ptr[me * nc + color];
    A.dst[i][0] = (double*)(q.peer_base[r] + ph.off_field[fa]);
    A.dst[i][1] = b ? (double*)(q.peer_base[r] + ph.off_field[fb]) : nullptr;
    A.peer_flag[i] = (unsigned long long*)(q.peer_base[r] + ph.off_flags) + p.rank;
    total += A.cnt[i];
  }
  A.seq = seq; A.ticket = q.ticket; A.ctl = h->ctl; A.ne_global = (double)h->ne_global; A.local_sm = local_sm;
  for (int r = 0; r < p.nranks; ++r) {
    A.peer_red_val[r] = (double*)(q.peer_base[r] + q.peer_hdr[r].off_red_val);
    A.peer_red_seq[r] = (unsigned long long*)(q.peer_base[r] + q.peer_hdr[r].off_red_seq);
  }
  A.my_red_val = (const double*)(q.slab + q.hdr.off_red_val);
  A.my_red_seq = (const unsigned long long*)(q.slab + q.hdr.off_red_seq);
  A.err = q.err;
  const int ctas = std::max(1, std::min(32, (total + 255) / 256));
  p2p_push_kernel<<<ctas, 256, 0, S(h)>>>(A);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// ---- staged exchange of any field (all colours at once) -----------------------------------------
// One launch per exchange: every CTA copies its share of the interface values into the
// neighbours' landing zones (laid out like their ghost range), the last CTA to finish raises the
// neighbours' flag words, then all CTAs wait for the neighbours' flags and move the landed values
// into the field's ghost cells.  Two landing zones alternate: a rank can only deliver exchange k+2
// after it has consumed k+1, which its neighbour sent after consuming k.  The grid is far below
// one wave, so the CTAs that spin never keep a publishing CTA from running.
namespace {

struct StageArgs {
  int nnbr, ncomp, N, G;
  int nf;                     // fields exchanged by this launch (1..3), ncomp components each: one rendezvous for all of them
  const int32_t* send_cells;
  double* field[3];           // fields (device numbering, ncomp interleaved)
  int s0[8], cnt[8], d0[8];   // per neighbour: first send cell, count, first ghost slot at the neighbour
  int g0[8], gcnt[8];         // per neighbour: the ghost range that receives (one colour: a sub-range of the neighbour's ghosts)
  int whole;                  // 1: all colours — the landing zone is copied as one block
  double* peer_stage[8];
  unsigned long long* peer_flag[8];
  const double* my_stage;
  const unsigned long long* my_flags;
  int nbr_rank[8];
  unsigned long long seq;
  unsigned int* ticket;
  int* err;
};

__global__ void __launch_bounds__(256) stage_exchange_kernel(const __grid_constant__ StageArgs A) {
  int total = 0;
  for (int i = 0; i < A.nnbr; ++i) total += A.cnt[i];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int i = 0, k = t;
    while (k >= A.cnt[i]) { k -= A.cnt[i]; ++i; }
    const int cell = A.send_cells[A.s0[i] + k];
    const int nct = A.nf * A.ncomp;
    for (int fi = 0; fi < A.nf; ++fi)
      for (int q = 0; q < A.ncomp; ++q) A.peer_stage[i][(size_t)(A.d0[i] + k) * nct + fi * A.ncomp + q] = A.field[fi][(size_t)cell * A.ncomp + q];
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < A.nnbr) *((volatile unsigned long long*)A.peer_flag[threadIdx.x]) = A.seq;
    if (threadIdx.x == 0) *A.ticket = 0;
  }
  if ((int)threadIdx.x < A.nnbr) {
    spin_until_ge(A.my_flags + A.nbr_rank[threadIdx.x], A.seq, A.err);
    __threadfence_system();
  }
  __syncthreads();
  const int nct = A.nf * A.ncomp;
  auto land = [&](size_t t) {  // t indexes the landing zone: ghost g, field fi, component q
    const size_t g = t / nct;
    const int r = (int)(t % nct), fi = r / A.ncomp, q = r % A.ncomp;
    A.field[fi][((size_t)A.N + g) * A.ncomp + q] = __ldcg(&A.my_stage[t]);
  };
  if (A.whole) {
    const size_t n = (size_t)A.G * nct;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) land(t);
    return;
  }
  for (int i = 0; i < A.nnbr; ++i) {  // one colour: only the ghost sub-ranges that received
    const size_t b = (size_t)A.g0[i] * nct, n = (size_t)A.gcnt[i] * nct;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) land(b + t);
  }
}

// small all-to-all through the peers' mailboxes: every rank posts two doubles to every rank,
// waits for all, and combines in rank order (identical result everywhere)
struct MailArgs {
  int nranks, rank, mode, root, parity;
  unsigned long long seq;
  double* dev;  // in: this rank's two values; out: the combination
  double* peer_val[64];
  unsigned long long* peer_seq[64];
  const double* my_val;
  const unsigned long long* my_seq;
  int* err;
};

__global__ void __launch_bounds__(64) mail_kernel(const __grid_constant__ MailArgs A) {
  const int slot = A.parity * 64;
  const double v0 = A.dev[0], v1 = A.mode == 0 ? A.dev[1] : 0.0;
  if ((int)threadIdx.x < A.nranks) {
    const int r = threadIdx.x;
    A.peer_val[r][(slot + A.rank) * 2] = v0;
    A.peer_val[r][(slot + A.rank) * 2 + 1] = v1;
    __threadfence_system();
    *((volatile unsigned long long*)&A.peer_seq[r][slot + A.rank]) = A.seq;
    spin_until_ge(&A.my_seq[slot + r], A.seq, A.err);
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (A.mode == 0) {
      double s = 0.0, m = 0.0;
      for (int r = 0; r < A.nranks; ++r) {
        s += *((volatile const double*)&A.my_val[(slot + r) * 2]);
        m = fmax(m, *((volatile const double*)&A.my_val[(slot + r) * 2 + 1]));
      }
      A.dev[0] = s; A.dev[1] = m;
    } else {
      A.dev[0] = *((volatile const double*)&A.my_val[(slot + A.root) * 2]);
    }
  }
}

}  // namespace

// several fields of equal width in one exchange (peer-to-peer: one launch and one rendezvous for all of them)
int comm_exchange_multi(Handle* h, double* const* fields, int nf, int ncomp) {
  if (h->prep.nranks == 1 || h->nnbr == 0) return CFDL_OK;
  if (nf >= 1 && nf <= 3 && ncomp <= 3 && h->p2p.connected && h->use_p2p) return p2p_exchange(h, fields, nf, ncomp, -1);
  for (int i = 0; i < nf; ++i) {
    int rc = comm_exchange(h, fields[i], ncomp, -1);
    if (rc) return rc;
  }
  return CFDL_OK;
}

static int p2p_exchange(Handle* h, double* const* fields, int nf, int ncomp, int color) {
  P2P& q = h->p2p;
  const Prep& p = h->prep;
  StageArgs A;
  std::memset(&A, 0, sizeof A);
  const int nc = p.ncolors;
  const unsigned long long seq = ++q.xseq;
  const int buf = (int)(seq & 1);
  A.nnbr = h->nnbr; A.ncomp = ncomp; A.N = h->N; A.G = h->G; A.send_cells = h->send_cells; A.nf = nf;
  for (int i = 0; i < nf; ++i) A.field[i] = fields[i];
  int total = 0;
  for (int i = 0; i < h->nnbr; ++i) {
    const int r = p.nbr_rank[i];
    const P2PHeader& ph = q.peer_hdr[r];
    int me = -1;
    for (int k = 0; k < ph.nnbr; ++k) if (ph.nbr_rank[k] == p.rank) me = k;
    if (me < 0) return fail(CFDL_ERR_INTERNAL, "p2p_exchange: rank %d does not list rank %d as a neighbour", r, p.rank);
    const int c0 = color >= 0 ? color : 0, c1 = color >= 0 ? color + 1 : nc;  // colour slices of a neighbour are adjacent
    A.s0[i] = p.send_ptr[(size_t)i * nc + c0];
    A.cnt[i] = p.send_ptr[(size_t)i * nc + c1] - A.s0[i];
    if (ph.recv_ptr[me * nc + c1] - ph.recv_ptr[me * nc + c0] != A.cnt[i]) return fail(CFDL_ERR_INTERNAL, "p2p_exchange: interface size mismatch with rank %d", r);
    A.d0[i] = ph.recv_ptr[me * nc + c0];
    A.g0[i] = p.recv_ptr[(size_t)i * nc + c0];
    A.gcnt[i] = p.recv_ptr[(size_t)i * nc + c1] - A.g0[i];
    A.peer_stage[i] = (double*)(q.peer_base[r] + ph.off_stage[buf]);
    A.peer_flag[i] = (unsigned long long*)(q.peer_base[r] + ph.off_xflag) + p.rank;
    A.nbr_rank[i] = r;
    total += A.cnt[i];
  }
  A.my_stage = (const double*)(q.slab + q.hdr.off_stage[buf]);
  A.my_flags = (const unsigned long long*)(q.slab + q.hdr.off_xflag);
  A.seq = seq; A.ticket = q.xticket; A.whole = color < 0 ? 1 : 0; A.err = q.err;
  // (all CTAs of this launch must be co-resident: they wait for flags that the last CTA to finish its stores raises.  64 CTAs
  // always are on a GPU; the host emulation of tests/emul runs a grid on as many workers as it has cores, hence the cap there)
#ifdef CUEMU
  constexpr int kMaxCtas = 4;
#else
  constexpr int kMaxCtas = 64;
#endif
  const int ctas = std::max(1, std::min(kMaxCtas, (std::max(total, h->G) * ncomp * nf + 255) / 256));
  stage_exchange_kernel<<<ctas, 256, 0, S(h)>>>(A);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

static int p2p_mail(Handle* h, double* dev, int mode, int root) {
  P2P& q = h->p2p;
  const Prep& p = h->prep;
  MailArgs A;
  std::memset(&A, 0, sizeof A);
  A.nranks = p.nranks; A.rank = p.rank; A.mode = mode; A.root = root;
  A.seq = ++q.mseq; A.parity = (int)(A.seq & 1); A.dev = dev;
  for (int r = 0; r < p.nranks; ++r) {
    A.peer_val[r] = (double*)(q.peer_base[r] + q.peer_hdr[r].off_mail_val);
    A.peer_seq[r] = (unsigned long long*)(q.peer_base[r] + q.peer_hdr[r].off_mail_seq);
  }
  A.my_val = (const double*)(q.slab + q.hdr.off_mail_val);
  A.my_seq = (const unsigned long long*)(q.slab + q.hdr.off_mail_seq);
  A.err = q.err;
  mail_kernel<<<1, 64, 0, S(h)>>>(A);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

// vector form of the mailbox: every rank posts npairs (sum, max) pairs to every rank, waits for all and combines
// element by element in rank order — one launch and one rendezvous for a whole block of solver iterations
namespace {
struct MailVArgs {
  int nranks, rank, parity, n;  // n doubles = 2 * npairs
  unsigned long long seq;
  double* dev;
  double* peer_val[64];
  unsigned long long* peer_seq[64];
  const double* my_val;
  const unsigned long long* my_seq;
  int* err;
};
__global__ void __launch_bounds__(256) mailv_kernel(const __grid_constant__ MailVArgs A) {
  const int slot = A.parity * 64;
  for (int r = 0; r < A.nranks; ++r)
    for (int i = threadIdx.x; i < A.n; i += blockDim.x) A.peer_val[r][(size_t)(slot + A.rank) * MAILV_LEN + i] = A.dev[i];
  __syncthreads();  // the CTA's remote stores are ordered before the publishing threads' system-scope fences
  if ((int)threadIdx.x < A.nranks) {
    const int r = threadIdx.x;
    __threadfence_system();
    *((volatile unsigned long long*)&A.peer_seq[r][slot + A.rank]) = A.seq;
    spin_until_ge(&A.my_seq[slot + r], A.seq, A.err);
    __threadfence_system();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < A.n; i += blockDim.x) {
    double v = 0.0;
    for (int r = 0; r < A.nranks; ++r) {
      const double x = *((volatile const double*)&A.my_val[(size_t)(slot + r) * MAILV_LEN + i]);
      v = (i & 1) ? fmax(v, x) : v + x;
    }
    A.dev[i] = v;
  }
}
}  // namespace

int comm_allreduce_pairs(Handle* h, double* dev, int npairs) {
  if (h->prep.nranks == 1) return CFDL_OK;
  P2P& q = h->p2p;
  const Prep& p = h->prep;
  if (!(q.connected && h->use_p2p)) return fail(CFDL_ERR_INTERNAL, "comm_allreduce_pairs needs the peer-to-peer slabs");
  if (2 * npairs > MAILV_LEN) return fail(CFDL_ERR_INTERNAL, "comm_allreduce_pairs: %d pairs exceed the mailbox", npairs);
  MailVArgs A;
  std::memset(&A, 0, sizeof A);
  A.nranks = p.nranks; A.rank = p.rank; A.n = 2 * npairs;
  A.seq = ++q.vseq; A.parity = (int)(A.seq & 1); A.dev = dev;
  for (int r = 0; r < p.nranks; ++r) {
    A.peer_val[r] = (double*)(q.peer_base[r] + q.peer_hdr[r].off_mailv_val);
    A.peer_seq[r] = (unsigned long long*)(q.peer_base[r] + q.peer_hdr[r].off_mailv_seq);
  }
  A.my_val = (const double*)(q.slab + q.hdr.off_mailv_val);
  A.my_seq = (const unsigned long long*)(q.slab + q.hdr.off_mailv_seq);
  A.err = q.err;
  mailv_kernel<<<1, 256, 0, S(h)>>>(A);
  CFDL_CUDA(cudaGetLastError());
  return CFDL_OK;
}

int p2p_check(Handle* h) {
  if (h->prep.nranks == 1 || !h->p2p.err) return CFDL_OK;
  CFDL_CUDA(cudaMemcpyAsync(h->scal_host + 500, h->p2p.err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CFDL_CUDA(cudaStreamSynchronize(h->stream));
  if (*reinterpret_cast<const int*>(h->scal_host + 500))
    return fail(CFDL_ERR_COMM, "rank %d: a wait on another rank's flag word ran into its time limit (did a rank die or never launch?)", h->prep.rank);
  return CFDL_OK;
}

}  // namespace cfdl

using namespace cfdl;

extern "C" {

int cfdl_comm_unique_id(uint8_t id[128]) {
  if (!id) return fail(CFDL_ERR_ARG, "cfdl_comm_unique_id: NULL");
  int rc = load_nccl();
  if (rc) return rc;
  NcclUniqueId u;
  CFDL_NCCL(g_nccl.GetUniqueId(&u));
  std::memcpy(id, u.internal, 128);
  return CFDL_OK;
}

int cfdl_comm_init(cfdl_handle h, const uint8_t id[128], int32_t rank, int32_t nranks) {
  if (!h || !id) return fail(CFDL_ERR_ARG, "cfdl_comm_init: NULL argument");
  if (rank != h->prep.rank || nranks != h->prep.nranks) return fail(CFDL_ERR_ARG, "cfdl_comm_init: handle was created as rank %d of %d", h->prep.rank, h->prep.nranks);
  int rc = load_nccl();
  if (rc) return rc;
  CFDL_CUDA(cudaSetDevice(h->device));
  NcclUniqueId u;
  std::memcpy(u.internal, id, 128);
  NcclComm c = nullptr;
  CFDL_NCCL(g_nccl.CommInitRank(&c, nranks, u, rank));
  h->comm = c;
  return CFDL_OK;
}

int cfdl_comm_ipc_handle(cfdl_handle h, uint8_t handle[64]) {
  if (!h || !handle) return fail(CFDL_ERR_ARG, "cfdl_comm_ipc_handle: NULL argument");
  if (!h->p2p.slab) return fail(CFDL_ERR_UNSUPPORTED, "cfdl_comm_ipc_handle: this handle has no exportable slab (single rank, or too many colours/neighbours)");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CFDL_CUDA(cudaSetDevice(h->device));
  cudaIpcMemHandle_t m;
  CFDL_CUDA(cudaIpcGetMemHandle(&m, h->p2p.slab));
  std::memcpy(handle, &m, 64);
  return CFDL_OK;
}

int cfdl_comm_ipc_connect(cfdl_handle h, const uint8_t* handles) {
  if (!h || !handles) return fail(CFDL_ERR_ARG, "cfdl_comm_ipc_connect: NULL argument");
  P2P& q = h->p2p;
  if (!q.slab) return fail(CFDL_ERR_UNSUPPORTED, "cfdl_comm_ipc_connect: this handle has no exportable slab");
  CFDL_CUDA(cudaSetDevice(h->device));
  const int P = h->prep.nranks;
  q.peer_base.assign(P, nullptr);
  q.peer_hdr.assign(P, P2PHeader());
  for (int r = 0; r < P; ++r) {
    if (r == h->prep.rank) { q.peer_base[r] = q.slab; q.peer_hdr[r] = q.hdr; continue; }
    cudaIpcMemHandle_t m;
    std::memcpy(&m, handles + 64 * (size_t)r, 64);
    void* base = nullptr;
    CFDL_CUDA(cudaIpcOpenMemHandle(&base, m, cudaIpcMemLazyEnablePeerAccess));
    q.opened.push_back(base);
    q.peer_base[r] = (char*)base;
    CFDL_CUDA(cudaMemcpy(&q.peer_hdr[r], base, sizeof(P2PHeader), cudaMemcpyDeviceToHost));
    if (q.peer_hdr[r].magic != kMagic || q.peer_hdr[r].rank != r || q.peer_hdr[r].nranks != P)
      return fail(CFDL_ERR_ARG, "cfdl_comm_ipc_connect: handle %d is not the slab of rank %d", r, r);
  }
  q.connected = true;
  return CFDL_OK;
}

}  // extern "C"
