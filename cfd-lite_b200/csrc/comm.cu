// Multi-GPU plumbing: one process per GPU, one partition per process.  Ghost-cell values move
// by NCCL send/recv over NVLink (grouped per exchange), residual scalars by NCCL all-reduce,
// the reference pressure pc(1) by NCCL broadcast.  NCCL is dlopen'ed (libnccl.so.2 — the copy
// torch already loaded when the host is Python), so the library loads on boxes without NCCL
// and single-GPU use never touches it.  This is the GPU analogue of update_halos
// (src/modules/mod_subdomains.f90:191-212) and of the residual accumulation loop of
// multi_subdomain_solver (src/modules/mod_solver.f90:144-150,172-178).
#include <dlfcn.h>
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

int comm_exchange(Handle* h, double* field, int ncomp, int color) {
  if (h->prep.nranks == 1 || h->nnbr == 0) return CFDL_OK;
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
  if (!h->comm) return fail(CFDL_ERR_NCCL, "cfdl_comm_init has not been called");
  CFDL_NCCL(g_nccl.GroupStart());
  CFDL_NCCL(g_nccl.AllReduce(dev2, dev2, 1, kNcclDouble, kNcclSum, h->comm, h->stream));
  CFDL_NCCL(g_nccl.AllReduce(dev2 + 1, dev2 + 1, 1, kNcclDouble, kNcclMax, h->comm, h->stream));
  CFDL_NCCL(g_nccl.GroupEnd());
  return CFDL_OK;
}

int comm_bcast(Handle* h, double* dev, int count, int root) {
  if (h->prep.nranks == 1) return CFDL_OK;
  if (!h->comm) return fail(CFDL_ERR_NCCL, "cfdl_comm_init has not been called");
  CFDL_NCCL(g_nccl.Broadcast(dev, dev, (size_t)count, kNcclDouble, root, h->comm, h->stream));
  return CFDL_OK;
}

void comm_destroy(Handle* h) {
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr;
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

}  // extern "C"
