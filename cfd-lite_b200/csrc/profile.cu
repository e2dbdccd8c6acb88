// Instrumentation used by bench.py: CUDA-event timers on the library's own stream (events on
// torch's current stream would not see these launches), optional per-launch timing of the
// dominant kernels for the roofline figure, and pinned host buffers for the end-to-end path.
#include "state.h"

namespace cfdl {

// level 1 (profile=1): every launch of the profiled kinds is bracketed by its own event pair.
// level 2 (profile=2): only whole batches of back-to-back solver passes are bracketed, so the
// passes keep overlapping as they do in production; the batch owner adds the launch count.
int prof_begin(Handle* h, int kind, int level) {
  if (h->profile != level) return CFDL_OK;
  if (h->prof_used + 2 > h->prof_ev.size()) {
    const size_t old = h->prof_ev.size(), grow = 4096;
    h->prof_ev.resize(old + grow, nullptr);
    h->prof_kind.resize((old + grow) / 2, 0);
    h->prof_cnt.resize((old + grow) / 2, 1);
    for (size_t i = old; i < old + grow; ++i) CFDL_CUDA(cudaEventCreate(&h->prof_ev[i]));
  }
  h->prof_kind[h->prof_used / 2] = kind;
  CFDL_CUDA(cudaEventRecord(h->prof_ev[h->prof_used], h->stream));
  return CFDL_OK;
}
int prof_end(Handle* h, int count, int level) {
  if (h->profile != level) return CFDL_OK;
  h->prof_cnt[h->prof_used / 2] = count;
  CFDL_CUDA(cudaEventRecord(h->prof_ev[h->prof_used + 1], h->stream));
  h->prof_used += 2;
  return CFDL_OK;
}
int prof_collect(Handle* h) {
  if (!h->profile || h->prof_used == 0) return CFDL_OK;
  CFDL_CUDA(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < h->prof_used; i += 2) {
    float ms = 0.f;
    CFDL_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]));
    const int k = h->prof_kind[i / 2];
    h->prof_ms[k] += ms;
    h->prof_n[k] += h->prof_cnt[i / 2];
  }
  h->prof_used = 0;
  return CFDL_OK;
}

}  // namespace cfdl

using namespace cfdl;

extern "C" {

int cfdl_timer_record(cfdl_handle h, int32_t slot) {
  if (!h || slot < 0 || slot >= 4) return fail(CFDL_ERR_ARG, "cfdl_timer_record: bad handle/slot");
  CFDL_CUDA(cudaSetDevice(h->device));
  if (!h->timer_ev[slot]) CFDL_CUDA(cudaEventCreate(&h->timer_ev[slot]));
  CFDL_CUDA(cudaEventRecord(h->timer_ev[slot], h->stream));
  return CFDL_OK;
}
int cfdl_timer_elapsed_ms(cfdl_handle h, int32_t slot_begin, int32_t slot_end, double* ms) {
  if (!h || !ms || slot_begin < 0 || slot_begin >= 4 || slot_end < 0 || slot_end >= 4 || !h->timer_ev[slot_begin] || !h->timer_ev[slot_end])
    return fail(CFDL_ERR_ARG, "cfdl_timer_elapsed_ms: bad handle/slot");
  CFDL_CUDA(cudaSetDevice(h->device));
  CFDL_CUDA(cudaEventSynchronize(h->timer_ev[slot_end]));
  float f = 0.f;
  CFDL_CUDA(cudaEventElapsedTime(&f, h->timer_ev[slot_begin], h->timer_ev[slot_end]));
  *ms = f;
  return CFDL_OK;
}
int cfdl_host_alloc(void** ptr, uint64_t bytes) {
  if (!ptr) return fail(CFDL_ERR_ARG, "cfdl_host_alloc: NULL");
  CFDL_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
  return CFDL_OK;
}
// page-lock memory the caller owns (e.g. the allocatable arrays of uvwp_t), so that cfdl_step_host's
// transfers run beside the computation; unregister before the memory is freed
int cfdl_host_register(void* ptr, uint64_t bytes) {
  if (!ptr || !bytes) return fail(CFDL_ERR_ARG, "cfdl_host_register: NULL / empty range");
  CFDL_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return CFDL_OK;
}
int cfdl_host_unregister(void* ptr) {
  if (ptr) CFDL_CUDA(cudaHostUnregister(ptr));
  return CFDL_OK;
}
int cfdl_host_free(void* ptr) {
  if (ptr) CFDL_CUDA(cudaFreeHost(ptr));
  return CFDL_OK;
}

}  // extern "C"
