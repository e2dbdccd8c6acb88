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

// order-independent checksum of an array's values: the wrapping sum of the 64-bit patterns, +0 and -0 alike
__global__ void __launch_bounds__(256) checksum_kernel(size_t n, const double* __restrict__ a, unsigned long long* out) {
  unsigned long long s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = a[i];
    s += v == 0.0 ? 0ull : (unsigned long long)__double_as_longlong(v);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

static bool outputs_checksum(Handle* h, const std::vector<TuneOutput>& outs, unsigned long long* cs) {
  unsigned long long* dev = reinterpret_cast<unsigned long long*>(h->scal + 444);
  if (cudaMemsetAsync(dev, 0, sizeof(unsigned long long), h->stream) != cudaSuccess) return false;
  for (const TuneOutput& o : outs)
    if (o.p && o.n) checksum_kernel<<<grid_for(h, (int64_t)o.n, 256), 256, 0, S(h)>>>(o.n, o.p, dev);
  return cudaMemcpyAsync(cs, dev, sizeof *cs, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess && cudaStreamSynchronize(h->stream) == cudaSuccess;
}

int autotune_pick(Handle* h, Handle::Tuned& T, const int* cands, int n, const std::function<int(int)>& run, int reps,
                  const std::vector<TuneOutput>& outputs) {
  T.done = 1;
  unsigned long long ref_cs = 0;
  bool have_ref = false;
  T.ncand = 0;
  T.choice = cands[0];
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    if (e0) cudaEventDestroy(e0);
    cudaGetLastError();
    return CFDL_OK;  // no timing possible: keep the first candidate
  }
  float best = 0.f;
  bool have = false;
  for (int i = 0; i < n && T.ncand < 16; ++i) {
    const int c = cands[i];
    float ms = -1.f;
    bool ok = run(c) == CFDL_OK && cudaStreamSynchronize(h->stream) == cudaSuccess;  // warm-up (and a check that it launches)
    bool wrong = false;
    if (ok && !outputs.empty()) {  // the candidate's outputs against the first candidate's
      unsigned long long cs = 0;
      if (outputs_checksum(h, outputs, &cs)) {
        if (!have_ref) { ref_cs = cs; have_ref = true; }
        else if (cs != ref_cs) wrong = true;
      }
    }
    if (wrong) {
      T.cand[T.ncand] = c; T.ms[T.ncand] = -2.f; T.ncand++;  // ran, but did not reproduce the reference form: never chosen
      continue;
    }
    if (ok) ok = cudaEventRecord(e0, h->stream) == cudaSuccess;
    for (int r = 0; ok && r < reps; ++r) ok = run(c) == CFDL_OK;
    if (ok) ok = cudaEventRecord(e1, h->stream) == cudaSuccess && cudaEventSynchronize(e1) == cudaSuccess &&
                 cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess;
    if (!ok) { cudaGetLastError(); ms = -1.f; }
    T.cand[T.ncand] = c;
    T.ms[T.ncand] = ok ? ms / reps : -1.f;
    T.ncand++;
    if (ok && (!have || ms < best)) { best = ms; have = true; T.choice = c; }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
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
