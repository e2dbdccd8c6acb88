// cuemu runtime: fiber-scheduled CTAs, guarded allocations.  TEST INFRASTRUCTURE ONLY — see
// cuda_runtime.h in this directory.
#include "cuda_runtime.h"

#include <sched.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>
#include <time.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace {

constexpr size_t kStack = 256 * 1024;

// ---- context switch ---------------------------------------------------------------------------------
// swapcontext() makes a signal-mask system call per switch; a CTA of 256 threads meeting at a
// barrier switches ~500 times, so x86-64 gets a register-only switch and other hosts ucontext.
#if defined(__x86_64__)
extern "C" void cuemu_switch(void** save_sp, void* next_sp);
asm(R"(
.text
.globl cuemu_switch
.type cuemu_switch,@function
cuemu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size cuemu_switch,.-cuemu_switch
)");
struct Ctx { void* sp = nullptr; };
inline void ctx_make(Ctx& c, char* stack, size_t size, void (*entry)()) {
  uintptr_t top = ((uintptr_t)stack + size) & ~(uintptr_t)15;
  void** p = (void**)top;
  *--p = nullptr;         // return address of entry (never used: fibers do not return)
  *--p = (void*)entry;    // popped by `ret` of the first switch; rsp is then 8 mod 16 as after a call
  for (int i = 0; i < 6; ++i) *--p = nullptr;
  c.sp = (void*)p;
}
inline void ctx_switch(Ctx& from, Ctx& to) { cuemu_switch(&from.sp, to.sp); }
#else
struct Ctx { ucontext_t uc; };
inline void ctx_make(Ctx& c, char* stack, size_t size, void (*entry)()) {
  getcontext(&c.uc);
  c.uc.uc_stack.ss_sp = stack; c.uc.uc_stack.ss_size = size; c.uc.uc_link = nullptr;
  makecontext(&c.uc, entry, 0);
}
inline void ctx_switch(Ctx& from, Ctx& to) { swapcontext(&from.uc, &to.uc); }
#endif

struct Warp {
  int live = 0, arrived = 0;
  unsigned gen = 0;
  unsigned long long slot[32];
};

struct Cta {
  int n = 0, live = 0, cur = -1;
  int bar_arrived = 0;
  unsigned bar_gen = 0;
  std::vector<Ctx> ctx;
  std::vector<char> done;
  std::vector<Warp> warps;
  Ctx sched;
  cuemu::Body body;
  dim3 bdim;
};

// per OS thread: the CTA being run and a pool of fiber stacks that is reused by every CTA this
// thread executes
thread_local Cta* t_cta = nullptr;
thread_local std::vector<char*> t_stacks;

char* stack_for(int i) {
  while ((int)t_stacks.size() <= i) {
    void* p = mmap(nullptr, kStack + 4096, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { std::perror("cuemu: mmap stack"); std::abort(); }
    mprotect(p, 4096, PROT_NONE);  // stack overflow faults instead of corrupting the neighbour
    t_stacks.push_back((char*)p + 4096);
  }
  return t_stacks[i];
}

void set_thread_idx(const Cta& c, int t) {
  threadIdx.x = t % c.bdim.x;
  threadIdx.y = (t / c.bdim.x) % c.bdim.y;
  threadIdx.z = t / (c.bdim.x * c.bdim.y);
}

void yield_to_sched() {
  Cta* c = t_cta;
  ctx_switch(c->ctx[c->cur], c->sched);
}

void release_if_complete(Cta* c) {
  if (c->live > 0 && c->bar_arrived == c->live) { c->bar_arrived = 0; c->bar_gen++; }
}
void release_warp_if_complete(Warp& w) {
  if (w.live > 0 && w.arrived == w.live) { w.arrived = 0; w.gen++; }
}

void fiber_main() {
  Cta* c = t_cta;
  c->body.call(c->body.ctx);
  const int t = c->cur;
  c->done[t] = 1;
  c->live--;
  Warp& w = c->warps[t >> 5];
  w.live--;
  // threads that exit no longer take part in barriers (CUDA semantics since Volta)
  release_if_complete(c);
  release_warp_if_complete(w);
  ctx_switch(c->ctx[t], c->sched);
  std::fprintf(stderr, "cuemu: finished fiber resumed\n");
  std::abort();
}

void run_cta(cuemu::Body body, dim3 bdim, dim3 gdim, uint3 bidx) {
  Cta c;
  c.n = (int)(bdim.x * bdim.y * bdim.z);
  c.live = c.n;
  c.body = body;
  c.bdim = bdim;
  c.ctx.resize(c.n);
  c.done.assign(c.n, 0);
  c.warps.resize((c.n + 31) / 32);
  for (int t = 0; t < c.n; ++t) c.warps[t >> 5].live++;
  blockDim = bdim; gridDim = gdim; blockIdx = bidx;
  Cta* saved = t_cta;
  t_cta = &c;
  for (int t = 0; t < c.n; ++t) ctx_make(c.ctx[t], stack_for(t), kStack, fiber_main);
  int ndone = 0;
  unsigned long long rounds = 0;
  while (ndone < c.n) {
    for (int t = 0; t < c.n; ++t) {
      if (c.done[t]) continue;
      c.cur = t;
      set_thread_idx(c, t);
      ctx_switch(c.sched, c.ctx[t]);
      if (c.done[t]) ndone++;
    }
    if (++rounds > 2000000000ull) { std::fprintf(stderr, "cuemu: CTA does not terminate (barrier deadlock?)\n"); std::abort(); }
  }
  t_cta = saved;
}

// ---- worker pool (one per launching OS thread) ------------------------------------------------------
// Ordinary launches spread their CTAs over the host cores; a cooperative launch gets one worker per
// CTA so that all CTAs are alive at once.  The launching thread is worker 0.
struct Pool {
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::vector<std::thread> th;
  unsigned long long job = 0;
  int nworkers = 0, remaining = 0;
  bool quit = false;
  std::function<void(int)> fn;
  void worker(int w) {
    unsigned long long seen = 0;
    for (;;) {
      std::function<void(int)> f;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return quit || (job != seen && w < nworkers); });
        if (quit) return;
        seen = job;
        f = fn;
      }
      f(w);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--remaining == 0) cv_done.notify_all();
      }
    }
  }
  void run(int n, const std::function<void(int)>& f) {
    if (n <= 1) { f(0); return; }
    {
      std::lock_guard<std::mutex> lk(mu);
      while ((int)th.size() < n - 1) { const int w = (int)th.size() + 1; th.emplace_back([this, w] { worker(w); }); }
      fn = f; nworkers = n; remaining = n - 1; job++;
    }
    cv_work.notify_all();
    f(0);
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return remaining == 0; });
    nworkers = 0;
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(mu); quit = true; }
    cv_work.notify_all();
    for (auto& t : th) t.join();
  }
};
thread_local std::unique_ptr<Pool> t_pool;
thread_local bool t_in_worker = false;

}  // namespace

namespace cuemu {

void sync_threads() {
  Cta* c = t_cta;
  if (!c) return;
  const unsigned g = c->bar_gen;
  c->bar_arrived++;
  release_if_complete(c);
  while (c->bar_gen == g) yield_to_sched();
}

void warp_exchange(const void* mine, void* out, size_t bytes) {
  Cta* c = t_cta;
  const int t = c->cur, lane = t & 31;
  Warp& w = c->warps[t >> 5];
  std::memcpy(&w.slot[lane], mine, bytes);
  // phase 1: everyone has published
  unsigned g = w.gen;
  w.arrived++;
  release_warp_if_complete(w);
  while (w.gen == g) yield_to_sched();
  for (int l = 0; l < 32; ++l) std::memcpy((char*)out + l * bytes, &w.slot[l], bytes);
  // phase 2: everyone has read (the slots may be overwritten by the next exchange)
  g = w.gen;
  w.arrived++;
  release_warp_if_complete(w);
  while (w.gen == g) yield_to_sched();
}

// a spinning thread lets its siblings in the CTA run (independent thread scheduling), then the core
void spin_pause() {
  if (t_cta) yield_to_sched();
  sched_yield();
}

void run_grid(dim3 grid, dim3 block, Body body, bool cooperative) {
  const uint3 s_t = threadIdx, s_b = blockIdx;
  const dim3 s_bd = blockDim, s_gd = gridDim;
  const long long nblocks = (long long)grid.x * grid.y * grid.z;
  auto block_at = [&](long long i) { return uint3{(unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((long long)grid.x * grid.y))}; };
  static const int hw = [] { const char* e = std::getenv("CUEMU_THREADS"); int n = e ? std::atoi(e) : (int)std::thread::hardware_concurrency(); return n < 1 ? 1 : n; }();
  if (nblocks == 1 || (!cooperative && hw == 1)) {
    for (long long i = 0; i < nblocks; ++i) run_cta(body, block, grid, block_at(i));
  } else {
    if (!t_pool) t_pool.reset(new Pool);
    if (cooperative) {
      t_pool->run((int)nblocks, [&](int w) { run_cta(body, block, grid, block_at(w)); });
    } else {
      std::atomic<long long> next{0};
      t_pool->run((int)std::min<long long>(nblocks, hw), [&](int) {
        for (long long i; (i = next.fetch_add(1)) < nblocks;) run_cta(body, block, grid, block_at(i));
      });
    }
  }
  threadIdx = s_t; blockIdx = s_b; blockDim = s_bd; gridDim = s_gd;
}

}  // namespace cuemu

// ---- memory ---------------------------------------------------------------------------------------
namespace {
std::mutex g_mu;
struct Region { void* base; size_t len; };
std::map<void*, Region> g_regions;
size_t page() { static size_t p = (size_t)sysconf(_SC_PAGESIZE); return p; }
}  // namespace

cudaError_t cuemu_malloc(void** out, size_t n) {
  const size_t pg = page();
  const size_t need = (n + 15) & ~(size_t)15;
  const size_t data_pages = (need + pg - 1) / pg;
  const size_t len = (data_pages + 2) * pg;  // guard page in front and behind
  char* base = (char*)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (base == (char*)MAP_FAILED) return cudaErrorMemoryAllocation;
  mprotect(base, pg, PROT_NONE);
  mprotect(base + (data_pages + 1) * pg, pg, PROT_NONE);
  char* p = base + (data_pages + 1) * pg - need;  // the allocation ends at the guard page
  std::memset(base + pg, 0xFF, data_pages * pg);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_regions[p] = Region{base, len};
  }
  *out = p;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_regions.find(p);
  if (it == g_regions.end()) return cudaErrorInvalidValue;
  munmap(it->second.base, it->second.len);
  g_regions.erase(it);
  return cudaSuccess;
}
cudaError_t cuemu_malloc_host(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) std::memset(d, v, n); return cudaSuccess; }

// ---- device / streams / events --------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) { const char* e = std::getenv("CUEMU_DEVICES"); *n = e ? std::atoi(e) : 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  std::memset(p, 0, sizeof *p);
  std::snprintf(p->name, sizeof p->name, "cuemu (host emulation, tests only)");
  p->major = 10; p->minor = 0; p->cooperativeLaunch = 1;
  const char* e = std::getenv("CUEMU_SMS");
  p->multiProcessorCount = e ? std::atoi(e) : 3;
  p->totalGlobalMem = (size_t)8 << 30;
  p->l2CacheSize = 126 << 20;
  return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuemu error"; }
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
// CUEMU_RANDOM_TIMES=<seed>: event intervals become pseudo-random, so that code which chooses between
// (bit-identical) kernel variants by timing them takes a different choice in every run of the tests
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
  static const char* e = std::getenv("CUEMU_RANDOM_TIMES");
  static std::atomic<unsigned long long> state{e ? std::strtoull(e, nullptr, 10) * 2654435761ull + 12345ull : 0ull};
  if (!e) { *ms = 1e-3f; return cudaSuccess; }
  unsigned long long x = state.fetch_add(0x9E3779B97F4A7C15ull) + 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; x ^= x >> 31;
  *ms = 0.5f + (float)(x >> 40) / (float)(1 << 24);
  return cudaSuccess;
}
// "ranks" of an emulated multi-GPU run are threads of one process: the handle is the pointer
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof *h); std::memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// "SM clock": MICROseconds of the host's steady clock — the kernels only use it for time limits on their spin
// waits (a few 1e9 ticks); oversubscribed host threads can wait long, so a tick is deliberately slow here
long long clock64() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (long long)ts.tv_sec * 1000000ll + ts.tv_nsec / 1000;
}

extern "C" int cfdl_emulated(void) { return 1; }
