// cuemu — a small host emulation of the CUDA runtime + device language subset that libcfdl uses.
//
// TEST INFRASTRUCTURE ONLY (like oracle/): it exists so that the kernels and the orchestration in
// cfd-lite_b200/csrc can be executed, bounds-checked and compared with the oracle in a container
// that has no GPU.  Nothing in the product includes, links or loads it: tests/emul/Makefile
// compiles the *.cu sources (kernel launches rewritten by rewrite.py) against this header into
// tests/emul/_build/libcfdl_emul.so, which only tests/test_emul_*.py load, explicitly.
//
// Execution model: a launch is synchronous.  CTAs run one after the other on the calling OS thread
// (cooperative launches: one OS thread per CTA, all alive at once, so grid barriers that spin on
// global memory work); the threads of a CTA are ucontext fibers scheduled round-robin, so
// __syncthreads() and warp shuffles have their CUDA meaning.  FP64 arithmetic is IEEE on both
// sides (library built -fmad=false, this build -ffp-contract=off) and the reductions use the same
// shuffle trees, so results are expected to be bit-identical to the GPU's.
// Device memory is host memory with a PROT_NONE guard page after every allocation (an
// out-of-bounds access faults) and filled with 0xFF bytes (NaN / -1) so that reads of
// uninitialised memory show up in the results.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <utility>

#define CUEMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local
#define __constant__ static const
#define __CUDACC_VER_MAJOR__ 12

using std::max;
using std::min;

struct uint3 { unsigned x, y, z; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct cuemu_stream_s* cudaStream_t;
typedef struct cuemu_event_s* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaDeviceProp {
  char name[256];
  int major, minor, multiProcessorCount, cooperativeLaunch, l2CacheSize;
  size_t totalGlobalMem;
};
struct cudaIpcMemHandle_t { char reserved[64]; };

cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cuemu_malloc(void** p, size_t n);
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cuemu_malloc((void**)p, n); }
cudaError_t cudaFree(void* p);
cudaError_t cuemu_malloc_host(void** p, size_t n);
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cuemu_malloc_host((void**)p, n); }
cudaError_t cudaFreeHost(void* p);
enum { cudaHostRegisterDefault = 0 };
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void* dst, int v, size_t n);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t* s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaDeviceSynchronize();
inline cudaError_t cudaCtxResetPersistingL2Cache() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e);
enum { cudaEventDisableTiming = 2 };
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
enum { cudaFuncAttributePreferredSharedMemoryCarveout = 9, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class Kern> inline cudaError_t cudaFuncSetAttribute(Kern, int, int) { return cudaSuccess; }
template <class Kern> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, Kern, int, size_t) { *n = 2; return cudaSuccess; }

// ---- launches -----------------------------------------------------------------------------------
namespace cuemu {
struct Cfg { dim3 grid, block; };
inline Cfg cfg(dim3 g, dim3 b, size_t = 0, cudaStream_t = nullptr) { return Cfg{g, b}; }
struct Body { void (*call)(void*); void* ctx; };
void run_grid(dim3 grid, dim3 block, Body body, bool cooperative);
template <class F> inline void launch(Cfg c, F&& f, bool cooperative = false) {
  Body b{[](void* p) { (*static_cast<typename std::remove_reference<F>::type*>(p))(); }, (void*)&f};
  run_grid(c.grid, c.block, b, cooperative);
}
void sync_threads();
void warp_exchange(const void* mine, void* slots_out, size_t bytes);  // all lanes publish `mine`, every lane gets the 32 values
void spin_pause();
template <class... Exp, size_t... I>
inline void call_packed(void (*k)(Exp...), void** a, std::index_sequence<I...>) { k(*static_cast<typename std::remove_reference<Exp>::type*>(a[I])...); }
}  // namespace cuemu

template <class... Exp>
inline cudaError_t cudaLaunchCooperativeKernel(void (*kern)(Exp...), dim3 grid, dim3 block, void** args, size_t, cudaStream_t) {
  cuemu::launch(cuemu::Cfg{grid, block}, [&]() { cuemu::call_packed(kern, args, std::index_sequence_for<Exp...>{}); }, true);
  return cudaSuccess;
}
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; unsigned numAttrs; };
template <class... Exp, class... Act>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* c, void (*kern)(Exp...), Act&&... args) {
  cuemu::launch(cuemu::Cfg{c->gridDim, c->blockDim}, [&]() { kern(args...); });
  return cudaSuccess;
}

// ---- device intrinsics --------------------------------------------------------------------------
inline void __syncthreads() { cuemu::sync_threads(); }
inline int __syncthreads_or(int pred) {
  __shared__ int acc;
  cuemu::sync_threads();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) acc = 0;
  cuemu::sync_threads();
  if (pred) acc = 1;
  cuemu::sync_threads();
  return acc;
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> inline T __ldcg(const T* p) { return *(const volatile T*)p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline void __stcg(T* p, T v) { *(volatile T*)p = v; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int delta) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  T all[32];
  cuemu::warp_exchange(&v, all, sizeof(T));
  const int lane = (int)((threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31u);
  return lane + delta < 32 ? all[lane + delta] : v;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) {
  T all[32];
  cuemu::warp_exchange(&v, all, sizeof(T));
  const int lane = (int)((threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31u);
  return all[(lane ^ m) & 31];
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) {
  T all[32];
  cuemu::warp_exchange(&v, all, sizeof(T));
  return all[src & 31];
}
inline void __syncwarp(unsigned = 0xffffffffu) { char c = 0, all[32]; cuemu::warp_exchange(&c, all, 1); }
long long clock64();
inline long long __double_as_longlong(double v) { long long r; std::memcpy(&r, &v, sizeof r); return r; }
inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, sizeof r); return r; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __drcp_rn(double a) { return 1.0 / a; }
// spin loops with an empty body appear in the sources as `while (cond) {}`; rewrite.py turns the
// body into CUEMU_SPIN so that a waiting CTA yields its core
#define CUEMU_SPIN ::cuemu::spin_pause()
