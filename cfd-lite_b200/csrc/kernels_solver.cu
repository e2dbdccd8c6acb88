// Linear solvers for a_p phi_e = b + sum a_nb phi_nb (sm_100a), replacing solve_gs /
// smoother_gs / calc_residual / multi_subdomain_solver (src/modules/mod_solver.f90:124-344).
//
//  PARITY  exact reference order.  The sequential sweep 1..ne (or the block-local order of
//          multi_subdomain_solver) is executed as a level schedule: cells whose same-block
//          lower-numbered neighbours are all done form one level; levels run one after the
//          other inside ONE persistent kernel separated by a grid barrier, backward sweeps walk
//          the levels in reverse.  Every cell sees exactly the operand values it sees in the
//          sequential loop, so phi is bit-identical to the reference sweep.
//  MCSGS   the same update formula in multicolour order (device numbering is colour-major, so
//          each colour is one fully coalesced launch); identical to solve_gs applied to the
//          colour-permuted system.
//  PCG     Jacobi-preconditioned conjugate gradients for the symmetric pc system, with the
//          reference's stopping rule (10x RMS-residual drop or nit iterations); kernels_pcg.inc.
#include <cstdlib>
#include <cstring>
#include "state.h"
#include "device_math.cuh"

namespace cfdl {

#define TPB 256

// ---------------------------------------------------------------------------------------------
// block reduction helpers (warp shuffle, then shared memory across warps); deterministic
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}
template <int NT>
__device__ __forceinline__ void block_sum_max(double& s, double& m) {
  __shared__ double sh_s[NT / 32], sh_m[NT / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  s = warp_sum(s); m = warp_max(m);
  if (lane == 0) { sh_s[wid] = s; sh_m[wid] = m; }
  __syncthreads();
  if (wid == 0) {
    s = (lane < NT / 32) ? sh_s[lane] : 0.0;
    m = (lane < NT / 32) ? sh_m[lane] : 0.0;
    s = warp_sum(s); m = warp_max(m);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// one Gauss-Seidel / SOR update over a contiguous range of mutually independent cells
// (mod_solver.f90:290-297): phi = (b + sum anb*phi_nb + (sor-1)*ap*phi) / ap / sor
template <int K>
__global__ void __launch_bounds__(TPB) sgs_range_kernel(int c0, int c1, int Np, const int32_t* __restrict__ nbi,
                                                        const double* __restrict__ ap, const double* __restrict__ anb,
                                                        const double* __restrict__ b, double* phi, double sor,
                                                        const SolveCtl* __restrict__ ctl) {
  if (ctl && ctl->done) return;
  const double sm1 = sor - 1.0;
  for (int c = c0 + blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += gridDim.x * blockDim.x) {
    double sumnb = b[c];
#pragma unroll
    for (int k = 0; k < K; ++k) sumnb = sumnb + anb[(size_t)k * Np + c] * phi[nbi[(size_t)k * Np + c]];
    const double a = ap[c];
    phi[c] = (sumnb + sm1 * a * phi[c]) / a / sor;
  }
}

// residual r = b + sum anb*phi_nb - ap*phi (mod_solver.f90:309-321 / 230-253): per-CTA partial
// sums of r^2 and max|r| (or max(0, r) in signed mode), combined in fixed order by the last
// CTA to finish, which also advances the solve control block.
enum { RES_INIT = 0, RES_ITER = 1, RES_PLAIN = 2, RES_LOCAL = 3, RES_LOCAL_GUARDED = 4 };
template <int K>
__global__ void __launch_bounds__(TPB) residual_kernel(int N, int Np, const int32_t* __restrict__ nbi,
                                                       const double* __restrict__ ap, const double* __restrict__ anb,
                                                       const double* __restrict__ b, const double* __restrict__ phi,
                                                       double* partial, SolveCtl* ctl, int mode, int signed_max, double* out2) {
  if ((mode == RES_ITER || mode == RES_LOCAL_GUARDED) && ctl->done) return;
  double s = 0.0, m = 0.0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    double sumnb = b[c];
#pragma unroll
    for (int k = 0; k < K; ++k) sumnb = sumnb + anb[(size_t)k * Np + c] * phi[nbi[(size_t)k * Np + c]];
    double r = sumnb - ap[c] * phi[c];
    if (!signed_max) r = fabs(r);
    m = fmax(m, r);
    s = s + r * r;
  }
  block_sum_max<TPB>(s, m);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = s;
    partial[2 * blockIdx.x + 1] = m;
    __threadfence();
    unsigned t = atomicAdd(&ctl->ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  s = 0.0; m = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {  // fixed order: deterministic
    s += __ldcg(&partial[2 * i]);
    m = fmax(m, __ldcg(&partial[2 * i + 1]));
  }
  block_sum_max<TPB>(s, m);
  if (threadIdx.x == 0) {
    ctl->ticket = 0;
    const double res = sqrt(s / N);
    if (mode == RES_INIT) {
      ctl->it = 0;
      ctl->res_i = res; ctl->res_f = res; ctl->res_max = 0.0;
      ctl->res_target = res / 10.0;
      ctl->done = !(0 < ctl->nit && res > ctl->res_target);
    } else if (mode == RES_ITER) {
      ctl->it += 1;
      ctl->res_f = res; ctl->res_max = m;
      ctl->done = !(ctl->it < ctl->nit && res > ctl->res_target);
    } else if (mode == RES_PLAIN) {
      out2[0] = res; out2[1] = m;
    } else {  // RES_LOCAL: this rank's sum of r^2 and max, combined across ranks by NCCL
      out2[0] = s; out2[1] = m;
    }
  }
}

// several GPUs: after the all-reduce of (sum r^2, max) advance the control block on every rank
__global__ void finalize_residual_kernel(SolveCtl* ctl, const double* sm, double ne_global, int mode) {
  if (mode == RES_ITER && ctl->done) return;
  const double res = sqrt(sm[0] / ne_global);
  if (mode == RES_INIT) {
    ctl->it = 0;
    ctl->res_i = res; ctl->res_f = res; ctl->res_max = 0.0;
    ctl->res_target = res / 10.0;
    ctl->done = !(0 < ctl->nit && res > ctl->res_target);
  } else {
    ctl->it += 1;
    ctl->res_f = res; ctl->res_max = sm[1];
    ctl->done = !(ctl->it < ctl->nit && res > ctl->res_target);
  }
}

// ---------------------------------------------------------------------------------------------
// PARITY: persistent level-scheduled sweep kernel.  All CTAs are co-resident (cooperative
// launch); levels are separated by a grid-wide barrier built on one global counter.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int nctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nctas;
    __threadfence();
    atomicAdd(counter, 1u);
    while (*((volatile unsigned int*)counter) < target) {}
    __threadfence();
  }
  __syncthreads();
}

template <int K>
__device__ __forceinline__ void level_update(int s, int Np, const int32_t* __restrict__ nbs, const double* __restrict__ ap,
                                             const double* __restrict__ anb, const double* __restrict__ b, double* phi,
                                             double sor, double sm1) {
  double sumnb = b[s];
#pragma unroll
  for (int k = 0; k < K; ++k) sumnb = sumnb + anb[(size_t)k * Np + s] * __ldcg(&phi[nbs[(size_t)k * Np + s]]);
  const double a = ap[s];
  __stcg(&phi[s], (sumnb + sm1 * a * __ldcg(&phi[s])) / a / sor);
}

// nsweeps symmetric iterations (forward 1..ne then backward ne..1 each), then the lagged
// copies other blocks read are refreshed (update_halos, mod_subdomains.f90:191-212)
template <int K>
__global__ void __launch_bounds__(TPB) level_sgs_kernel(int nlevels, const int32_t* __restrict__ lvl_ptr, int Np, int H,
                                                        const int32_t* __restrict__ nbs, const double* __restrict__ ap,
                                                        const double* __restrict__ anb, const double* __restrict__ b,
                                                        double* phi, double sor, int niter, int nlag,
                                                        const int32_t* __restrict__ lag_src, unsigned int* counter) {
  unsigned int target = 0;
  const double sm1 = sor - 1.0;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int it = 0; it < niter; ++it) {
    for (int l = 0; l < nlevels; ++l) {
      const int s1 = lvl_ptr[l + 1];
      for (int s = lvl_ptr[l] + tid; s < s1; s += nth) level_update<K>(s, Np, nbs, ap, anb, b, phi, sor, sm1);
      grid_barrier(counter, target, gridDim.x);
    }
    for (int l = nlevels - 1; l >= 0; --l) {
      const int s1 = lvl_ptr[l + 1];
      for (int s = lvl_ptr[l] + tid; s < s1; s += nth) level_update<K>(s, Np, nbs, ap, anb, b, phi, sor, sm1);
      grid_barrier(counter, target, gridDim.x);
    }
  }
  for (int i = tid; i < nlag; i += nth) { const int s = lag_src[i]; __stcg(&phi[H + s], __ldcg(&phi[s])); }
}

// sweep-space staging: gather device-numbered system into level order and back
template <int K>
__global__ void __launch_bounds__(TPB) to_sweep_kernel(int N, int Np, const int32_t* __restrict__ s2c, const double* __restrict__ ap,
                                                       const double* __restrict__ anb, const double* __restrict__ b,
                                                       const double* __restrict__ phi, double* ap_s, double* anb_s, double* b_s,
                                                       double* phi_s, int H) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < N; s += gridDim.x * blockDim.x) {
    const int c = s2c[s];
    ap_s[s] = ap[c]; b_s[s] = b[c];
    const double p = phi[c];
    phi_s[s] = p;
    phi_s[H + s] = p;  // lag copy starts equal (assemble_coef copies phi into every block's halos)
#pragma unroll
    for (int k = 0; k < K; ++k) anb_s[(size_t)k * Np + s] = anb[(size_t)k * Np + c];
  }
}
__global__ void __launch_bounds__(TPB) halo_to_sweep_kernel(int N, int B, const double* __restrict__ phi, double* phi_s) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < B; j += gridDim.x * blockDim.x) phi_s[N + j] = phi[N + j];
}
__global__ void __launch_bounds__(TPB) from_sweep_kernel(int N, const int32_t* __restrict__ s2c, const double* __restrict__ phi_s, double* phi) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < N; s += gridDim.x * blockDim.x) phi[s2c[s]] = phi_s[s];
}

// residual in sweep space written per cell in block order (for per-block reductions)
template <int K>
__global__ void __launch_bounds__(TPB) residual_cells_kernel(int N, int Np, const int32_t* __restrict__ nbs, const int32_t* __restrict__ bpos,
                                                             const double* __restrict__ ap, const double* __restrict__ anb,
                                                             const double* __restrict__ b, const double* __restrict__ phi, double* rr) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < N; s += gridDim.x * blockDim.x) {
    double sumnb = b[s];
#pragma unroll
    for (int k = 0; k < K; ++k) sumnb = sumnb + anb[(size_t)k * Np + s] * phi[nbs[(size_t)k * Np + s]];
    rr[bpos[s]] = sumnb - ap[s] * phi[s];
  }
}
// one CTA per (chunk, block): sum r^2 and max(0, r) over a slice of a block's cells
#define SEG_CHUNK 8192
__global__ void __launch_bounds__(TPB) seg_reduce_kernel(const double* __restrict__ rr, const int32_t* __restrict__ blk_ptr, double* partial,
                                                         int nchunks_max) {
  const int blk = blockIdx.y, chunk = blockIdx.x;
  const int b0 = blk_ptr[blk] + chunk * SEG_CHUNK, b1 = min(blk_ptr[blk + 1], b0 + SEG_CHUNK);
  double s = 0.0, m = 0.0;
  for (int i = b0 + threadIdx.x; i < b1; i += blockDim.x) { const double r = rr[i]; s += r * r; m = fmax(m, r); }
  block_sum_max<TPB>(s, m);
  if (threadIdx.x == 0) { partial[2 * ((size_t)blk * nchunks_max + chunk)] = s; partial[2 * ((size_t)blk * nchunks_max + chunk) + 1] = m; }
}
__global__ void seg_final_kernel(const double* __restrict__ partial, int nchunks_max, int nblocks, double* out) {
  const int blk = threadIdx.x;
  if (blk >= nblocks) return;
  double s = 0.0, m = 0.0;
  for (int i = 0; i < nchunks_max; ++i) { s += partial[2 * ((size_t)blk * nchunks_max + i)]; m = fmax(m, partial[2 * ((size_t)blk * nchunks_max + i) + 1]); }
  out[2 * blk] = s; out[2 * blk + 1] = m;
}

// ---------------------------------------------------------------------------------------------
static int upload_schedule(Handle* h, const Schedule& S, DevSchedule& D) {
  D.nlevels = S.nlevels; D.nblocks = S.nblocks; D.nlag = (int)S.lag_src.size(); D.blk_ptr = S.blk_ptr;
  auto up = [&](int32_t*& dst, const std::vector<int32_t>& src) -> int {
    if (src.empty()) { dst = nullptr; return CFDL_OK; }
    CFDL_CUDA(cudaMalloc(&dst, sizeof(int32_t) * src.size()));
    h->allocs.push_back(dst);
    CFDL_CUDA(cudaMemcpy(dst, src.data(), sizeof(int32_t) * src.size(), cudaMemcpyHostToDevice));
    return CFDL_OK;
  };
  int rc;
  if ((rc = up(D.lvl_ptr, S.lvl_ptr))) return rc;
  if ((rc = up(D.s2c, S.s2c))) return rc;
  if ((rc = up(D.nbs, S.nbs))) return rc;
  if ((rc = up(D.bpos, S.bpos))) return rc;
  if ((rc = up(D.lag_src, S.lag_src))) return rc;
  D.max_level_cells = 0;
  for (int l = 0; l < S.nlevels; ++l) D.max_level_cells = std::max(D.max_level_cells, S.lvl_ptr[l + 1] - S.lvl_ptr[l]);
  return CFDL_OK;
}

template <int K>
static int coop_ctas_for(Handle* h) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, level_sgs_kernel<K>, TPB, 0);
  if (per_sm < 1) per_sm = 1;
  return h->num_sms * std::min(per_sm, 2);
}

int solver_init(Handle* h) {
  int rc;
  if ((rc = upload_schedule(h, h->prep.natural, h->natural))) return rc;
  if (h->prep.n_subdomains > 1 && (rc = upload_schedule(h, h->prep.blocks, h->blocks))) return rc;
  auto dalloc = [&](double*& p, size_t n) -> int {
    CFDL_CUDA(cudaMalloc(&p, sizeof(double) * n));
    h->allocs.push_back(p);
    CFDL_CUDA(cudaMemset(p, 0, sizeof(double) * n));
    return CFDL_OK;
  };
  if ((rc = dalloc(h->ap_s, h->Np))) return rc;
  if ((rc = dalloc(h->b_s, h->Np))) return rc;
  if ((rc = dalloc(h->anb_s, (size_t)h->K * h->Np))) return rc;
  if ((rc = dalloc(h->phi_s, (size_t)h->H + h->N + 32))) return rc;
  if ((rc = dalloc(h->rr, h->Np))) return rc;
  if (!h->rb_work && (rc = dalloc(h->rb_work, (size_t)h->H + 32))) return rc;
  h->coop_ctas = (h->K <= 4) ? coop_ctas_for<4>(h) : coop_ctas_for<6>(h);
  return CFDL_OK;
}

// launch the persistent level kernel for `niter` symmetric iterations
static int launch_levels(Handle* h, const DevSchedule& D, double sor, int niter) {
  CFDL_CUDA(cudaMemsetAsync(h->barrier, 0, sizeof(unsigned int), h->stream));
  int nlevels = D.nlevels, Np = h->Np, H = h->H, nlag = D.nlag;
  const int32_t *lvl_ptr = D.lvl_ptr, *nbs = D.nbs, *lag_src = D.lag_src;
  const double *ap = h->ap_s, *anb = h->anb_s, *b = h->b_s;
  double* phi = h->phi_s;
  unsigned int* counter = h->barrier;
  void* args[] = {&nlevels, &lvl_ptr, &Np, &H, &nbs, &ap, &anb, &b, &phi, &sor, &niter, &nlag, &lag_src, &counter};
  // no more CTAs than the widest level can feed
  int ctas = std::min(h->coop_ctas, std::max(1, (D.max_level_cells + TPB - 1) / TPB));
  auto go = [&](auto kern) { return cudaLaunchCooperativeKernel(kern, dim3(ctas), dim3(TPB), args, 0, S(h)); };
  prof_begin(h, PROF_LEVELS);
  CFDL_CUDA(h->K <= 4 ? go(level_sgs_kernel<4>) : go(level_sgs_kernel<6>));
  prof_end(h);
  return CFDL_OK;
}

template <int K>
static int parity_solve_t(Handle* h, int eq, double* phi, const double* rhs, int nit, double* out4, bool dispatch) {
  const bool is_pc = (eq == CFDL_EQ_PC);
  const double sor = is_pc ? 1.02 : 1.0;
  const bool multi = dispatch && h->prep.n_subdomains > 1;  // solve() dispatch, mod_solver.f90:338-342
  const DevSchedule& D = multi ? h->blocks : h->natural;
  const int N = h->N, Np = h->Np, g = grid_for(h, N, TPB);
  to_sweep_kernel<K><<<g, TPB, 0, S(h)>>>(N, Np, D.s2c, h->fld[CFDL_F_AP], h->fld[CFDL_F_ANB], rhs, phi, h->ap_s, h->anb_s,
                                               h->b_s, h->phi_s, h->H);
  if (h->B) halo_to_sweep_kernel<<<grid_for(h, h->B, TPB), TPB, 0, S(h)>>>(N, h->B, phi, h->phi_s);
  CFDL_CUDA(cudaGetLastError());
  int rc;
  int it = 0;
  double res_i = 0, res_f = 0, res_max = 0;
  if (!multi) {
    // solve_gs, mod_solver.f90:255-327
    CFDL_CUDA(cudaMemsetAsync(&h->ctl->ticket, 0, sizeof(unsigned int), h->stream));
    residual_kernel<K><<<g, TPB, 0, S(h)>>>(N, Np, D.nbs, h->ap_s, h->anb_s, h->b_s, h->phi_s, h->partial, h->ctl, RES_PLAIN, 0, h->scal + 8);
    CFDL_CUDA(cudaMemcpyAsync(h->scal_host, h->scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CFDL_CUDA(cudaStreamSynchronize(h->stream));
    res_i = h->scal_host[0];
    res_f = res_i;
    const double res_target = res_i / 10.0;
    while (it < nit && res_f > res_target) {
      it += 1;
      if ((rc = launch_levels(h, D, sor, 1))) return rc;
      residual_kernel<K><<<g, TPB, 0, S(h)>>>(N, Np, D.nbs, h->ap_s, h->anb_s, h->b_s, h->phi_s, h->partial, h->ctl, RES_PLAIN, 0, h->scal + 8);
      CFDL_CUDA(cudaMemcpyAsync(h->scal_host, h->scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CFDL_CUDA(cudaStreamSynchronize(h->stream));
      res_f = h->scal_host[0];
      res_max = h->scal_host[1];
    }
  } else {
    // multi_subdomain_solver, mod_solver.f90:124-189 (incl. the un-reset res_f_tot and signed res_max)
    const int P = D.nblocks;
    int maxlen = 0;
    for (int b = 0; b < P; ++b) maxlen = std::max(maxlen, D.blk_ptr[b + 1] - D.blk_ptr[b]);
    const int nchunks = (maxlen + SEG_CHUNK - 1) / SEG_CHUNK;
    if (2 * (size_t)P * nchunks > (size_t)h->partial_len || P > 64)
      return fail(CFDL_ERR_UNSUPPORTED, "too many subdomains (%d) for the residual work space", P);
    int32_t* blk_ptr_dev = (int32_t*)(h->scal + 16);  // small device scratch: P+1 ints
    CFDL_CUDA(cudaMemcpyAsync(blk_ptr_dev, D.blk_ptr.data(), sizeof(int32_t) * (P + 1), cudaMemcpyHostToDevice, h->stream));
    auto block_residuals = [&](double& sum_res2, double& mx) -> int {
      residual_cells_kernel<K><<<g, TPB, 0, S(h)>>>(N, Np, D.nbs, D.bpos, h->ap_s, h->anb_s, h->b_s, h->phi_s, h->rr);
      seg_reduce_kernel<<<dim3(nchunks, P), TPB, 0, S(h)>>>(h->rr, blk_ptr_dev, h->partial, nchunks);
      seg_final_kernel<<<1, 1024, 0, S(h)>>>(h->partial, nchunks, P, h->scal + 64);
      CFDL_CUDA(cudaMemcpyAsync(h->scal_host, h->scal + 64, 2 * P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CFDL_CUDA(cudaStreamSynchronize(h->stream));
      for (int b = 0; b < P; ++b) {
        const int ne_b = D.blk_ptr[b + 1] - D.blk_ptr[b];
        const double res = sqrt(h->scal_host[2 * b] / ne_b);
        sum_res2 = sum_res2 + res * res;
        mx = std::max(mx, h->scal_host[2 * b + 1]);
      }
      return CFDL_OK;
    };
    double res_i_tot = 0.0, res_max_tot = 0.0;
    if ((rc = block_residuals(res_i_tot, res_max_tot))) return rc;
    res_i_tot = sqrt(res_i_tot / P);
    double res_f_tot = res_i_tot;
    const double res_target = res_i_tot / 10.0;
    while (it < nit && res_f_tot > res_target) {
      if ((rc = launch_levels(h, D, sor, 2))) return rc;  // smoother_gs(...,2) per block, then update_halos
      it += 2;
      if (it % 10 == 0) {
        if ((rc = block_residuals(res_f_tot, res_max_tot))) return rc;  // accumulates onto the old value, :175
        res_f_tot = sqrt(res_f_tot / P);
      }
    }
    res_i = res_i_tot; res_f = res_f_tot; res_max = res_max_tot;
  }
  from_sweep_kernel<<<g, TPB, 0, S(h)>>>(N, D.s2c, h->phi_s, phi);
  CFDL_CUDA(cudaGetLastError());
  if (out4) { out4[0] = it; out4[1] = res_i; out4[2] = res_f; out4[3] = res_max; }
  return CFDL_OK;
}

#include "kernels_rb.inc"
#include "kernels_rb3.inc"
#include "kernels_pcg.inc"

// MCSGS: colour-ordered symmetric Gauss-Seidel with the reference's stopping rule; iterations
// are enqueued in growing batches, each kernel returning at once when ctl->done is set
template <int K>
static int mcsgs_solve_t(Handle* h, int eq, double* phi, const double* rhs, int nit, double* out4) {
  const double sor = (eq == CFDL_EQ_PC) ? 1.02 : 1.0;
  const int N = h->N, Np = h->Np, g = grid_for(h, N, TPB);
  const int nc = h->prep.ncolors;
  const int32_t* cp = h->prep.color_ptr.data();
  const double *ap = h->fld[CFDL_F_AP], *anb = h->fld[CFDL_F_ANB];
  SolveCtl init = {};
  init.nit = nit;
  *h->ctl_host = init;
  CFDL_CUDA(cudaMemcpyAsync(h->ctl, h->ctl_host, sizeof(SolveCtl), cudaMemcpyHostToDevice, h->stream));
  // Several GPUs: the colouring is global, and the ghost copies of a colour are refreshed right
  // after that colour's sweep, so every update sees exactly the operands of the single-GPU
  // sweep; the residual norm is all-reduced (sum of r^2, max) before the stopping test.
  const bool dist = h->prep.nranks > 1;
  int rc;
  auto residual = [&](int mode) -> int {
    prof_begin(h, PROF_RESIDUAL);
    if (!dist) {
      residual_kernel<K><<<g, TPB, 0, S(h)>>>(N, Np, h->ell_nb, ap, anb, rhs, phi, h->partial, h->ctl, mode, 0, nullptr);
    } else {
      residual_kernel<K><<<g, TPB, 0, S(h)>>>(N, Np, h->ell_nb, ap, anb, rhs, phi, h->partial, h->ctl,
                                              mode == RES_ITER ? RES_LOCAL_GUARDED : RES_LOCAL, 0, h->scal + 8);
      int r = comm_allreduce_sum_max(h, h->scal + 8);
      if (r) return r;
      finalize_residual_kernel<<<1, 1, 0, S(h)>>>(h->ctl, h->scal + 8, (double)h->ne_global, mode);
    }
    prof_end(h);
    return CFDL_OK;
  };
  auto sweep = [&](int c) -> int {
    const int n = cp[c + 1] - cp[c];
    if (n > 0) {
      prof_begin(h, PROF_SGS_SWEEP);
      sgs_range_kernel<K><<<grid_for(h, n, TPB), TPB, 0, S(h)>>>(cp[c], cp[c + 1], Np, h->ell_nb, ap, anb, rhs, phi, sor, h->ctl);
      prof_end(h);
    }
    return dist ? comm_exchange(h, phi, 1, c) : CFDL_OK;
  };
  if (dist && (rc = comm_exchange(h, phi, 1, -1))) return rc;
  if ((rc = residual(RES_INIT))) return rc;
  // first batch: the iteration count of this equation's previous solve (consecutive SIMPLE
  // iterations need about the same), enqueued without waiting for the opening residual — every
  // kernel returns at once when ctl->done is set; an overshoot continues in small batches
  int launched = 0, batch = std::max(1, std::min(h->last_passes[eq], nit));
  for (bool opening = true;; opening = false) {
    if (!opening) {
      CFDL_CUDA(cudaMemcpyAsync(h->ctl_host, h->ctl, sizeof(SolveCtl), cudaMemcpyDeviceToHost, h->stream));
      CFDL_CUDA(cudaStreamSynchronize(h->stream));
      if (h->ctl_host->done || launched >= nit) break;
    }
    const int m = std::min(batch, nit - launched);
    for (int i = 0; i < m; ++i) {
      for (int c = 0; c < nc; ++c) if ((rc = sweep(c))) return rc;
      for (int c = nc - 1; c >= 0; --c) if ((rc = sweep(c))) return rc;
      if ((rc = residual(RES_ITER))) return rc;
    }
    CFDL_CUDA(cudaGetLastError());
    launched += m;
    batch = (launched > 8) ? 8 : 2;
  }
  h->last_passes[eq] = h->ctl_host->it;
  if (out4) { out4[0] = h->ctl_host->it; out4[1] = h->ctl_host->res_i; out4[2] = h->ctl_host->res_f; out4[3] = h->ctl_host->res_max; }
  return CFDL_OK;
}

// calc_residual (mod_solver.f90:230-253) on the device-numbered system: RMS of r and max(0, r)
int residual_plain(Handle* h, const double* phi, const double* rhs, bool signed_max, double* res, double* res_max) {
  if (h->K > 6) return fail(CFDL_ERR_UNSUPPORTED, "cells with more than 6 faces are not supported");
  const int g = grid_for(h, h->N, TPB);
  CFDL_CUDA(cudaMemsetAsync(&h->ctl->ticket, 0, sizeof(unsigned int), h->stream));
  if (h->K <= 4)
    residual_kernel<4><<<g, TPB, 0, S(h)>>>(h->N, h->Np, h->ell_nb, h->fld[CFDL_F_AP], h->fld[CFDL_F_ANB], rhs, phi, h->partial, h->ctl, RES_PLAIN, signed_max ? 1 : 0, h->scal + 8);
  else
    residual_kernel<6><<<g, TPB, 0, S(h)>>>(h->N, h->Np, h->ell_nb, h->fld[CFDL_F_AP], h->fld[CFDL_F_ANB], rhs, phi, h->partial, h->ctl, RES_PLAIN, signed_max ? 1 : 0, h->scal + 8);
  CFDL_CUDA(cudaMemcpyAsync(h->scal_host, h->scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CFDL_CUDA(cudaStreamSynchronize(h->stream));
  *res = h->scal_host[0];
  *res_max = h->scal_host[1];
  return CFDL_OK;
}

// The three momentum solves of solve_uvwp (mod_uvwp.f90:114-116): side by side (kernels_rb3.inc) when
// the mode allows it — multicolour SGS on a two-colour mesh, one GPU — else one after the other.
// Both orders give the same bits.
static bool momentum_fusable(const Handle* h) {
  if (h->K > 6 || h->solver_mode == CFDL_SOLVER_PARITY) return false;
  if (h->prep.nranks == 1) return true;
  // partitioned: the fused two-colour passes with the peer-to-peer exchange (slab with the extra value arrays)
  return h->fused_rb && h->prep.ncolors == 2 && h->p2p.connected && h->use_p2p && h->rb3_work[0] != nullptr;
}
static int momentum_run(Handle* h, bool fused, int nit, double* out12) {
  if (fused && h->fused_rb && h->prep.ncolors == 2) return h->K <= 4 ? rb3_solve_t<4>(h, nit, out12) : rb3_solve_t<6>(h, nit, out12);
  if (fused) return h->K <= 4 ? mcsgs3_solve_t<4>(h, nit, out12) : mcsgs3_solve_t<6>(h, nit, out12);
  for (int eq = CFDL_EQ_U; eq <= CFDL_EQ_W; ++eq) {
    int rc = solve_equation(h, eq, h->fld[CFDL_F_U + eq], h->fld[CFDL_F_BU + eq], nit, out12 ? out12 + 4 * eq : nullptr, false);
    if (rc) return rc;
  }
  return CFDL_OK;
}

int solve_momentum(Handle* h, int nit, double* out12) {
  // side by side wherever the mode allows it (measured on a B200, round 2: 0.54 against 0.79 ms at 128^3);
  // uvw_fused = 0 keeps the reference's one-after-the-other order (same bits, tests compare the two)
  return momentum_run(h, momentum_fusable(h) && h->uvw_fused != 0, nit, out12);
}

int solve_equation(Handle* h, int eq, double* phi, const double* rhs, int nit, double* out4, bool dispatch) {
  if (h->K > 6) return fail(CFDL_ERR_UNSUPPORTED, "cells with more than 6 faces are not supported");
  const bool k4 = h->K <= 4;
  if (h->prep.nranks > 1 && h->solver_mode == CFDL_SOLVER_PARITY)
    return fail(CFDL_ERR_UNSUPPORTED, "the exact natural-order solver is sequential across partitions; use solver=mcsgs on several GPUs");
  switch (h->solver_mode) {
    case CFDL_SOLVER_PARITY:
      return k4 ? parity_solve_t<4>(h, eq, phi, rhs, nit, out4, dispatch) : parity_solve_t<6>(h, eq, phi, rhs, nit, out4, dispatch);
    case CFDL_SOLVER_PCG:  // conjugate gradients for the (symmetric) pc system; the momentum equations stay on MCSGS
      if (eq == CFDL_EQ_PC) return k4 ? pcg_solve_t<4>(h, eq, phi, rhs, nit, out4) : pcg_solve_t<6>(h, eq, phi, rhs, nit, out4);
      // fall through
    case CFDL_SOLVER_MCSGS:
      if (h->fused_rb && h->prep.ncolors == 2)
        return k4 ? rb_solve_t<4>(h, eq, phi, rhs, nit, out4) : rb_solve_t<6>(h, eq, phi, rhs, nit, out4);
      return k4 ? mcsgs_solve_t<4>(h, eq, phi, rhs, nit, out4) : mcsgs_solve_t<6>(h, eq, phi, rhs, nit, out4);
  }
  return fail(CFDL_ERR_ARG, "unknown solver mode %d", h->solver_mode);
}

}  // namespace cfdl
