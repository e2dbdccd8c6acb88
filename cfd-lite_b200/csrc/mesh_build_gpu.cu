// cfdl_mesh_build_gpu: connectivity + geometry of the mesh on the GPU — the arrays the reference's set-up produces:
// find_element_nb (src/setup/mod_mg_lvl_uns.f90:283-433), calc_aip_xyzip_uns (src/setup/calc_aip_xyzip.f90:7-75),
// calc_vol_cv_centers_uns (src/setup/calc_vol_cv_centers.f90:4-61), including the reference's global face numbering, bit for bit
// what the host builder cfdl_mesh_build and the reference's own source give (tests).
//
// No sort of face keys is needed.  Two elements that share a face see the same vertex set, hence the same SMALLEST vertex, and
// the reference numbers faces in the lexicographic order of (smallest vertex, e1, f1) — the order in which its loop over the
// vertices first meets them.  So:
//   1. every (element, local face) record is counted at its smallest vertex (histogram with atomics), an exclusive scan turns
//      the counts into bucket offsets, the records are scattered into the buckets (arrival order: arbitrary);
//   2. one thread per vertex sorts its bucket by (element, face) — a few dozen entries — and pairs the records whose other
//      vertices agree; exactly two records per face, else the mesh is refused as the reference does (`stop`);
//   3. a scan over the pairs per vertex gives the global face numbers; a second pass per vertex writes s2g, ef2nb(:,1:2), bs;
//   4. boundary-edge flags (the sign of bs), face area vectors / centroids, cell volumes / centroids and halo centres with the
//      reference's formulas (geom_formulas.h, shared with the host builder: same operations, same bits).
// Everything is deterministic: atomics only decide the arrival order inside a bucket, which step 2 sorts away.
#include <algorithm>
#include <vector>
#include "state.h"
#include "geom_formulas.h"

namespace cfdl {
namespace {

#define MB_TPB 256
#define MB_CUDA(call)                                                                                          \
  do {                                                                                                         \
    cudaError_t err__ = (call);                                                                                \
    if (err__ != cudaSuccess) { rc = fail(CFDL_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); goto done; } \
  } while (0)

__constant__ int c_faces[4][6][4] = {
    {{1, 3, 2, 0}, {1, 2, 4, 0}, {2, 3, 4, 0}, {3, 1, 4, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},   // TETRA_4 (10)
    {{1, 4, 3, 2}, {1, 2, 5, 0}, {2, 3, 5, 0}, {3, 4, 5, 0}, {4, 1, 5, 0}, {0, 0, 0, 0}},   // PYRA_5 (12)
    {{1, 2, 5, 4}, {2, 3, 6, 5}, {3, 1, 4, 6}, {1, 3, 2, 0}, {4, 5, 6, 0}, {0, 0, 0, 0}},   // PENTA_6 (14)
    {{1, 4, 3, 2}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 4, 8, 7}, {1, 5, 8, 4}, {5, 6, 7, 8}}};  // HEXA_8 (17)

__host__ __device__ inline int mb_nface(int t) { return t == 17 ? 6 : t == 10 ? 4 : (t == 12 || t == 14) ? 5 : (t == 5 || t == 7) ? 1 : 0; }
__host__ __device__ inline int mb_nvx(int t) { return t == 17 ? 8 : t == 10 ? 4 : t == 12 ? 5 : t == 14 ? 6 : t == 5 ? 3 : t == 7 ? 4 : 0; }
__device__ inline int mb_tab(int t) { return t == 10 ? 0 : t == 12 ? 1 : t == 14 ? 2 : 3; }

// vertex list of local face f (0-based) of an element of CGNS type t, in the reference's order (mod_util.f90:55-85,1362-1426)
__device__ inline int mb_face_vx(const int32_t* vx, int t, int f, int32_t* lst) {
  if (t == 5) { lst[0] = vx[0]; lst[1] = vx[1]; lst[2] = vx[2]; return 3; }
  if (t == 7) { lst[0] = vx[0]; lst[1] = vx[1]; lst[2] = vx[2]; lst[3] = vx[3]; return 4; }
  const int tb = mb_tab(t);
  int nl = 0;
  for (int l = 0; l < 4; ++l)
    if (c_faces[tb][f][l] > 0) lst[nl++] = vx[c_faces[tb][f][l] - 1];
  return nl;
}
// the face's vertices sorted ascending; a triangle gets 0 in front (as the host builder's key)
__device__ inline void mb_sorted(const int32_t* vx, int t, int f, int32_t* s) {
  int32_t lst[4] = {0, 0, 0, 0};
  mb_face_vx(vx, t, f, lst);
  for (int i = 1; i < 4; ++i) { const int32_t v = lst[i]; int j = i - 1; while (j >= 0 && lst[j] > v) { lst[j + 1] = lst[j]; --j; } lst[j + 1] = v; }
  for (int i = 0; i < 4; ++i) s[i] = lst[i];
}

struct MeshDev {
  int64_t nvx, nelem, nrec;
  int32_t ne, nf, nbf, w, nsec;
  const int32_t *e2vx, *sec_lo, *sec_hi, *sec_type;
  const double *x, *y, *z;
  int32_t* ef2nb_idx;   // ne + 1 (1-based offsets)
  unsigned int *vcount, *vstart, *vcursor, *pcount, *fbase;  // per vertex (nvx + 1)
  unsigned long long* bucket;  // per record: (element << 3) | local face (0-based)
  int32_t* tuple;              // per record: the three larger vertices of its sorted vertex set
  int32_t* partner;            // per bucket position: position of the matching record
  int32_t *ef2nb_nb, *ef2nb_fg, *s2g, *bs;
  double *xc, *yc, *zc, *aip, *rip, *vol;
  int* err;                    // 1: a face without partner / shared by more than two elements / bad vertex id
};

__device__ inline int mb_type_of(const MeshDev& M, int64_t e) {
  for (int s = 0; s < M.nsec; ++s)
    if (e >= M.sec_lo[s] && e <= M.sec_hi[s]) return M.sec_type[s];
  return 0;
}

// faces per cell -> ef2nb_idx before its scan (entry e holds the face count of cell e; entry 0 stays 0)
__global__ void __launch_bounds__(MB_TPB) mb_nface_kernel(const MeshDev M, unsigned int* cnt) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1; e <= M.ne; e += (int64_t)gridDim.x * blockDim.x) cnt[e - 1] = (unsigned)mb_nface(mb_type_of(M, e));
}
__device__ inline int64_t mb_rec_of(const MeshDev& M, int64_t e, int f) { return e <= M.ne ? (int64_t)M.ef2nb_idx[e - 1] - 1 + f : (int64_t)M.ef2nb_idx[M.ne] - 1 + (e - M.ne - 1); }

__global__ void __launch_bounds__(MB_TPB) mb_count_kernel(const MeshDev M) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1; e <= M.nelem; e += (int64_t)gridDim.x * blockDim.x) {
    const int t = mb_type_of(M, e);
    const int32_t* vx = M.e2vx + (size_t)M.w * (e - 1);
    for (int f = 0; f < mb_nface(t); ++f) {
      int32_t s[4];
      mb_sorted(vx, t, f, s);
      const int32_t vmin = s[0] ? s[0] : s[1];
      if (vmin < 1 || s[3] > M.nvx) { *M.err = 1; continue; }
      atomicAdd(&M.vcount[vmin - 1], 1u);
    }
  }
}
__global__ void __launch_bounds__(MB_TPB) mb_scatter_kernel(const MeshDev M) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1; e <= M.nelem; e += (int64_t)gridDim.x * blockDim.x) {
    const int t = mb_type_of(M, e);
    const int32_t* vx = M.e2vx + (size_t)M.w * (e - 1);
    for (int f = 0; f < mb_nface(t); ++f) {
      int32_t s[4];
      mb_sorted(vx, t, f, s);
      const int32_t vmin = s[0] ? s[0] : s[1];
      if (vmin < 1 || s[3] > M.nvx) continue;
      const unsigned pos = M.vstart[vmin - 1] + atomicAdd(&M.vcursor[vmin - 1], 1u);
      M.bucket[pos] = ((unsigned long long)e << 3) | (unsigned)f;
      // the rest of the key: a quad's three larger vertices, a triangle's two (first entry 0)
      M.tuple[3 * (size_t)pos] = s[0] ? s[1] : 0; M.tuple[3 * (size_t)pos + 1] = s[2]; M.tuple[3 * (size_t)pos + 2] = s[3];
    }
  }
}
// one thread per vertex: order its bucket by (element, face), pair the records with equal keys, count the pairs
__global__ void __launch_bounds__(MB_TPB) mb_match_kernel(const MeshDev M) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < M.nvx; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned b0 = M.vstart[v], b1 = M.vstart[v + 1];
    for (unsigned i = b0 + 1; i < b1; ++i) {  // insertion sort, tuple moves with its record
      const unsigned long long key = M.bucket[i];
      const int32_t t0 = M.tuple[3 * (size_t)i], t1 = M.tuple[3 * (size_t)i + 1], t2 = M.tuple[3 * (size_t)i + 2];
      unsigned j = i;
      while (j > b0 && M.bucket[j - 1] > key) {
        M.bucket[j] = M.bucket[j - 1];
        M.tuple[3 * (size_t)j] = M.tuple[3 * (size_t)(j - 1)]; M.tuple[3 * (size_t)j + 1] = M.tuple[3 * (size_t)(j - 1) + 1]; M.tuple[3 * (size_t)j + 2] = M.tuple[3 * (size_t)(j - 1) + 2];
        --j;
      }
      M.bucket[j] = key; M.tuple[3 * (size_t)j] = t0; M.tuple[3 * (size_t)j + 1] = t1; M.tuple[3 * (size_t)j + 2] = t2;
    }
    for (unsigned i = b0; i < b1; ++i) M.partner[i] = -1;
    unsigned npairs = 0;
    for (unsigned i = b0; i < b1; ++i) {
      if (M.partner[i] >= 0) continue;
      int found = 0;
      for (unsigned j = i + 1; j < b1; ++j) {
        if (M.tuple[3 * (size_t)j] != M.tuple[3 * (size_t)i] || M.tuple[3 * (size_t)j + 1] != M.tuple[3 * (size_t)i + 1] || M.tuple[3 * (size_t)j + 2] != M.tuple[3 * (size_t)i + 2]) continue;
        if (found || M.partner[j] >= 0) { found = 2; break; }  // a third element on the same face
        M.partner[i] = (int32_t)j; M.partner[j] = (int32_t)i;
        found = 1;
      }
      // exactly two records per face, the first of them a cell (two boundary elements on one face are refused)
      if (found != 1 || (int64_t)(M.bucket[i] >> 3) > M.ne) *M.err = 1;
      else ++npairs;
    }
    M.pcount[v] = npairs;
  }
}
__device__ inline int32_t mb_pack(int64_t g, int s) { return (int32_t)(((uint32_t)g << 5) | (uint32_t)s); }
__global__ void __launch_bounds__(MB_TPB) mb_write_kernel(const MeshDev M) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < M.nvx; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned b0 = M.vstart[v], b1 = M.vstart[v + 1];
    int32_t fg = (int32_t)M.fbase[v];
    for (unsigned i = b0; i < b1; ++i) {
      const int32_t j = M.partner[i];
      if (j < 0 || (unsigned)j < i) continue;  // second record of its pair (or an error already flagged)
      ++fg;  // 1-based global face number, in the order (smallest vertex, e1, f1)
      const int64_t e1 = (int64_t)(M.bucket[i] >> 3), e2 = (int64_t)(M.bucket[j] >> 3);
      const int f1 = (int)(M.bucket[i] & 7) + 1, f2 = (int)(M.bucket[j] & 7) + 1;
      const int64_t i1 = (int64_t)M.ef2nb_idx[e1 - 1] - 1 + f1 - 1;
      M.s2g[fg - 1] = mb_pack(e1, f1);  // owner = lower-numbered cell
      M.ef2nb_fg[i1] = fg;
      if (e2 > M.ne) {  // boundary: the 2-D element is the halo cell
        M.ef2nb_nb[i1] = mb_pack(e2, 0);
        M.bs[e2 - M.ne - 1] = mb_pack(e1, f1);
      } else {
        const int64_t i2 = (int64_t)M.ef2nb_idx[e2 - 1] - 1 + f2 - 1;
        M.ef2nb_nb[i1] = mb_pack(e2, f2);
        M.ef2nb_nb[i2] = mb_pack(e1, f1);
        M.ef2nb_fg[i2] = -fg;
      }
    }
  }
}
// "edge boundary" flag in the sign of bs (mod_mg_lvl_uns.f90:421-433): the reference's sequential loop leaves bs negative
// exactly for the halos of cells with two or more boundary faces (consumers take abs())
__global__ void __launch_bounds__(MB_TPB) mb_bs_sign_kernel(const MeshDev M) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1; e <= M.ne; e += (int64_t)gridDim.x * blockDim.x) {
    int nb = 0;
    for (int32_t idx = M.ef2nb_idx[e - 1]; idx <= M.ef2nb_idx[e] - 1; ++idx) {
      const uint32_t p = (uint32_t)M.ef2nb_nb[idx - 1];
      if ((p & 31u) == 0 && (int64_t)(p >> 5) > M.ne) ++nb;
    }
    if (nb < 2) continue;
    for (int32_t idx = M.ef2nb_idx[e - 1]; idx <= M.ef2nb_idx[e] - 1; ++idx) {
      const uint32_t p = (uint32_t)M.ef2nb_nb[idx - 1];
      if ((p & 31u) == 0 && (int64_t)(p >> 5) > M.ne) { int32_t& b = M.bs[(p >> 5) - M.ne - 1]; b = b < 0 ? b : -b; }
    }
  }
}
// face area vectors and centroids (calc_aip_xyzip.f90:25-72), owner's vertex order
__global__ void __launch_bounds__(MB_TPB) mb_face_geom_kernel(const MeshDev M) {
  for (int64_t fg = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; fg < M.nf; fg += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = (int64_t)((uint32_t)M.s2g[fg] >> 5);
    const int fl = M.s2g[fg] & 31;
    int32_t lst[4];
    const int nl = mb_face_vx(M.e2vx + (size_t)M.w * (e - 1), mb_type_of(M, e), fl - 1, lst);
    double r[4][3];
    for (int i = 0; i < nl; ++i) { r[i][0] = M.x[lst[i] - 1]; r[i][1] = M.y[lst[i] - 1]; r[i][2] = M.z[lst[i] - 1]; }
    face_area_centroid(r, nl, M.aip + 3 * (size_t)fg, M.rip + 3 * (size_t)fg);
  }
}
// cell volumes / centroids by face pyramids; halo centre = face centroid (calc_vol_cv_centers.f90:20-55)
__global__ void __launch_bounds__(MB_TPB) mb_cell_geom_kernel(const MeshDev M) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1; e <= M.ne; e += (int64_t)gridDim.x * blockDim.x) {
    const int nv = mb_nvx(mb_type_of(M, e));
    const int32_t* vx = M.e2vx + (size_t)M.w * (e - 1);
    CellAccumulator acc;
    double* gc = acc.gc;
    gc[0] = gc[1] = gc[2] = 0.0;
    for (int i = 0; i < nv; ++i) { gc[0] = gc[0] + M.x[vx[i] - 1]; gc[1] = gc[1] + M.y[vx[i] - 1]; gc[2] = gc[2] + M.z[vx[i] - 1]; }
    for (int i = 0; i < 3; ++i) gc[i] = gc[i] / nv;
    for (int32_t idx = M.ef2nb_idx[e - 1]; idx <= M.ef2nb_idx[e] - 1; ++idx) {
      int32_t gf = M.ef2nb_fg[idx - 1];
      const int sg = gf >= 0 ? 1 : -1;
      gf = gf < 0 ? -gf : gf;
      const double* cs = M.rip + 3 * (size_t)(gf - 1);
      const double* a = M.aip + 3 * (size_t)(gf - 1);
      const uint32_t p = (uint32_t)M.ef2nb_nb[idx - 1];
      if ((p & 31u) == 0) { const size_t enb = p >> 5; M.xc[enb - 1] = cs[0]; M.yc[enb - 1] = cs[1]; M.zc[enb - 1] = cs[2]; }
      acc.add_face(sg, a, cs);
    }
    double ctr[3];
    acc.finish(ctr, &M.vol[e - 1]);
    M.xc[e - 1] = ctr[0]; M.yc[e - 1] = ctr[1]; M.zc[e - 1] = ctr[2];
  }
}

// ---- exclusive scan of unsigned counts (n entries in, n + 1 out: out[n] = total) -------------------------------------------
#define SCAN_PER_CTA 1024  // 256 threads x 4 entries
__global__ void __launch_bounds__(MB_TPB) mb_scan_block_kernel(const unsigned int* in, unsigned int* out, unsigned int* sums, int64_t n) {
  __shared__ unsigned int sh[MB_TPB];
  const int64_t base = (int64_t)blockIdx.x * SCAN_PER_CTA + 4 * (int64_t)threadIdx.x;
  unsigned int v[4], t = 0;
  for (int i = 0; i < 4; ++i) { v[i] = (base + i < n) ? in[base + i] : 0u; t += v[i]; }
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int off = 1; off < MB_TPB; off <<= 1) {  // Hillis-Steele inclusive scan of the per-thread sums
    unsigned int a = threadIdx.x >= (unsigned)off ? sh[threadIdx.x - off] : 0u;
    __syncthreads();
    sh[threadIdx.x] += a;
    __syncthreads();
  }
  unsigned int run = sh[threadIdx.x] - t;  // exclusive prefix of this thread inside the CTA
  for (int i = 0; i < 4; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
  if (threadIdx.x == MB_TPB - 1) sums[blockIdx.x] = sh[MB_TPB - 1];
}
__global__ void __launch_bounds__(MB_TPB) mb_scan_add_kernel(unsigned int* out, const unsigned int* offs, int64_t n) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_PER_CTA + 4 * (int64_t)threadIdx.x;
  const unsigned int o = offs[blockIdx.x];
  for (int i = 0; i < 4; ++i) if (base + i < n) out[base + i] += o;
}
// out[0..n) = exclusive scan of in[0..n); returns the total through *total_dev (a device word, may alias out[n])
int mb_scan(const unsigned int* in, unsigned int* out, int64_t n, cudaStream_t st, std::vector<void*>& tmp) {
  const int64_t nb = (n + SCAN_PER_CTA - 1) / SCAN_PER_CTA;
  unsigned int* sums = nullptr;
  if (cudaMalloc(&sums, sizeof(unsigned int) * (size_t)(2 * nb + 2)) != cudaSuccess) return fail(CFDL_ERR_CUDA, "cfdl_mesh_build_gpu: cudaMalloc failed");
  tmp.push_back(sums);
  mb_scan_block_kernel<<<(unsigned)nb, MB_TPB, 0, st>>>(in, out, sums, n);
  if (nb > 1) {
    unsigned int* sums_scanned = sums + nb + 1;
    int rc = mb_scan(sums, sums_scanned, nb, st, tmp);
    if (rc) return rc;
    mb_scan_add_kernel<<<(unsigned)nb, MB_TPB, 0, st>>>(out, sums_scanned, n);
  }
  return CFDL_OK;
}
__global__ void mb_total_kernel(const unsigned int* counts, const unsigned int* excl, int64_t n, unsigned int* total) { *total = n > 0 ? excl[n - 1] + counts[n - 1] : 0u; }
__global__ void __launch_bounds__(MB_TPB) mb_idx_kernel(const unsigned int* excl, int32_t* ef2nb_idx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) ef2nb_idx[i] = (int32_t)excl[i] + 1;
}

}  // namespace
}  // namespace cfdl

using namespace cfdl;

extern "C" int cfdl_mesh_build_gpu(int32_t device, int64_t nvx, const double* x, const double* y, const double* z, int nsec, const int32_t* etype,
                                   const int32_t* esec, int w, const int32_t* e2vx, int32_t ne, int32_t nf, int32_t nbf, int32_t* ef2nb_idx,
                                   int32_t* ef2nb_nb, int32_t* ef2nb_fg, int32_t* s2g, int32_t* bs, double* xc, double* yc, double* zc, double* aip,
                                   double* rip, double* vol) {
  const int64_t nelem = (int64_t)ne + nbf;
  if (!x || !y || !z || !etype || !esec || !e2vx || !ef2nb_idx || !ef2nb_nb || !ef2nb_fg || !s2g || (!bs && nbf) || !xc || !yc || !zc || !aip || !rip || !vol)
    return fail(CFDL_ERR_ARG, "cfdl_mesh_build_gpu: NULL argument");
  if (nelem >= (int64_t(1) << 26)) return fail(CFDL_ERR_RANGE, "cfdl_mesh_build_gpu: ne+nbf=%lld exceeds the reference's 2^26 packing limit", (long long)nelem);
  if (nsec < 1 || nsec > 64) return fail(CFDL_ERR_ARG, "cfdl_mesh_build_gpu: %d sections", nsec);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(CFDL_ERR_CUDA, "cfdl_mesh_build_gpu: no CUDA device is usable (this library has no CPU path; cfdl_mesh_build is the host tool)"); }
  if (device < 0 || device >= ndev) return fail(CFDL_ERR_ARG, "cfdl_mesh_build_gpu: device %d of %d", device, ndev);
  // sections: contiguous, 3-D first (checked on the host: a handful of entries)
  std::vector<int32_t> lo(nsec), hi(nsec), ty(nsec);
  {
    std::vector<int> order(nsec);
    for (int s = 0; s < nsec; ++s) order[s] = s;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return esec[2 * a] < esec[2 * b]; });
    int64_t next = 1, slots = 0;
    for (int k = 0; k < nsec; ++k) {
      const int s = order[k];
      if (esec[2 * s] != next) return fail(CFDL_ERR_MESH, "cfdl_mesh_build_gpu: sections are not contiguous");
      if (mb_nface(etype[s]) == 0) return fail(CFDL_ERR_UNSUPPORTED, "cfdl_mesh_build_gpu: element type %d", etype[s]);
      next = (int64_t)esec[2 * s + 1] + 1;
      if ((etype[s] >= 10) != (esec[2 * s + 1] <= ne)) return fail(CFDL_ERR_MESH, "cfdl_mesh_build_gpu: 3-D sections must come first and hold ne cells");
      if (etype[s] >= 10) slots += (int64_t)mb_nface(etype[s]) * (esec[2 * s + 1] - esec[2 * s] + 1);
      lo[k] = esec[2 * s]; hi[k] = esec[2 * s + 1]; ty[k] = etype[s];
    }
    if (next != nelem + 1) return fail(CFDL_ERR_MESH, "cfdl_mesh_build_gpu: sections hold %lld elements, expected %lld", (long long)(next - 1), (long long)nelem);
    if (slots != 2 * (int64_t)nf - nbf) return fail(CFDL_ERR_MESH, "cfdl_mesh_build_gpu: nf=%d inconsistent with the element faces (%lld slots)", nf, (long long)slots);
  }
  const int64_t Z = 2 * (int64_t)nf - nbf, nrec = Z + nbf, H = nelem;
  int rc = CFDL_OK;
  std::vector<void*> tmp;
  cudaStream_t st = nullptr;
  int herr = 0;
  unsigned int htotal = 0;
  MeshDev M = {};
  auto dalloc = [&](auto*& p, size_t n) -> bool {
    void* q = nullptr;
    if (cudaMalloc(&q, sizeof(*p) * (n + 8)) != cudaSuccess) return false;
    tmp.push_back(q);
    p = reinterpret_cast<decltype(p)>(q);
    return true;
  };
  int32_t *d_e2vx = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_ty = nullptr;
  double *d_x = nullptr, *d_y = nullptr, *d_z = nullptr;
  unsigned int *d_cnt = nullptr, *d_excl = nullptr, *d_total = nullptr;
  int grid = 0;
  MB_CUDA(cudaSetDevice(device));
  {
    cudaDeviceProp prop;
    MB_CUDA(cudaGetDeviceProperties(&prop, device));
    grid = prop.multiProcessorCount * 8;
  }
  MB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  if (!dalloc(d_e2vx, (size_t)w * nelem) || !dalloc(d_lo, nsec) || !dalloc(d_hi, nsec) || !dalloc(d_ty, nsec) || !dalloc(d_x, nvx) || !dalloc(d_y, nvx) ||
      !dalloc(d_z, nvx) || !dalloc(M.ef2nb_idx, (size_t)ne + 1) || !dalloc(M.vcount, nvx + 1) || !dalloc(M.vstart, nvx + 1) || !dalloc(M.vcursor, nvx + 1) ||
      !dalloc(M.pcount, nvx + 1) || !dalloc(M.fbase, nvx + 1) || !dalloc(M.bucket, nrec) || !dalloc(M.tuple, 3 * (size_t)nrec) || !dalloc(M.partner, nrec) ||
      !dalloc(M.ef2nb_nb, Z) || !dalloc(M.ef2nb_fg, Z) || !dalloc(M.s2g, nf) || !dalloc(M.bs, (size_t)nbf) || !dalloc(M.xc, H) || !dalloc(M.yc, H) ||
      !dalloc(M.zc, H) || !dalloc(M.aip, 3 * (size_t)nf) || !dalloc(M.rip, 3 * (size_t)nf) || !dalloc(M.vol, ne) || !dalloc(M.err, 1) ||
      !dalloc(d_cnt, (size_t)ne + 1) || !dalloc(d_excl, (size_t)ne + 2) || !dalloc(d_total, 1)) {
    cudaGetLastError();
    rc = fail(CFDL_ERR_CUDA, "cfdl_mesh_build_gpu: out of device memory");
    goto done;
  }
  M.nvx = nvx; M.nelem = nelem; M.nrec = nrec; M.ne = ne; M.nf = nf; M.nbf = nbf; M.w = w; M.nsec = nsec;
  M.e2vx = d_e2vx; M.sec_lo = d_lo; M.sec_hi = d_hi; M.sec_type = d_ty; M.x = d_x; M.y = d_y; M.z = d_z;
  MB_CUDA(cudaMemcpyAsync(d_e2vx, e2vx, sizeof(int32_t) * (size_t)w * nelem, cudaMemcpyHostToDevice, st));
  MB_CUDA(cudaMemcpyAsync(d_lo, lo.data(), sizeof(int32_t) * nsec, cudaMemcpyHostToDevice, st));
  MB_CUDA(cudaMemcpyAsync(d_hi, hi.data(), sizeof(int32_t) * nsec, cudaMemcpyHostToDevice, st));
  MB_CUDA(cudaMemcpyAsync(d_ty, ty.data(), sizeof(int32_t) * nsec, cudaMemcpyHostToDevice, st));
  MB_CUDA(cudaMemcpyAsync(d_x, x, sizeof(double) * nvx, cudaMemcpyHostToDevice, st));
  MB_CUDA(cudaMemcpyAsync(d_y, y, sizeof(double) * nvx, cudaMemcpyHostToDevice, st));
  MB_CUDA(cudaMemcpyAsync(d_z, z, sizeof(double) * nvx, cudaMemcpyHostToDevice, st));
  MB_CUDA(cudaMemsetAsync(M.err, 0, sizeof(int), st));
  MB_CUDA(cudaMemsetAsync(M.vcount, 0, sizeof(unsigned int) * (size_t)(nvx + 1), st));
  MB_CUDA(cudaMemsetAsync(M.vcursor, 0, sizeof(unsigned int) * (size_t)(nvx + 1), st));
  MB_CUDA(cudaMemsetAsync(M.bs, 0, sizeof(int32_t) * (size_t)(nbf + 1), st));
  // ef2nb_idx = 1 + exclusive scan of the faces per cell
  mb_nface_kernel<<<grid, MB_TPB, 0, st>>>(M, d_cnt);
  MB_CUDA(cudaMemsetAsync(d_cnt + ne, 0, sizeof(unsigned int), st));
  if ((rc = mb_scan(d_cnt, d_excl, (int64_t)ne + 1, st, tmp))) goto done;
  mb_idx_kernel<<<grid, MB_TPB, 0, st>>>(d_excl, M.ef2nb_idx, ne);
  // 1. records bucketed by the smallest vertex of their face
  mb_count_kernel<<<grid, MB_TPB, 0, st>>>(M);
  if ((rc = mb_scan(M.vcount, M.vstart, nvx + 1, st, tmp))) goto done;  // vcount[nvx] = 0: vstart[nvx] = number of records
  mb_scatter_kernel<<<grid, MB_TPB, 0, st>>>(M);
  // 2. pairs per vertex, 3. global face numbers
  mb_match_kernel<<<grid, MB_TPB, 0, st>>>(M);
  MB_CUDA(cudaMemsetAsync(M.pcount + nvx, 0, sizeof(unsigned int), st));
  if ((rc = mb_scan(M.pcount, M.fbase, nvx + 1, st, tmp))) goto done;
  mb_total_kernel<<<1, 1, 0, st>>>(M.pcount, M.fbase, nvx + 1, d_total);
  MB_CUDA(cudaMemcpyAsync(&herr, M.err, sizeof(int), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(&htotal, d_total, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  if (herr || (int64_t)htotal != nf) {
    rc = fail(CFDL_ERR_MESH, "Error in creation of element neighbour list: %u faces matched, %d expected%s", htotal, nf,
              herr ? " (a face without partner, a face shared by more than two elements, or a vertex id out of range)" : "");
    goto done;
  }
  mb_write_kernel<<<grid, MB_TPB, 0, st>>>(M);
  mb_bs_sign_kernel<<<grid, MB_TPB, 0, st>>>(M);
  // 4. geometry
  mb_face_geom_kernel<<<grid, MB_TPB, 0, st>>>(M);
  mb_cell_geom_kernel<<<grid, MB_TPB, 0, st>>>(M);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaMemcpyAsync(ef2nb_idx, M.ef2nb_idx, sizeof(int32_t) * ((size_t)ne + 1), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(ef2nb_nb, M.ef2nb_nb, sizeof(int32_t) * (size_t)Z, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(ef2nb_fg, M.ef2nb_fg, sizeof(int32_t) * (size_t)Z, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(s2g, M.s2g, sizeof(int32_t) * (size_t)nf, cudaMemcpyDeviceToHost, st));
  if (nbf) MB_CUDA(cudaMemcpyAsync(bs, M.bs, sizeof(int32_t) * (size_t)nbf, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(xc, M.xc, sizeof(double) * (size_t)H, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(yc, M.yc, sizeof(double) * (size_t)H, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(zc, M.zc, sizeof(double) * (size_t)H, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(aip, M.aip, sizeof(double) * 3 * (size_t)nf, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(rip, M.rip, sizeof(double) * 3 * (size_t)nf, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(vol, M.vol, sizeof(double) * (size_t)ne, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  for (int32_t j = 0; j < nbf; ++j)
    if (bs[j] == 0) { rc = fail(CFDL_ERR_MESH, "cfdl_mesh_build_gpu: 2-D element %d matches no cell face", ne + 1 + j); break; }
done:
  for (void* p : tmp) cudaFree(p);
  if (st) cudaStreamDestroy(st);
  return rc;
}
