// Host-side partitioning: (1) the reference's recursive-coordinate-bisection block
// decomposition and block-local cell order (src/setup/mod_agglomeration.f90:380-561 `split_leaf`
// / `grow`, src/setup/mod_mg_lvl_uns.f90:883-902 + src/modules/mod_util.f90:1683-1730 for the
// order inside a block), used for the pc block solver and as the GPU partition of the
// multi-GPU path; (2) extraction of one rank's partition (owned cells + ghost cells + halos)
// from the global mesh, with the interface lists for the halo exchange.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>
#include "cfdl_common.h"

namespace {

struct Leaf {
  std::vector<int32_t> cells;  // ascending original ids (0-based)
  double lo[3], hi[3], vol;
};

// key/payload quicksort with the reference's exact (unstable) behaviour: first element is the
// pivot, partitions are scanned from both ends, the smaller partition is processed first.
void sort_by_key_like_reference(std::vector<int32_t>& key, std::vector<int32_t>& val) {
  const int n = (int)key.size();
  std::vector<int> beg(1000), end(1000);
  int i = 0;
  beg[0] = 0; end[0] = n;
  while (i >= 0) {
    int L = beg[i], R = end[i] - 1;
    if (L < R) {
      const int32_t piv = key[L], v0 = val[L];
      while (L < R) {
        while (key[R] >= piv && L < R) --R;
        if (L < R) { key[L] = key[R]; val[L] = val[R]; ++L; }
        while (key[L] <= piv && L < R) ++L;
        if (L < R) { key[R] = key[L]; val[R] = val[L]; --R; }
      }
      key[L] = piv; val[L] = v0;
      beg[i + 1] = L + 1; end[i + 1] = end[i]; end[i] = L;
      ++i;
      if (i >= 999) return;
      if (end[i] - beg[i] > end[i - 1] - beg[i - 1]) { std::swap(beg[i], beg[i - 1]); std::swap(end[i], end[i - 1]); }
    } else {
      --i;
    }
  }
}

}  // namespace

// cell2sub(ne): block 1..P of every cell; g2gf_p(ne): cells (1-based) sorted by block in the
// reference's order; g2gf_idx(P+1): 1-based block offsets.  Any output may be NULL.
extern "C" int cfdl_partition_rcb(int32_t ne, const double* xc, const double* yc, const double* zc, const double* vol, int32_t P,
                                  int32_t* cell2sub, int32_t* g2gf_p, int32_t* g2gf_idx) {
  using cfdl::fail;
  if (ne < 1 || P < 1 || P > ne || !xc || !yc || !zc || !vol) return fail(CFDL_ERR_ARG, "cfdl_partition_rcb: bad arguments");
  const double* pos[3] = {xc, yc, zc};
  std::vector<Leaf> leaves(1);
  Leaf& root = leaves[0];
  root.cells.resize(ne);
  std::iota(root.cells.begin(), root.cells.end(), 0);
  double total = 0.0;
  for (int a = 0; a < 3; ++a) { root.lo[a] = *std::min_element(pos[a], pos[a] + ne); root.hi[a] = *std::max_element(pos[a], pos[a] + ne); }
  for (int32_t e = 0; e < ne; ++e) total = total + vol[e] * 1.0;
  root.vol = 2.1 * total;                 // construct_seed_gen: the root always splits
  const double vol_ave = total * 1 / P;   // grow(): vol_ave*(target_ncv)/(new_target_ncv)
  double threshold = 2.0;
  int guard = 0;
  while ((int)leaves.size() < P) {
    const size_t before = leaves.size();
    size_t i = 0;
    while (i < leaves.size()) {
      Leaf& lf = leaves[i];
      if (lf.vol / vol_ave < threshold || lf.cells.size() == 1) { ++i; continue; }
      double d = 0.0;
      int axis = -1;
      for (int a = 0; a < 3; ++a)
        if (std::fabs(lf.lo[a] - lf.hi[a]) > d) { d = std::fabs(lf.lo[a] - lf.hi[a]); axis = a; }
      if (axis < 0) return fail(CFDL_ERR_MESH, "cfdl_partition_rcb: degenerate bounding box");
      d = (lf.lo[axis] + lf.hi[axis]) / 2.0;
      Leaf l, r;
      for (int a = 0; a < 3; ++a) { l.lo[a] = r.lo[a] = 1e20; l.hi[a] = r.hi[a] = -1e20; }
      l.vol = r.vol = 0.0;
      for (int32_t e : lf.cells) {
        Leaf& t = (pos[axis][e] < d) ? l : r;
        t.cells.push_back(e);
        t.vol = t.vol + vol[e] * 1.0;
        for (int a = 0; a < 3; ++a) { t.lo[a] = std::min(t.lo[a], pos[a][e]); t.hi[a] = std::max(t.hi[a], pos[a][e]); }
      }
      if (l.cells.empty() || r.cells.empty()) return fail(CFDL_ERR_MESH, "cfdl_partition_rcb: one-sided split");
      leaves[i] = std::move(l);
      leaves.insert(leaves.begin() + i + 1, std::move(r));
      if ((int)leaves.size() == P) break;
      i += 2;  // children created in this pass are not revisited until the next one
    }
    if (leaves.size() == before) threshold *= 0.75;
    if (++guard > 100000) return fail(CFDL_ERR_MESH, "cfdl_partition_rcb: cannot reach %d blocks", P);
  }
  std::vector<int32_t> key(ne), val(ne);
  for (int b = 0; b < P; ++b)
    for (int32_t e : leaves[b].cells) key[e] = b + 1;
  if (cell2sub) std::copy(key.begin(), key.end(), cell2sub);
  if (g2gf_p || g2gf_idx) {
    std::iota(val.begin(), val.end(), 1);
    sort_by_key_like_reference(key, val);
    if (g2gf_p) std::copy(val.begin(), val.end(), g2gf_p);
    if (g2gf_idx) {
      int32_t g = 0;
      for (int32_t i = 0; i < ne; ++i)
        if (key[i] != g) { g = key[i]; g2gf_idx[g - 1] = i + 1; }
      g2gf_idx[P] = ne + 1;
    }
  }
  return CFDL_OK;
}

// ---- partition plan (host only; what cfdl_create_distributed derives internally) ---------------
#include "prep.h"

// For rank `rank` of `nranks`: owned cells and ghost cells in device order (1-based ids of the
// global mesh), the neighbour ranks, and per neighbour the cells sent to it / the slice of the
// ghost range received from it.  send_ptr/recv_ptr have n_nbr+1 entries; the sender's list and
// the receiver's ghost slice enumerate the same cells in the same order on both ranks.
extern "C" int cfdl_partition_plan(int32_t ne, int32_t nf, int32_t nbf, const int32_t* ef2nb_idx, const int32_t* ef2nb_nb,
                                   const int32_t* ef2nb_fg, const int32_t* s2g, const int32_t* bs, const double* xc,
                                   const double* yc, const double* zc, const int32_t* cell2rank, int32_t nranks, int32_t rank,
                                   int32_t* n_owned, int32_t* owned, int32_t* n_ghost, int32_t* ghost, int32_t* n_nbr,
                                   int32_t* nbr_rank, int32_t* send_ptr, int32_t* send_cells, int32_t* recv_ptr,
                                   int32_t* ncolors, int32_t* owned_color_ptr) {
  using namespace cfdl;
  Prep p;
  int rc = prepare(p, ne, nf, nbf, ef2nb_idx, ef2nb_nb, ef2nb_fg, s2g, bs, xc, yc, zc, 0, nullptr, nullptr, nullptr, 1, nullptr, nullptr, 2,
                   cell2rank, rank, nranks);
  if (rc) return rc;
  const int nnbr = (int)p.nbr_rank.size(), nc = p.ncolors;
  if (n_owned) *n_owned = p.N;
  if (n_ghost) *n_ghost = p.G;
  if (n_nbr) *n_nbr = nnbr;
  if (ncolors) *ncolors = nc;
  if (owned) for (int32_t c = 0; c < p.N; ++c) owned[c] = p.c2o[c] + 1;
  if (ghost) for (int32_t g = 0; g < p.G; ++g) ghost[g] = p.c2o[p.N + g] + 1;
  if (nbr_rank) for (int r = 0; r < nnbr; ++r) nbr_rank[r] = p.nbr_rank[r];
  if (send_ptr) for (int r = 0; r <= nnbr; ++r) send_ptr[r] = p.send_ptr[(size_t)r * nc];
  if (recv_ptr) for (int r = 0; r <= nnbr; ++r) recv_ptr[r] = p.recv_ptr[(size_t)r * nc];
  if (send_cells) for (size_t i = 0; i < p.send_cells.size(); ++i) send_cells[i] = p.c2o[p.send_cells[i]] + 1;
  if (owned_color_ptr) for (int c = 0; c <= nc; ++c) owned_color_ptr[c] = p.color_ptr[c];
  return CFDL_OK;
}
