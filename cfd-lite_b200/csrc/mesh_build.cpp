// Fast connectivity + geometry build: produces exactly the arrays the reference's set-up
// produces — find_element_nb (src/setup/mod_mg_lvl_uns.f90:283-433), calc_aip_xyzip_uns
// (src/setup/calc_aip_xyzip.f90:7-75), calc_vol_cv_centers_uns
// (src/setup/calc_vol_cv_centers.f90:4-61) — including the reference's global face numbering,
// but by sorting face keys (O(F log F)) instead of the reference's per-vertex pairwise
// comparison (O(nvx k^2)).
//
// Face numbering: the reference numbers a face when it first meets it in `do v=1,nvx` /
// `do l=1,n; do m=l+1,n`; the incident (element, local face) list of a vertex is ordered by
// element id then face id, so the discovery order is the lexicographic order of
// (smallest vertex of the face, e1, f1, e2, f2) with e1 < e2.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "cfdl_common.h"
#include "geom_formulas.h"

using cfdl::face_area_centroid;
using cfdl::CellAccumulator;

namespace {

const int faces_tetra4[4][4] = {{1, 3, 2, 0}, {1, 2, 4, 0}, {2, 3, 4, 0}, {3, 1, 4, 0}};
const int faces_pyra5[5][4] = {{1, 4, 3, 2}, {1, 2, 5, 0}, {2, 3, 5, 0}, {3, 4, 5, 0}, {4, 1, 5, 0}};
const int faces_penta6[5][4] = {{1, 2, 5, 4}, {2, 3, 6, 5}, {3, 1, 4, 6}, {1, 3, 2, 0}, {4, 5, 6, 0}};
const int faces_hexa8[6][4] = {{1, 4, 3, 2}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 4, 8, 7}, {1, 5, 8, 4}, {5, 6, 7, 8}};

int nface_of(int t) { return t == 17 ? 6 : t == 10 ? 4 : (t == 12 || t == 14) ? 5 : (t == 5 || t == 7) ? 1 : 0; }
int nvx_of(int t) { return t == 17 ? 8 : t == 10 ? 4 : t == 12 ? 5 : t == 14 ? 6 : t == 5 ? 3 : t == 7 ? 4 : 0; }

// vertex list of local face f (0-based) of an element of CGNS type t (mod_util.f90:55-85,1362-1426)
int face_vx(const int32_t* vx, int t, int f, int32_t* lst) {
  const int(*tab)[4] = nullptr;
  switch (t) {
    case 10: tab = faces_tetra4; break;
    case 12: tab = faces_pyra5; break;
    case 14: tab = faces_penta6; break;
    case 17: tab = faces_hexa8; break;
    case 5: lst[0] = vx[0]; lst[1] = vx[1]; lst[2] = vx[2]; return 3;
    case 7: lst[0] = vx[0]; lst[1] = vx[1]; lst[2] = vx[2]; lst[3] = vx[3]; return 4;
    default: return 0;
  }
  int nl = 0;
  for (int l = 0; l < 4; ++l)
    if (tab[f][l] > 0) lst[nl++] = vx[tab[f][l] - 1];
  return nl;
}

struct FaceRec {
  int32_t v[4];  // sorted ascending, 0-padded in front for triangles
  int32_t e, f;  // 1-based element, 1-based local face
};
inline bool key_less(const FaceRec& a, const FaceRec& b) {
  for (int i = 0; i < 4; ++i) if (a.v[i] != b.v[i]) return a.v[i] < b.v[i];
  return a.e < b.e;
}
inline bool key_eq(const FaceRec& a, const FaceRec& b) { return !std::memcmp(a.v, b.v, sizeof a.v); }

struct Pair {
  int32_t vmin, e1, f1, e2, f2;
};


}  // namespace

extern "C" int cfdl_mesh_build(int64_t nvx, const double* x, const double* y, const double* z, int nsec, const int32_t* etype,
                               const int32_t* esec, int w, const int32_t* e2vx, int32_t ne, int32_t nf, int32_t nbf,
                               int32_t* ef2nb_idx, int32_t* ef2nb_nb, int32_t* ef2nb_fg, int32_t* s2g, int32_t* bs, double* xc,
                               double* yc, double* zc, double* aip, double* rip, double* vol) {
  using cfdl::fail;
  const int64_t nelem = (int64_t)ne + nbf;
  if (nelem >= (int64_t(1) << 26)) return fail(CFDL_ERR_RANGE, "cfdl_mesh_build: ne+nbf=%lld exceeds the reference's 2^26 packing limit", (long long)nelem);
  // element -> type, sections must be contiguous, 3-D first
  std::vector<int8_t> et((size_t)nelem + 1, 0);
  {
    std::vector<int> order(nsec);
    for (int s = 0; s < nsec; ++s) order[s] = s;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return esec[2 * a] < esec[2 * b]; });
    int64_t next = 1;
    for (int k = 0; k < nsec; ++k) {
      int s = order[k];
      if (esec[2 * s] != next) return fail(CFDL_ERR_MESH, "cfdl_mesh_build: sections are not contiguous");
      if (nface_of(etype[s]) == 0) return fail(CFDL_ERR_UNSUPPORTED, "cfdl_mesh_build: element type %d", etype[s]);
      for (int64_t e = esec[2 * s]; e <= esec[2 * s + 1]; ++e) et[e] = (int8_t)etype[s];
      next = (int64_t)esec[2 * s + 1] + 1;
      if ((etype[s] >= 10) != (esec[2 * s + 1] <= ne)) return fail(CFDL_ERR_MESH, "cfdl_mesh_build: 3-D sections must come first and hold ne cells");
    }
    if (next != nelem + 1) return fail(CFDL_ERR_MESH, "cfdl_mesh_build: sections hold %lld elements, expected %lld", (long long)(next - 1), (long long)nelem);
  }
  ef2nb_idx[0] = 1;
  for (int32_t e = 1; e <= ne; ++e) ef2nb_idx[e] = ef2nb_idx[e - 1] + nface_of(et[e]);
  const int64_t Z = 2 * (int64_t)nf - nbf;
  if (ef2nb_idx[ne] - 1 != Z) return fail(CFDL_ERR_MESH, "cfdl_mesh_build: nf=%d inconsistent with the element faces (%d slots)", nf, ef2nb_idx[ne] - 1);
  // all (element, face) records keyed by sorted vertex set
  std::vector<FaceRec> rec((size_t)(Z + nbf));
  size_t nr = 0;
  for (int64_t e = 1; e <= nelem; ++e) {
    const int t = et[e];
    const int32_t* vx = e2vx + (size_t)w * (e - 1);
    for (int f = 0; f < nface_of(t); ++f) {
      FaceRec& r = rec[nr++];
      int32_t lst[4] = {0, 0, 0, 0};
      int nl = face_vx(vx, t, f, lst);
      for (int i = 0; i < nl; ++i) if (lst[i] < 1 || lst[i] > nvx) return fail(CFDL_ERR_MESH, "cfdl_mesh_build: element %lld has vertex %d", (long long)e, lst[i]);
      std::sort(lst, lst + 4);
      std::memcpy(r.v, lst, sizeof lst);
      r.e = (int32_t)e; r.f = f + 1;
    }
  }
  std::sort(rec.begin(), rec.end(), key_less);
  std::vector<Pair> pairs((size_t)nf);
  size_t np = 0;
  for (size_t i = 0; i < nr;) {
    size_t j = i + 1;
    while (j < nr && key_eq(rec[i], rec[j])) ++j;
    if (j - i != 2) return fail(CFDL_ERR_MESH, "Error in creation of element neighbour list: a face of element %d is shared by %d elements", rec[i].e, (int)(j - i));
    if (rec[i].e > ne) return fail(CFDL_ERR_MESH, "cfdl_mesh_build: two boundary elements cover the same face");
    if (np >= (size_t)nf) return fail(CFDL_ERR_MESH, "cfdl_mesh_build: more faces than nf");
    Pair& p = pairs[np++];
    p.vmin = rec[i].v[0] ? rec[i].v[0] : rec[i].v[1];
    p.e1 = rec[i].e; p.f1 = rec[i].f; p.e2 = rec[j - 1].e; p.f2 = rec[j - 1].f;
    i = j;
  }
  if (np != (size_t)nf) return fail(CFDL_ERR_MESH, "Error in creation of element neighbour list ... %d %d", (int)np, nf);
  rec.clear(); rec.shrink_to_fit();
  std::sort(pairs.begin(), pairs.end(), [](const Pair& a, const Pair& b) {
    if (a.vmin != b.vmin) return a.vmin < b.vmin;
    if (a.e1 != b.e1) return a.e1 < b.e1;
    if (a.f1 != b.f1) return a.f1 < b.f1;
    if (a.e2 != b.e2) return a.e2 < b.e2;
    return a.f2 < b.f2;
  });
  auto pack = [](int32_t g, int32_t s) { return (int32_t)(((uint32_t)g << 5) | (uint32_t)s); };
  for (int32_t j = 0; j < nbf; ++j) bs[j] = 0;
  for (int32_t fgi = 0; fgi < nf; ++fgi) {
    const Pair& p = pairs[fgi];
    const int32_t fg = fgi + 1;
    const int32_t i1 = ef2nb_idx[p.e1 - 1] - 1 + p.f1 - 1;
    s2g[fgi] = pack(p.e1, p.f1);  // owner = lower-numbered cell
    ef2nb_fg[i1] = fg;
    if (p.e2 > ne) {              // boundary: 2-D element is the halo cell
      ef2nb_nb[i1] = pack(p.e2, 0);
      bs[p.e2 - ne - 1] = pack(p.e1, p.f1);
    } else {
      const int32_t i2 = ef2nb_idx[p.e2 - 1] - 1 + p.f2 - 1;
      ef2nb_nb[i1] = pack(p.e2, p.f2);
      ef2nb_nb[i2] = pack(p.e1, p.f1);
      ef2nb_fg[i2] = -fg;
    }
  }
  pairs.clear(); pairs.shrink_to_fit();
  for (int32_t j = 0; j < nbf; ++j) if (bs[j] == 0) return fail(CFDL_ERR_MESH, "cfdl_mesh_build: 2-D element %d matches no cell face", ne + 1 + j);
  // "edge boundary" flag in the sign of bs (mod_mg_lvl_uns.f90:421-433); consumers take abs()
  for (int32_t e1 = ne + 1; e1 <= ne + nbf; ++e1) {
    if (bs[e1 - ne - 1] < 0) continue;
    const int32_t e = (int32_t)((uint32_t)bs[e1 - ne - 1] >> 5), f = bs[e1 - ne - 1] & 31;
    for (int32_t idx = ef2nb_idx[e - 1]; idx <= ef2nb_idx[e] - 1; ++idx) {
      if (idx - ef2nb_idx[e - 1] + 1 == f) continue;
      const int32_t e2 = (int32_t)((uint32_t)ef2nb_nb[idx - 1] >> 5), f2 = ef2nb_nb[idx - 1] & 31;
      if (f2 == 0 && e2 > ne) {
        bs[e1 - ne - 1] = -std::abs(bs[e1 - ne - 1]);
        bs[e2 - ne - 1] = -bs[e2 - ne - 1];
      }
    }
  }
  // ---- face area vectors and centroids (calc_aip_xyzip.f90:25-72), owner's vertex order ----
  auto P = [&](int32_t v, double* r) { r[0] = x[v - 1]; r[1] = y[v - 1]; r[2] = z[v - 1]; };
  for (int32_t fg = 0; fg < nf; ++fg) {
    const int32_t e = (int32_t)((uint32_t)s2g[fg] >> 5), fl = s2g[fg] & 31;
    int32_t lst[4];
    const int nl = face_vx(e2vx + (size_t)w * (e - 1), et[e], fl - 1, lst);
    double r[4][3];
    for (int i = 0; i < nl; ++i) P(lst[i], r[i]);
    face_area_centroid(r, nl, aip + 3 * (size_t)fg, rip + 3 * (size_t)fg);
  }
  // ---- cell volumes / centroids by face pyramids; halo centre = face centroid ----------------
  for (int32_t e = 1; e <= ne; ++e) {
    const int nv = nvx_of(et[e]);
    const int32_t* vx = e2vx + (size_t)w * (e - 1);
    CellAccumulator acc;
    double* gc = acc.gc;
    gc[0] = gc[1] = gc[2] = 0.0;
    for (int i = 0; i < nv; ++i) { gc[0] = gc[0] + x[vx[i] - 1]; gc[1] = gc[1] + y[vx[i] - 1]; gc[2] = gc[2] + z[vx[i] - 1]; }
    for (int i = 0; i < 3; ++i) gc[i] = gc[i] / nv;
    for (int32_t idx = ef2nb_idx[e - 1]; idx <= ef2nb_idx[e] - 1; ++idx) {
      int32_t gf = ef2nb_fg[idx - 1];
      const int sg = gf >= 0 ? 1 : -1;
      gf = std::abs(gf);
      const double* cs = rip + 3 * (size_t)(gf - 1);
      const double* a = aip + 3 * (size_t)(gf - 1);
      const int32_t enb = (int32_t)((uint32_t)ef2nb_nb[idx - 1] >> 5), lf = ef2nb_nb[idx - 1] & 31;
      if (lf == 0) { xc[enb - 1] = cs[0]; yc[enb - 1] = cs[1]; zc[enb - 1] = cs[2]; }
      acc.add_face(sg, a, cs);
    }
    double ctr[3];
    acc.finish(ctr, &vol[e - 1]);
    xc[e - 1] = ctr[0]; yc[e - 1] = ctr[1]; zc[e - 1] = ctr[2];
  }
  return CFDL_OK;
}
