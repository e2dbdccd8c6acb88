// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_setup.hpp).  Parity pinned to the reference's source text (tests/test_oracle_vs_reference_source.py).
// Restates the reference's serial mesh set-up; each function cites the lines it follows.
#include "oracle_setup.hpp"
#include <cstring>

namespace orc {

// ---- element tables: mod_util.f90:55-85 (faces, CGNS order), :166-169 (counts) -------------
static const int faces_tetra4[4][3] = {{1, 3, 2}, {1, 2, 4}, {2, 3, 4}, {3, 1, 4}};
static const int faces_pyra5[5][4] = {{1, 4, 3, 2}, {1, 2, 5, 0}, {2, 3, 5, 0}, {3, 4, 5, 0}, {4, 1, 5, 0}};
static const int nfaces_pyra5[5] = {4, 3, 3, 3, 3};
static const int faces_penta6[5][4] = {{1, 2, 5, 4}, {2, 3, 6, 5}, {3, 1, 4, 6}, {1, 3, 2, 0}, {4, 5, 6, 0}};
static const int nfaces_penta6[5] = {4, 4, 4, 3, 3};
static const int faces_hexa8[6][4] = {{1, 4, 3, 2}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 4, 8, 7}, {1, 5, 8, 4}, {5, 6, 7, 8}};

int element_nface(int t) {
  static const int tab[21] = {0, 6, 0, 0, 0, 1, 1, 1, 1, 1, 4, 4, 5, 5, 5, 5, 5, 6, 6, 6, 0};
  return (t >= 1 && t <= 20) ? tab[t] : 0;
}
int element_nvx(int t) {
  static const int tab[21] = {0, 8, 1, 2, 3, 3, 6, 4, 8, 9, 4, 10, 5, 14, 6, 15, 18, 8, 20, 27, 0};
  return (t >= 1 && t <= 20) ? tab[t] : 0;
}

void face_vxlist(const int* vxlist, int t, int f, int* lst, int& nl) {
  if (t == TETRA_4) {
    nl = 3;
    for (int l = 0; l < nl; ++l) lst[l] = vxlist[faces_tetra4[f - 1][l] - 1];
  } else if (t == PYRA_5) {
    nl = nfaces_pyra5[f - 1];
    for (int l = 0; l < nl; ++l) lst[l] = vxlist[faces_pyra5[f - 1][l] - 1];
  } else if (t == PENTA_6) {
    nl = nfaces_penta6[f - 1];
    for (int l = 0; l < nl; ++l) lst[l] = vxlist[faces_penta6[f - 1][l] - 1];
  } else if (t == HEXA_8) {
    nl = 4;
    for (int l = 0; l < nl; ++l) lst[l] = vxlist[faces_hexa8[f - 1][l] - 1];
  } else if (t == TRI_3) {
    nl = 3;
    for (int l = 0; l < nl; ++l) lst[l] = vxlist[l];
  } else if (t == QUAD_4) {
    nl = 4;
    for (int l = 0; l < nl; ++l) lst[l] = vxlist[l];
  } else {
    throw std::runtime_error("Unsupported element type in face_vxlist");  // mod_util.f90:1421-1424
  }
}

// face_compare, mod_util.f90:1450-1479 : same vertex SET (order-free)
static bool face_compare(int e1, int f1, int t1, const int* vx1, int e2, int f2, int t2, const int* vx2) {
  if (e1 == e2) return false;
  int lst1[4] = {0, 0, 0, 0}, lst2[4] = {0, 0, 0, 0}, nl1, nl2;
  face_vxlist(vx1, t1, f1, lst1, nl1);
  face_vxlist(vx2, t2, f2, lst2, nl2);
  if (nl1 != nl2) return false;
  int cnt = 0;
  for (int i = 0; i < nl1; ++i)
    for (int j = 0; j < nl2; ++j)
      if (lst1[i] == lst2[j]) { ++cnt; break; }
  return cnt == nl1;
}

// ---- find_element_nb, mod_mg_lvl_uns.f90:283-488 -----------------------------------------
void find_element_nb(Mesh& m) {
  const int ng = m.ne, nelem = m.nelem, nvx = m.nvx, nsec = m.nsec;
  const int Z = 2 * m.nf - m.nbf;
  m.ef2nb_idx.alloc(ng + 1, 0);
  m.ef2nb1.alloc(Z, 0);
  m.ef2nb2.alloc(Z, 0);
  m.s2g.alloc(m.nf + 1, 0);
  m.bs.alloc(m.nbf, 0, ng + 1);
  A1<int> vx2e_idx(nvx + 1, 0);
  m.ef2nb_idx(1) = 1;
  int fmax = 0;
  for (int s = 1; s <= nsec; ++s)  // :303-314
    for (int e = m.bs_idx(s); e <= m.bs_idx(s + 1) - 1; ++e) {
      int t = m.etype[s - 1];
      if (e < ng + 1) m.ef2nb_idx(e + 1) = m.ef2nb_idx(e) + element_nface(t);
      fmax = std::max(fmax, element_nface(t));
      for (int j = 1; j <= element_nvx(t); ++j) vx2e_idx(m.e2vx_of(e)[j - 1]) += 1;
    }
  if (m.ef2nb_idx(ng + 1) - 1 != Z) throw std::runtime_error("find_element_nb: face count does not match sections");
  int tmp = vx2e_idx(1), emax = tmp;  // :316-324 exclusive scan, 1-based
  vx2e_idx(1) = 1;
  for (int v = 2; v <= nvx + 1; ++v) {
    int tmp2 = vx2e_idx(v);
    emax = std::max(emax, tmp2);
    vx2e_idx(v) = vx2e_idx(v - 1) + tmp;
    tmp = tmp2;
  }
  A1<int> vx2e(vx2e_idx(nvx + 1) - 1, 0);
  {  // :332-345 (first free slot == append in visiting order)
    std::vector<int> cur((size_t)nvx + 1);
    for (int v = 1; v <= nvx; ++v) cur[v] = vx2e_idx(v);
    for (int s = 1; s <= nsec; ++s)
      for (int e = m.bs_idx(s); e <= m.bs_idx(s + 1) - 1; ++e)
        for (int j = 1; j <= element_nvx(m.etype[s - 1]); ++j) {
          int v = m.e2vx_of(e)[j - 1];
          vx2e(cur[v]++) = e;
        }
  }
  std::vector<int> indices((size_t)emax * fmax + 1), f2s((size_t)emax * fmax + 1);
  int nf = 0;
  for (int v = 1; v <= nvx; ++v) {  // :351-420
    int n = 0;
    for (int idx1 = vx2e_idx(v); idx1 <= vx2e_idx(v + 1) - 1; ++idx1) {
      int e = vx2e(idx1);
      int s = m.section_of(e);
      int t = m.etype[s - 1];
      for (int f = 1; f <= element_nface(t); ++f) {
        int lst1[4] = {0, 0, 0, 0}, nl1;
        face_vxlist(m.e2vx_of(e), t, f, lst1, nl1);
        for (int l = 0; l < nl1; ++l)
          if (lst1[l] == v) { ++n; indices[n] = index_t(e, f); f2s[n] = s; }
      }
    }
    for (int l = 1; l <= n; ++l)
      for (int mm = l + 1; mm <= n; ++mm) {
        int t1 = m.etype[f2s[l] - 1], t2 = m.etype[f2s[mm] - 1];
        int e1, f1, e2, f2;
        get_idx(indices[l], e1, f1);
        get_idx(indices[mm], e2, f2);
        if (!face_compare(e1, f1, t1, m.e2vx_of(e1), e2, f2, t2, m.e2vx_of(e2))) continue;
        if (e1 > ng && e2 <= ng) {
          int idx2 = m.ef2nb_idx(e2) + f2 - 1;
          if (m.ef2nb1(idx2) != 0) continue;
          ++nf;
          if (nf > m.nf) throw std::runtime_error("find_element_nb: more faces than nf");
          m.ef2nb1(idx2) = index_t(e1, 0);
          m.ef2nb2(idx2) = nf;
          m.s2g(nf) = indices[mm];
          m.bs(e1) = index_t(e2, f2);
        } else if (e2 > ng && e1 <= ng) {
          int idx1 = m.ef2nb_idx(e1) + f1 - 1;
          if (m.ef2nb1(idx1) != 0) continue;
          ++nf;
          if (nf > m.nf) throw std::runtime_error("find_element_nb: more faces than nf");
          m.ef2nb1(idx1) = index_t(e2, 0);
          m.ef2nb2(idx1) = nf;
          m.s2g(nf) = indices[l];
          m.bs(e2) = index_t(e1, f1);
        } else if (e1 <= ng && e2 <= ng) {
          int idx1 = m.ef2nb_idx(e1) + f1 - 1, idx2 = m.ef2nb_idx(e2) + f2 - 1;
          if (m.ef2nb1(idx1) != 0 && m.ef2nb1(idx2) != 0) continue;
          ++nf;
          if (nf > m.nf) throw std::runtime_error("find_element_nb: more faces than nf");
          m.ef2nb1(idx1) = indices[mm];
          m.ef2nb1(idx2) = indices[l];
          m.ef2nb2(idx1) = nf;
          m.s2g(nf) = indices[l];
          if (e1 > e2) { m.ef2nb2(idx1) = -nf; m.s2g(nf) = indices[mm]; }
          m.ef2nb2(idx2) = -m.ef2nb2(idx1);
        }
      }
  }
  // :421-433 "edge boundary" flag in the sign bit of bs; every consumer takes abs()
  for (int e1 = ng + 1; e1 <= nelem; ++e1) {
    if (m.bs(e1) < 0) continue;
    int e, f;
    get_idx(m.bs(e1), e, f);
    if (e < 1 || e > ng) continue;  // unmatched 2-D element: reported below
    for (int idx = m.ef2nb_idx(e); idx <= m.ef2nb_idx(e + 1) - 1; ++idx) {
      if (idx - m.ef2nb_idx(e) + 1 == f) continue;
      int e2, f2;
      get_idx(m.ef2nb1(idx), e2, f2);
      if (f2 == 0 && e2 > ng) {
        m.bs(e1) = -std::abs(m.bs(e1));
        m.bs(e2) = -m.bs(e2);
      }
    }
  }
  if (nf != m.nf) {  // :456-485
    char msg[128];
    std::snprintf(msg, sizeof msg, "Error in creation of element neighbour list ... %d %d", nf, m.nf);
    throw std::runtime_error(msg);
  }
}

// ---- calc_aip_xyzip_uns, calc_aip_xyzip.f90:7-75 -------------------------------------------
static inline void cross(double* a, const double* b, const double* c) {  // vec_a_bcrossc, mod_util.f90:692-704
  a[0] = b[1] * c[2] - b[2] * c[1];
  a[1] = b[2] * c[0] - b[0] * c[2];
  a[2] = b[0] * c[1] - b[1] * c[0];
}

void calc_aip_xyzip_uns(Mesh& m) {
  m.aip.alloc(3 * (long)m.nf, 0.0);
  m.rip.alloc(3 * (long)m.nf, 0.0);
  auto P = [&](int v, double* r) { r[0] = m.x[v - 1]; r[1] = m.y[v - 1]; r[2] = m.z[v - 1]; };
  for (int fg = 1; fg <= m.nf; ++fg) {
    int e, fl;
    get_idx(m.s2g(fg), e, fl);
    int s;
    for (s = 1; s <= m.nsec; ++s)
      if (e >= m.esec[2 * (s - 1)] && e <= m.esec[2 * (s - 1) + 1]) break;
    int t = m.etype[s - 1], lst[8], nl;
    face_vxlist(m.e2vx_of(e), t, fl, lst, nl);
    double r1[3], r2[3], r3[3], r4[3], dr1[3], dr2[3], areavec[3], subcntr[3], sumcntr[3][3], A[3];
    if (nl == 3 || nl == 4) {
      P(lst[0], r1); P(lst[1], r2); P(lst[2], r3);
      for (int i = 0; i < 3; ++i) { dr1[i] = r2[i] - r1[i]; dr2[i] = r3[i] - r1[i]; }
      cross(areavec, dr1, dr2);
      for (int i = 0; i < 3; ++i) { areavec[i] = 0.5 * areavec[i]; A[i] = areavec[i]; }
      for (int i = 0; i < 3; ++i) subcntr[i] = (r1[i] + r2[i] + r3[i]) / 3.0;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sumcntr[i][j] = subcntr[i] * areavec[j];
      if (nl == 4) {
        P(lst[3], r4);
        for (int i = 0; i < 3; ++i) { dr1[i] = r3[i] - r1[i]; dr2[i] = r4[i] - r1[i]; }
        cross(areavec, dr1, dr2);
        for (int i = 0; i < 3; ++i) { areavec[i] = 0.5 * areavec[i]; A[i] = A[i] + areavec[i]; }
        for (int i = 0; i < 3; ++i) subcntr[i] = (r1[i] + r3[i] + r4[i]) / 3.0;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sumcntr[i][j] = sumcntr[i][j] + subcntr[i] * areavec[j];
      }
    } else {
      throw std::runtime_error("Unsupported face in calc_xyzip_uns");
    }
    const double aa = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
    for (int i = 0; i < 3; ++i) {
      double c = 0.0;
      for (int j = 0; j < 3; ++j) c = c + sumcntr[i][j] * A[j];
      m.aip(3 * (long)fg - 3 + i + 1) = A[i];
      m.rip(3 * (long)fg - 3 + i + 1) = c / aa;
    }
  }
}

// ---- calc_vol_cv_centers_uns, calc_vol_cv_centers.f90:4-61 -------------------------------
void calc_vol_cv_centers_uns(Mesh& m) {
  const int ne = m.ne;
  m.xc.alloc(m.nelem, 0.0); m.yc.alloc(m.nelem, 0.0); m.zc.alloc(m.nelem, 0.0);
  m.vol.alloc(ne, 0.0);
  for (int s = 1; s <= m.nsec; ++s)
    for (int e = m.esec[2 * (s - 1)]; e <= m.esec[2 * (s - 1) + 1]; ++e) {
      if (e > ne) continue;
      int t = m.etype[s - 1], nv = element_nvx(t);
      double gc[3] = {0, 0, 0};
      for (int i = 1; i <= nv; ++i) {
        int v = m.e2vx_of(e)[i - 1];
        gc[0] = gc[0] + m.x[v - 1]; gc[1] = gc[1] + m.y[v - 1]; gc[2] = gc[2] + m.z[v - 1];
      }
      for (int i = 0; i < 3; ++i) gc[i] = gc[i] / nv;
      double sum_vol = 0.0, rc[3] = {0, 0, 0};
      for (int idx = m.ef2nb_idx(e); idx <= m.ef2nb_idx(e + 1) - 1; ++idx) {
        int gf = m.ef2nb2(idx);
        int gf_sgn = sgn(gf);
        gf = std::abs(gf);
        const double* cs = &m.rip(3 * (long)gf - 2);
        const double* a = &m.aip(3 * (long)gf - 2);
        int enb, lf;
        get_idx(m.ef2nb1(idx), enb, lf);
        if (lf == 0) { m.xc(enb) = cs[0]; m.yc(enb) = cs[1]; m.zc(enb) = cs[2]; }  // halo centre = face centroid
        double h[3] = {cs[0] - gc[0], cs[1] - gc[1], cs[2] - gc[2]};
        double sub_vol = ((gf_sgn * a[0]) * h[0] + (gf_sgn * a[1]) * h[1] + (gf_sgn * a[2]) * h[2]) / 3.0;
        sum_vol = sum_vol + sub_vol;
        for (int i = 0; i < 3; ++i) rc[i] = rc[i] + (0.25 * gc[i] + 0.75 * cs[i]) * sub_vol;
      }
      m.xc(e) = rc[0] / sum_vol; m.yc(e) = rc[1] / sum_vol; m.zc(e) = rc[2] / sum_vol;
      m.vol(e) = sum_vol;
    }
}

// ---- sorts, mod_util.f90:1602-1623 and :1683-1730 -----------------------------------------
void qsort_key(int* key, int* b, int i, int f) {  // key(*), b(*) 1-based, recursive, pivot = last
  if (i >= f) return;
  int n = i, p = f;
  while (n < p) {
    if (key[n - 1] > key[p - 1]) {
      std::swap(key[n - 1], key[p - 1]); std::swap(b[n - 1], b[p - 1]);
      p = p - 1;
      if (p > n) { std::swap(key[n - 1], key[p - 1]); std::swap(b[n - 1], b[p - 1]); }
    } else {
      n = n + 1;
    }
  }
  qsort_key(key, b, i, p - 1);
  qsort_key(key, b, p + 1, f);
}

void qsort_key_nRec(int* key, int* b, int n) {  // key(0:n-1), NOT stable
  const int max_levels = 1000;
  std::vector<int> beg(max_levels), end(max_levels);
  int i = 0;
  beg[0] = 0; end[0] = n;
  while (i >= 0) {
    int L = beg[i], R = end[i] - 1;
    if (L < R) {
      int piv = key[L], b0 = b[L];
      while (L < R) {
        while (key[R] >= piv && L < R) R = R - 1;
        if (L < R) { key[L] = key[R]; b[L] = b[R]; L = L + 1; }
        while (key[L] <= piv && L < R) L = L + 1;
        if (L < R) { key[R] = key[L]; b[R] = b[L]; R = R - 1; }
      }
      key[L] = piv; b[L] = b0;
      beg[i + 1] = L + 1; end[i + 1] = end[i]; end[i] = L;
      i = i + 1;
      if (i >= max_levels) throw std::runtime_error("qsort_key_nRec: stack overflow");
      if (end[i] - beg[i] > end[i - 1] - beg[i - 1]) { std::swap(beg[i], beg[i - 1]); std::swap(end[i], end[i - 1]); }
    } else {
      i = i - 1;
    }
  }
}

// ---- RCB: seed_gen_t, split_leaf, grow — mod_agglomeration.f90:331-561 ---------------------
namespace {
struct BB {
  int parent_group_id = 0, group_id = -1;
  double rmin[3], rmax[3];
  double group_vol = 0.0;
  std::vector<int> list;  // FIFO of cell ids (push = back, pop = front)
};
struct Node {
  BB* val = nullptr;
  Node *left = nullptr, *right = nullptr, *parent = nullptr;
  bool leaf() const { return !left && !right; }
};
struct SeedGen {
  int nleaf = 1, target_ncv = 1;
  double vol_ave = 0.0;
  const Mesh* m = nullptr;
  Node* root = nullptr;
  std::vector<Node*> pool;
  ~SeedGen() { for (Node* n : pool) { delete n->val; delete n; } }
  Node* mk(BB* bb, Node* parent) { Node* n = new Node; n->val = bb; n->parent = parent; pool.push_back(n); return n; }
  Node* begin() const { Node* n = root; while (n->left) n = n->left; return n; }
  Node* end() const { Node* n = root; while (n->right) n = n->right; return n; }
  bool iterate_next(Node*& it) const {  // in-order successor (the reference threads the tree for this)
    if (it == end()) return false;
    if (it->right) { it = it->right; while (it->left) it = it->left; }
    else { Node* c = it; Node* p = c->parent; while (p && p->right == c) { c = p; p = p->parent; } it = p; }
    return true;
  }
  bool goto_next_leaf(Node*& it) const {
    bool ok = false;
    while (iterate_next(it)) { ok = it->leaf(); if (ok) break; }
    return ok;
  }
  bool goto_1st_leaf(Node*& it) const {
    it = begin();
    if (!it->leaf()) return goto_next_leaf(it);
    return true;
  }
  bool split_leaf(Node* it, double threshold_density, int& nchild) {  // :380-498
    nchild = 0;
    if (!it->leaf()) return false;
    BB* bb = it->val;
    if (bb->group_vol / vol_ave < threshold_density) return false;
    if (bb->list.size() == 1) return false;
    double d = 0.0;
    int isplit = 0;
    for (int i = 0; i < 3; ++i)
      if (std::fabs(bb->rmin[i] - bb->rmax[i]) > d) { d = std::fabs(bb->rmin[i] - bb->rmax[i]); isplit = i + 1; }
    if (isplit == 0) throw std::runtime_error("rcb: degenerate bounding box");
    d = (bb->rmin[isplit - 1] + bb->rmax[isplit - 1]) / 2.0;
    BB* left = new BB; BB* right = new BB;
    const A1<double>& pos = (isplit == 1) ? m->xc : (isplit == 2 ? m->yc : m->zc);
    double lrmin[3] = {1e20, 1e20, 1e20}, lrmax[3] = {-1e20, -1e20, -1e20};
    double rrmin[3] = {1e20, 1e20, 1e20}, rrmax[3] = {-1e20, -1e20, -1e20};
    double vol_l = 0.0, vol_r = 0.0;  // phi == 1 (construct_seed_gen :352-353)
    for (int e : bb->list) {
      const double c[3] = {m->xc(e), m->yc(e), m->zc(e)};
      if (pos(e) < d) {
        left->list.push_back(e);
        vol_l = vol_l + m->vol(e) * 1.0;
        for (int i = 0; i < 3; ++i) { lrmin[i] = std::min(lrmin[i], c[i]); lrmax[i] = std::max(lrmax[i], c[i]); }
      } else {
        right->list.push_back(e);
        vol_r = vol_r + m->vol(e) * 1.0;
        for (int i = 0; i < 3; ++i) { rrmin[i] = std::min(rrmin[i], c[i]); rrmax[i] = std::max(rrmax[i], c[i]); }
      }
    }
    bb->list.clear(); bb->list.shrink_to_fit();
    left->group_vol = vol_l; right->group_vol = vol_r;
    for (int i = 0; i < 3; ++i) { left->rmin[i] = lrmin[i]; left->rmax[i] = lrmax[i]; right->rmin[i] = rrmin[i]; right->rmax[i] = rrmax[i]; }
    int pg = (bb->group_id == -1) ? bb->parent_group_id : bb->group_id;
    left->parent_group_id = pg; right->parent_group_id = pg;
    if (!left->list.empty() && !right->list.empty()) { nchild = 2; it->left = mk(left, it); it->right = mk(right, it); }
    else if (!left->list.empty()) { nchild = 1; it->left = mk(left, it); delete right; }
    else if (!right->list.empty()) { nchild = 1; it->right = mk(right, it); delete left; }
    else throw std::runtime_error("Error: This parent node does not have leaves!");
    return true;
  }
  void grow(int new_target_ncv, A1<int>& gf2g, int lvl) {  // :500-561
    int nleaf0 = nleaf;
    double threshold_density = 2.0;
    vol_ave = vol_ave * target_ncv / new_target_ncv;
    target_ncv = new_target_ncv;
    Node* it = nullptr;
    if (lvl > 1) {
      long guard = 0;
      while (nleaf < new_target_ncv) {
        if (goto_1st_leaf(it)) {
          for (;;) {
            int nchild;
            if (split_leaf(it, threshold_density, nchild)) {
              goto_next_leaf(it);
              nleaf = nleaf + nchild - 1;
              if (nleaf == new_target_ncv) break;
            }
            if (!goto_next_leaf(it)) break;
          }
        }
        if (nleaf == nleaf0) threshold_density = threshold_density * 0.75;
        nleaf0 = nleaf;
        if (++guard > 100000) throw std::runtime_error("rcb: cannot reach the requested number of leaves");
      }
      int n = 0;
      if (goto_1st_leaf(it))
        for (;;) {
          ++n;
          BB* leaf = it->val;
          if (leaf->group_id != -1) leaf->parent_group_id = leaf->group_id;
          leaf->group_id = n;
          gf2g(n) = leaf->parent_group_id;
          if (!goto_next_leaf(it)) break;
        }
    } else if (lvl == 1) {
      if (goto_1st_leaf(it))
        for (;;) {
          BB* leaf = it->val;
          for (int e : leaf->list) gf2g(e) = leaf->group_id;
          leaf->list.clear();
          if (!goto_next_leaf(it)) break;
        }
    }
  }
};
}  // namespace

// generate_seeds (mod_mg_lvl_uns.f90:95-118) with nl = (n_subdomains, 1), npmax = 2
// (cell_input.f90:131-137), then add_transformation_bt's block ordering (:873-903)
void rcb_partition(Mesh& m, int n_subdomains, bool stable_order) {
  const int ne = m.ne;
  SeedGen sg;
  sg.m = &m;
  BB* root_bb = new BB;
  sg.root = sg.mk(root_bb, nullptr);
  root_bb->rmin[0] = *std::min_element(m.xc.d.begin(), m.xc.d.begin() + ne);
  root_bb->rmin[1] = *std::min_element(m.yc.d.begin(), m.yc.d.begin() + ne);
  root_bb->rmin[2] = *std::min_element(m.zc.d.begin(), m.zc.d.begin() + ne);
  root_bb->rmax[0] = *std::max_element(m.xc.d.begin(), m.xc.d.begin() + ne);
  root_bb->rmax[1] = *std::max_element(m.yc.d.begin(), m.yc.d.begin() + ne);
  root_bb->rmax[2] = *std::max_element(m.zc.d.begin(), m.zc.d.begin() + ne);
  sg.vol_ave = 0.0;
  root_bb->list.reserve(ne);
  for (int e = 1; e <= ne; ++e) { root_bb->list.push_back(e); sg.vol_ave = sg.vol_ave + m.vol(e) * 1.0; }
  root_bb->group_vol = 2.1 * sg.vol_ave;
  root_bb->group_id = 1;
  A1<int> gf2g2(n_subdomains, 0);
  sg.grow(n_subdomains, gf2g2, 2);
  m.gf2g.alloc(ne, 0);
  sg.grow(ne, m.gf2g, 1);
  // add_transformation_bt, lvl = 1
  m.g2gf_p.alloc(ne, 0);
  for (int gf = 1; gf <= ne; ++gf) m.g2gf_p(gf) = gf;
  std::vector<int> key(m.gf2g.d);
  if (stable_order) {  // TIMING RUNS ONLY: same blocks, cells ascending inside a block (the reference sort is quadratic here)
    std::vector<int> perm(ne);
    for (int i = 0; i < ne; ++i) perm[i] = i + 1;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a - 1] < key[b - 1]; });
    for (int i = 0; i < ne; ++i) m.g2gf_p(i + 1) = perm[i];
    std::sort(key.begin(), key.end());
  } else {
    qsort_key_nRec(key.data(), m.g2gf_p.data(), ne);
  }
  int ng_tmp = key[ne - 1];
  m.g2gf_idx.alloc(ng_tmp + 1, 0);
  int g = 0;
  for (int idx = 1; idx <= ne; ++idx)
    if (g != key[idx - 1]) { g = key[idx - 1]; m.g2gf_idx(g) = idx; }
  m.g2gf_idx(ng_tmp + 1) = ne + 1;
  if (ng_tmp != n_subdomains) throw std::runtime_error("rcb: block count mismatch");
  m.n_subdomains = n_subdomains;
}

// ---- cell_input.f90:10-165 without the CGNS calls -------------------------------------------
void setup_mesh(Mesh& m, int n_subdomains) {
  // counts, cell_input.f90:58-78
  int cvs1 = 0, cvs2 = 0, cvs3 = 0;
  for (int s = 0; s < m.nsec; ++s) {
    int cnt = m.esec[2 * s + 1] - m.esec[2 * s] + 1;
    if (m.etype[s] >= 10 && m.etype[s] <= 20) { cvs1 += cnt; cvs2 += element_nface(m.etype[s]) * cnt; }
    else cvs3 += cnt;
  }
  cvs2 = (cvs2 + cvs3) / 2;
  m.ne = cvs1; m.nf = cvs2; m.nbf = cvs3;
  if (m.nelem != m.ne + m.nbf) throw std::runtime_error("setup_mesh: nelem != ne+nbf");
  // add_meshds, mod_mg_lvl_uns.f90:154-184 : order sections by start; 2-D sections are the BC interfaces
  {
    std::vector<int> order(m.nsec);
    for (int s = 0; s < m.nsec; ++s) order[s] = s;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return m.esec[2 * a] < m.esec[2 * b]; });
    std::vector<int> et(m.nsec), es(2 * m.nsec);
    std::vector<std::string> nm(m.nsec);
    for (int s = 0; s < m.nsec; ++s) { et[s] = m.etype[order[s]]; es[2 * s] = m.esec[2 * order[s]]; es[2 * s + 1] = m.esec[2 * order[s] + 1]; nm[s] = m.sectionName[order[s]]; }
    m.etype = et; m.esec = es; m.sectionName = nm;
  }
  m.bs_idx.alloc(m.nsec + 1, 0);
  m.bs_idx(1) = 1;
  for (int s = 1; s <= m.nsec; ++s) {
    if (m.esec[2 * (s - 1)] != m.bs_idx(s)) throw std::runtime_error("setup_mesh: sections are not contiguous");
    m.bs_idx(s + 1) = m.esec[2 * (s - 1) + 1] + 1;
  }
  m.intf2sec.clear();
  for (int s = 1; s <= m.nsec; ++s) {
    if (m.etype[s - 1] < 10) m.intf2sec.push_back(s);
    else if (m.esec[2 * (s - 1) + 1] > m.ne) throw std::runtime_error("setup_mesh: 3-D sections must come first");
  }
  m.nintf_c2b = (int)m.intf2sec.size();
  find_element_nb(m);
  calc_aip_xyzip_uns(m);
  calc_vol_cv_centers_uns(m);
  m.n_subdomains = 1;
  if (n_subdomains > 1) rcb_partition(m, n_subdomains);
}

}  // namespace orc
