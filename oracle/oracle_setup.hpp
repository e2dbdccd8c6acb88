// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by, or called from the product
// (libcfdl.so / the cfdl python binding).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it.
//
// CPU restatement (FP64, one thread) of the reference's mesh set-up, i.e. of everything that
// produces the hot path's input arrays:
//   find_element_nb            src/setup/mod_mg_lvl_uns.f90:283-433
//   calc_aip_xyzip_uns         src/setup/calc_aip_xyzip.f90:7-75
//   calc_vol_cv_centers_uns    src/setup/calc_vol_cv_centers.f90:4-61
//   RCB seed generator         src/setup/mod_agglomeration.f90:331-561
//   g2gf block ordering        src/setup/mod_mg_lvl_uns.f90:873-903, src/modules/mod_util.f90:1683-1730
//
// PARITY PINNED TO THE REFERENCE'S SOURCE TEXT (not to a compiled binary): the reference ships no tests, golden
// vectors or runnable case for this path (bundled test/box.cgns.tar.gz is missing) and no Fortran compiler exists here
// or on the GPU box, so find_element_nb, calc_aip_xyzip_uns, calc_vol_cv_centers_uns, generate_seeds (which subdomain a
// cell belongs to) and add_transformation_bt are executed from the reference's unmodified files by
// oracle/f90run/f90py.py and this restatement must give the same arrays bit for bit
// (tests/test_oracle_vs_reference_source.py).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <stdexcept>

namespace orc {

// 1-based array, Fortran style: a(i), i in [lo, lo+n)
template <class T>
struct A1 {
  std::vector<T> d;
  long lo = 1;
  A1() {}
  explicit A1(long n, T v = T(), long lo_ = 1) : d((size_t)n, v), lo(lo_) {}
  void alloc(long n, T v = T(), long lo_ = 1) { d.assign((size_t)n, v); lo = lo_; }
  T& operator()(long i) { return d[(size_t)(i - lo)]; }
  const T& operator()(long i) const { return d[(size_t)(i - lo)]; }
  long size() const { return (long)d.size(); }
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
};

// mod_util.f90:9,1428-1448 (lvl = 0 everywhere on the hot path => 5 face bits)
constexpr int num_face_bits = 5;
inline int index_t(int group_no, int side_no) { return (int)(((uint32_t)group_no << num_face_bits) | (uint32_t)side_no); }
inline void get_idx(int idx, int& group_no, int& side_no) {
  side_no = idx & ((1 << num_face_bits) - 1);
  group_no = (int)((uint32_t)idx >> num_face_bits);
}
// mod_util.f90:832-843 : sgn(0) = +1
inline int sgn(int x) { return x >= 0 ? 1 : -1; }

// element tables, mod_util.f90:55-85,166-169 (CGNS element type codes)
enum { TRI_3 = 5, QUAD_4 = 7, TETRA_4 = 10, PYRA_5 = 12, PENTA_6 = 14, HEXA_8 = 17 };
int element_nface(int t);
int element_nvx(int t);
// face_vxlist, mod_util.f90:1362-1426 : vertex list of local face f (1-based) of an element
void face_vxlist(const int* vxlist, int t, int f, int* lst, int& nl);

struct Mesh {
  // raw "CGNS" content (cell_input.f90:36-99)
  int nvx = 0, nelem = 0, nsec = 0, ne2vx_max = 0;
  std::vector<double> x, y, z;                 // 0-based storage of vertex v at [v-1]
  std::vector<int> e2vx;                       // (ne2vx_max, nelem) column-major
  std::vector<int> etype;                      // per section
  std::vector<int> esec;                       // (2, nsec)
  std::vector<std::string> sectionName;
  // geometry_t / meshds_t (mod_mg_lvl_uns.f90:46-58, mod_meshds_uns.f90:14-26)
  int ne = 0, nf = 0, nbf = 0;
  A1<int> ef2nb_idx;                           // (ne+1)
  A1<int> ef2nb1, ef2nb2;                      // ef2nb(:,1), ef2nb(:,2), each (2nf-nbf)
  A1<int> s2g;                                 // (nf)
  A1<int> bs;                                  // (ne+1 : ne+nbf)
  A1<int> bs_idx;                              // (nsec+1)
  A1<double> xc, yc, zc;                       // (ne+nbf)
  A1<double> aip, rip;                         // (3nf) AoS
  A1<double> vol;                              // (ne)
  // boundary interfaces = 2-D sections (mod_mg_lvl_uns.f90:170-184)
  int nintf_c2b = 0;
  std::vector<int> intf2sec;                   // 1-based section of interface i at [i-1]
  // subdomain maps (n_subdomains > 1)
  int n_subdomains = 1;
  A1<int> gf2g;                                // cell -> block (1..P)
  A1<int> g2gf_p;                              // cells sorted by block
  A1<int> g2gf_idx;                            // (P+1)

  const int* e2vx_of(int e) const { return &e2vx[(size_t)ne2vx_max * (size_t)(e - 1)]; }
  int section_of(int e) const {
    for (int s = 1; s <= nsec; ++s) if (e < bs_idx(s + 1)) return s;
    return nsec;
  }
};

// cell_input.f90:10-165 minus the CGNS reads: fills every geometry array of `m` from the raw
// content already stored in it.  Throws std::runtime_error where the reference `stop`s.
void setup_mesh(Mesh& m, int n_subdomains);

void find_element_nb(Mesh& m);
void calc_aip_xyzip_uns(Mesh& m);
void calc_vol_cv_centers_uns(Mesh& m);
void rcb_partition(Mesh& m, int n_subdomains, bool stable_order = false);  // generate_seeds + grow (+ block order)
void qsort_key_nRec(int* key, int* b, int n);           // mod_util.f90:1683-1730
void qsort_key(int* key, int* b, int i, int f);         // mod_util.f90:1602-1623 (1-based i,f)

}  // namespace orc
