// cfdl_create_structured_hex: a solver handle for the synthetic n^3 lid-driven cavity without
// going through the reference's packed (cell<<5|face) int32 arrays, whose 2^26 limit
//  stops at 406^3.  The mesh is the one cfdl_meshgen_fill(HEX, n, 0, 0) +
// cfdl_mesh_build would produce — same cell, halo and local-face numbering, same geometry bits
// (geom_formulas.h) — but connectivity and geometry are generated analytically per rank, so
// the 512^3 (134 M cell) target of BASELINE.json fits.  Only the global FACE numbering differs
// from the reference's (x-normal faces, then y, then z), so face fields (mip) of such a handle
// are in that order on the host side.
//
// Cell (i,j,k) has 0-based id i + n (j + n k); local faces follow CGNS HEXA_8
// (z-, y-, x+, y+, x-, z+); boundary halos are numbered bottom, top, west, east, south, north
// with the generator's in-plane order.
#include <cstdint>
#include <vector>
#include "cfdl_common.h"
#include "geom_formulas.h"
#include "state.h"

namespace cfdl {
namespace {

// n x n x nz cells of edge 1/n: the unit cube for nz = n (the cavity of BASELINE.json), a cavity of depth nz/n otherwise
// (the weak-scaling series of bench.py stacks one 128^3 block per GPU along z)
struct Hex {
  int32_t n, nz;
  int64_t n2, n3, Fx, Fy, Fz, nB;
  int64_t hoff[7];  // first halo of the sections bottom, top, west, east, south, north
  explicit Hex(int32_t n_, int32_t nz_ = 0) : n(n_), nz(nz_ > 0 ? nz_ : n_) {
    n2 = (int64_t)n * n; n3 = n2 * nz; Fx = (int64_t)(n + 1) * n * nz; Fy = Fx; Fz = n2 * (nz + 1);
    const int64_t side = (int64_t)n * nz;
    hoff[0] = 0; hoff[1] = n2; hoff[2] = 2 * n2; hoff[3] = hoff[2] + side; hoff[4] = hoff[3] + side; hoff[5] = hoff[4] + side; hoff[6] = hoff[5] + side;
    nB = hoff[6];
  }
  double vx(int i) const { return i == n ? 1.0 : i * (1.0 / n); }  // x, y and (for i <= n) z vertex planes: the cube's bits
  // 0-based global face ids; i (resp. j, k) runs over n+1 planes
  int32_t fx(int i, int j, int k) const { return (int32_t)(i + (int64_t)(n + 1) * (j + (int64_t)n * k)); }
  int32_t fy(int i, int j, int k) const { return (int32_t)(Fx + i + (int64_t)n * (j + (int64_t)(n + 1) * k)); }
  int32_t fz(int i, int j, int k) const { return (int32_t)(Fx + Fy + i + (int64_t)n * (j + (int64_t)n * k)); }
  // halo offsets (0-based from gN): section s, in-plane (p, q)
  int32_t halo(int s, int p, int q) const { return (int32_t)(hoff[s] + (int64_t)q * n + p); }

  // vertices of local face lf (0..5) of cell (i,j,k), CGNS order of the reference's faces_hexa8
  void face_vertices(int i, int j, int k, int lf, double (*r)[3]) const {
    static const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    static const int fv[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {4, 5, 6, 7}};
    for (int q = 0; q < 4; ++q) {
      const int* c = corner[fv[lf][q]];
      r[q][0] = vx(i + c[0]); r[q][1] = vx(j + c[1]); r[q][2] = vx(k + c[2]);
    }
  }
  // area vector and centroid of global face f, computed from its owner's vertex order
  void face_geom(int32_t f, double* aip, double* rip) const {
    int i, j, k, lf;
    int64_t g = f;
    if (g < Fx) {
      i = (int)(g % (n + 1)); g /= (n + 1); j = (int)(g % n); k = (int)(g / n);
      if (i == 0) lf = 4; else { i -= 1; lf = 2; }
    } else if (g < Fx + Fy) {
      g -= Fx;
      i = (int)(g % n); g /= n; j = (int)(g % (n + 1)); k = (int)(g / (n + 1));
      if (j == 0) lf = 1; else { j -= 1; lf = 3; }
    } else {
      g -= Fx + Fy;
      i = (int)(g % n); g /= n; j = (int)(g % n); k = (int)(g / n);
      if (k == 0) lf = 0; else { k -= 1; lf = 5; }
    }
    double r[4][3];
    face_vertices(i, j, k, lf, r);
    face_area_centroid(r, 4, aip, rip);
  }
  // the six slots of cell (i,j,k): neighbour (cell id or gN + halo) and signed 1-based face id
  void slots(int i, int j, int k, int32_t* nb, int32_t* fg) const {
    const int32_t e = (int32_t)(i + (int64_t)n * (j + (int64_t)n * k));
    const int32_t gN = (int32_t)n3;
    nb[0] = k > 0 ? e - (int32_t)n2 : gN + halo(0, i, j);     fg[0] = k > 0 ? -(fz(i, j, k) + 1) : fz(i, j, k) + 1;
    nb[1] = j > 0 ? e - n : gN + halo(4, i, k);               fg[1] = j > 0 ? -(fy(i, j, k) + 1) : fy(i, j, k) + 1;
    nb[2] = i < n - 1 ? e + 1 : gN + halo(3, j, k);           fg[2] = fx(i + 1, j, k) + 1;
    nb[3] = j < n - 1 ? e + n : gN + halo(5, i, k);           fg[3] = fy(i, j + 1, k) + 1;
    nb[4] = i > 0 ? e - 1 : gN + halo(2, j, k);               fg[4] = i > 0 ? -(fx(i, j, k) + 1) : fx(i, j, k) + 1;
    nb[5] = k < nz - 1 ? e + (int32_t)n2 : gN + halo(1, i, j); fg[5] = fz(i, j, k + 1) + 1;
  }
  // centroid and volume of cell e with the reference's pyramid sums over its six faces
  void cell_geom(int32_t e, double* ctr, double* vol) const {
    const int i = e % n, j = (int)((e / n) % n), k = (int)(e / n2);
    int32_t nb[6], fg[6];
    slots(i, j, k, nb, fg);
    CellAccumulator acc;
    static const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    acc.gc[0] = acc.gc[1] = acc.gc[2] = 0.0;
    for (int q = 0; q < 8; ++q) {
      acc.gc[0] = acc.gc[0] + vx(i + corner[q][0]); acc.gc[1] = acc.gc[1] + vx(j + corner[q][1]); acc.gc[2] = acc.gc[2] + vx(k + corner[q][2]);
    }
    for (int q = 0; q < 3; ++q) acc.gc[q] = acc.gc[q] / 8;
    for (int s = 0; s < 6; ++s) {
      double a[3], r[3];
      face_geom(std::abs(fg[s]) - 1, a, r);
      acc.add_face(fg[s] > 0 ? 1 : -1, a, r);
    }
    acc.finish(ctr, vol);
  }
  // boundary face of halo h (0-based offset): the face of its interior cell
  void halo_owner(int32_t h, int* i, int* j, int* k, int* lf) const {
    int s = 0;
    while (h >= hoff[s + 1]) ++s;
    const int64_t l = h - hoff[s];
    const int p = (int)(l % n), q = (int)(l / n);
    switch (s) {
      case 0: *i = p; *j = q; *k = 0;      *lf = 0; break;
      case 1: *i = p; *j = q; *k = nz - 1; *lf = 5; break;
      case 2: *i = 0;     *j = p; *k = q; *lf = 4; break;
      case 3: *i = n - 1; *j = p; *k = q; *lf = 2; break;
      case 4: *i = p; *j = 0;     *k = q; *lf = 1; break;
      default: *i = p; *j = n - 1; *k = q; *lf = 3; break;
    }
  }
};

// recursive coordinate bisection of the index box, cutting x, y, z in turn at floor(m/2) — what
// the reference's bounding-box midpoint test (mod_agglomeration.f90) gives on a uniform cube
void bisect(std::vector<int32_t>& c2r, const Hex& hx, int lo[3], int hi[3], int axis, int first, int count) {
  if (count == 1) {
    for (int k = lo[2]; k < hi[2]; ++k)
      for (int j = lo[1]; j < hi[1]; ++j)
        for (int i = lo[0]; i < hi[0]; ++i) c2r[(size_t)(i + (int64_t)hx.n * (j + (int64_t)hx.n * k))] = first + 1;
    return;
  }
  const int mid = lo[axis] + (hi[axis] - lo[axis]) / 2;
  int l2[3] = {lo[0], lo[1], lo[2]}, h2[3] = {hi[0], hi[1], hi[2]};
  h2[axis] = mid;
  bisect(c2r, hx, l2, h2, (axis + 1) % 3, first, count / 2);
  h2[axis] = hi[axis]; l2[axis] = mid;
  bisect(c2r, hx, l2, h2, (axis + 1) % 3, first + count / 2, count / 2);
}

}  // namespace
}  // namespace cfdl

using namespace cfdl;

// slabs: 0 = recursive bisection (x, y, z in turn), 1 = nranks slabs along z, the slowest index of the numbering:
// every rank's interface cells are then the first and last planes of its own range — contiguous in the natural order
static int create_structured(cfdl_handle* out, int32_t n, int32_t nz, double rho, double mu, int32_t rank, int32_t nranks, int32_t device, int slabs) {
  if (!out) return fail(CFDL_ERR_ARG, "cfdl_create_structured_hex: out is NULL");
  *out = nullptr;
  if (n < 2 || n > 700) return fail(CFDL_ERR_RANGE, "cfdl_create_structured_hex: n=%d outside 2..700 (int32 slot index)", n);
  if (nz <= 0) nz = n;
  if (nz < 2 || (int64_t)n * n * nz > 343000000ll) return fail(CFDL_ERR_RANGE, "cfdl_create_structured_hex: %d x %d x %d cells exceed the int32 slot index", n, n, nz);
  if (nz != n && !slabs) return fail(CFDL_ERR_ARG, "cfdl_create_structured_hex: a depth other than n needs the slab partition");
  if (nranks < 1 || (!slabs && (nranks & (nranks - 1))) || rank < 0 || rank >= nranks)
    return fail(CFDL_ERR_ARG, "cfdl_create_structured_hex: rank %d of %d (power-of-two rank counts only)", rank, nranks);
  int ndev = cfdl_device_count();
  if (ndev < 1) return fail(CFDL_ERR_CUDA, "cfdl_create_structured_hex: no CUDA device is usable (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(CFDL_ERR_ARG, "cfdl_create_structured_hex: device %d of %d", device, ndev);
  if ((slabs ? nz : n) < nranks) return fail(CFDL_ERR_ARG, "cfdl_create_structured_hex: %d ranks for n=%d", nranks, n);
  const Hex hx(n, nz);
  cfdl_handle_s* h = new (std::nothrow) cfdl_handle_s;
  if (!h) return fail(CFDL_ERR_INTERNAL, "out of host memory");
  h->device = device;
  Prep& p = h->prep;
  const int32_t gN = (int32_t)hx.n3;
  p.gN = gN; p.gF = (int32_t)(hx.Fx + hx.Fy + hx.Fz); p.gB = (int32_t)hx.nB; p.gZ = (int32_t)(6 * hx.n3);
  p.K = 6; p.rank = rank; p.nranks = nranks; p.n_subdomains = 1;
  p.row_ptr.resize((size_t)gN + 1);
  for (int64_t e = 0; e <= gN; ++e) p.row_ptr[(size_t)e] = (int32_t)(6 * e);
  std::vector<int32_t> o_nb((size_t)p.gZ), o_fg((size_t)p.gZ), halo_e((size_t)p.gB), halo_lf((size_t)p.gB), c2r;
  {
    int64_t e = 0;
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i, ++e) hx.slots(i, j, k, &o_nb[(size_t)(6 * e)], &o_fg[(size_t)(6 * e)]);
  }
  for (int32_t hh = 0; hh < p.gB; ++hh) {
    int i, j, k, lf;
    hx.halo_owner(hh, &i, &j, &k, &lf);
    halo_e[hh] = (int32_t)(i + (int64_t)n * (j + (int64_t)n * k));
    halo_lf[hh] = lf + 1;
  }
  if (nranks > 1) {
    c2r.resize((size_t)gN);
    int lo[3] = {0, 0, 0}, hi[3] = {n, n, n};
    if (slabs) {
      for (int r = 0; r < nranks; ++r) {
        const int k0 = (int)((int64_t)nz * r / nranks), k1 = (int)((int64_t)nz * (r + 1) / nranks);
        std::fill(c2r.begin() + (size_t)k0 * hx.n2, c2r.begin() + (size_t)k1 * hx.n2, r + 1);
      }
    } else {
      bisect(c2r, hx, lo, hi, 0, 0, nranks);
    }
  }
  // the cavity's boundary conditions: six wall sections, the top one moving with (1,0,0)
  int32_t bc_esec[12], bc_kind[6];
  double bc_uvw[18] = {0};
  for (int s = 0; s < 6; ++s) {
    bc_esec[2 * s] = gN + 1 + (int32_t)hx.hoff[s];
    bc_esec[2 * s + 1] = gN + (int32_t)hx.hoff[s + 1];
    bc_kind[s] = (s == 1) ? CFDL_BC_LID : CFDL_BC_WALL;
  }
  bc_uvw[3] = 1.0;
  int rc = prepare_core(p, o_nb, o_fg, halo_e, halo_lf, nullptr, nullptr, nullptr, 6, bc_esec, bc_kind, bc_uvw, 1, nullptr, nullptr,
                        /*natural base order*/ 0, nranks > 1 ? c2r.data() : nullptr, rank, nranks);
  if (rc) { delete h; return rc; }
  std::vector<int32_t>().swap(o_nb); std::vector<int32_t>().swap(o_fg); std::vector<int32_t>().swap(c2r);
  GeomSource G;
  G.cell_xyz = [hx, gN](int32_t g, double* o) {
    if (g < gN) { double v; hx.cell_geom(g, o, &v); return; }
    int i, j, k, lf;
    hx.halo_owner(g - gN, &i, &j, &k, &lf);
    double r[4][3], a[3];
    hx.face_vertices(i, j, k, lf, r);
    face_area_centroid(r, 4, a, o);
  };
  G.vol = [hx](int32_t g) { double c[3], v; hx.cell_geom(g, c, &v); return v; };
  G.rho = [rho](int32_t) { return rho; };
  G.mu = [mu](int32_t) { return mu; };
  G.face = [hx](int32_t f, double* a, double* r) { hx.face_geom(f, a, r); };
  return create_from_prep(h, G, out);
}

extern "C" int cfdl_create_structured_hex(cfdl_handle* out, int32_t n, double rho, double mu, int32_t rank, int32_t nranks, int32_t device) {
  return create_structured(out, n, n, rho, mu, rank, nranks, device, 0);
}
extern "C" int cfdl_create_structured_hex_slabs(cfdl_handle* out, int32_t n, int32_t nz, double rho, double mu, int32_t rank, int32_t nranks,
                                                int32_t device) {
  return create_structured(out, n, nz, rho, mu, rank, nranks, device, 1);
}

// The arrays cfdl_create_structured_hex works from, for inspection and tests (no GPU needed).
// Any output may be NULL.  nb/fg: 6 n^3 slots (0-based neighbour cell or n^3 + halo; signed 1-based
// face id in this file's face numbering); xc,yc,zc: n^3 + 6 n^2; vol: n^3; aip,rip: 3 per face;
// cell2rank: n^3 entries in 1..nranks.
extern "C" int cfdl_structured_hex_arrays(int32_t n, int32_t nranks, int32_t* nb, int32_t* fg, double* xc, double* yc, double* zc,
                                          double* vol, double* aip, double* rip, int32_t* cell2rank) {
  if (n < 2 || n > 700) return fail(CFDL_ERR_RANGE, "cfdl_structured_hex_arrays: n=%d outside 2..700", n);
  if (nranks < 1 || (nranks & (nranks - 1)) || n < nranks) return fail(CFDL_ERR_ARG, "cfdl_structured_hex_arrays: %d ranks", nranks);
  const Hex hx(n);
  const int32_t gN = (int32_t)hx.n3;
  int64_t e = 0;
  for (int k = 0; k < n; ++k)
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i, ++e) {
        int32_t a[6], b[6];
        hx.slots(i, j, k, a, b);
        for (int s = 0; s < 6; ++s) { if (nb) nb[6 * e + s] = a[s]; if (fg) fg[6 * e + s] = b[s]; }
        double c[3], v;
        if (xc || yc || zc || vol) hx.cell_geom((int32_t)e, c, &v);
        if (xc) xc[e] = c[0];
        if (yc) yc[e] = c[1];
        if (zc) zc[e] = c[2];
        if (vol) vol[e] = v;
      }
  for (int32_t hh = 0; hh < 6 * hx.n2; ++hh) {
    int i, j, k, lf;
    hx.halo_owner(hh, &i, &j, &k, &lf);
    double r[4][3], a[3], c[3];
    hx.face_vertices(i, j, k, lf, r);
    face_area_centroid(r, 4, a, c);
    if (xc) xc[gN + hh] = c[0];
    if (yc) yc[gN + hh] = c[1];
    if (zc) zc[gN + hh] = c[2];
  }
  if (aip || rip)
    for (int64_t f = 0; f < 3 * hx.Fx; ++f) {
      double a[3], c[3];
      hx.face_geom((int32_t)f, a, c);
      for (int q = 0; q < 3; ++q) { if (aip) aip[3 * f + q] = a[q]; if (rip) rip[3 * f + q] = c[q]; }
    }
  if (cell2rank) {
    std::vector<int32_t> c2r((size_t)gN, 1);
    int lo[3] = {0, 0, 0}, hi[3] = {n, n, n};
    bisect(c2r, hx, lo, hi, 0, 0, nranks);
    std::copy(c2r.begin(), c2r.end(), cell2rank);
  }
  return CFDL_OK;
}
