// Synthetic unit-cube meshes in the CGNS conventions the reference's reader hands
// to its setup code (src/setup/cell_input.f90:36-99): one 3-D section followed by
// six named 2-D boundary sections, element->vertex lists in CGNS node order
// (HEXA_8=17, TETRA_4=10, QUAD_4=7, TRI_3=5; src/modules/mod_util.f90:13-37,55-85).
// Stands in for the CGNS file (no CGNS/HDF5 library in this image); pure host code.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "cfdl.h"

namespace {

// splitmix64: tiny, portable, deterministic (the mesh must be identical on every box).
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  uint64_t below(uint64_t n) { return next() % n; }
};

const char* kSectionNames[7] = {"interior", "bottom", "top", "west", "east", "south", "north"};

inline int64_t vid(int64_t i, int64_t j, int64_t k, int64_t n) {  // 1-based vertex id, i fastest
  return 1 + i + (n + 1) * (j + (n + 1) * k);
}

}  // namespace

extern "C" int cfdl_meshgen_sizes(int kind, int n, int64_t* nvx, int64_t* ne, int64_t* nbf,
                                  int* nsec, int* ne2vx_max) {
  if (n < 1 || (kind != CFDL_MESH_HEX && kind != CFDL_MESH_TET)) return CFDL_ERR_ARG;
  const int64_t n1 = n + 1, nn = n;
  *nvx = n1 * n1 * n1;
  if (kind == CFDL_MESH_HEX) {
    *ne = nn * nn * nn;
    *nbf = 6 * nn * nn;
    *ne2vx_max = 8;
  } else {
    *ne = 6 * nn * nn * nn;
    *nbf = 12 * nn * nn;
    *ne2vx_max = 4;
  }
  *nsec = 7;
  return CFDL_OK;
}

extern "C" int cfdl_meshgen_fill(int kind, int n, double jitter, int shuffle, uint64_t seed,
                                 double* x, double* y, double* z, int32_t* e2vx, int32_t* etype,
                                 int32_t* esec, char* names /* nsec*32, blank padded */) {
  int64_t nvx, ne, nbf;
  int nsec, w;
  int rc = cfdl_meshgen_sizes(kind, n, &nvx, &ne, &nbf, &nsec, &w);
  if (rc) return rc;
  if (ne + nbf >= (int64_t(1) << 26)) return CFDL_ERR_RANGE;  // reference packs (id<<5)|face in int32
  const double h = 1.0 / n;
  Rng rng(seed);
  // vertices; interior ones optionally jittered by U(-jitter*h, +jitter*h) per coordinate
  for (int k = 0; k <= n; ++k)
    for (int j = 0; j <= n; ++j)
      for (int i = 0; i <= n; ++i) {
        int64_t v = vid(i, j, k, n) - 1;
        double px = i * h, py = j * h, pz = k * h;
        if (i == n) px = 1.0;
        if (j == n) py = 1.0;
        if (k == n) pz = 1.0;
        if (jitter > 0.0 && i > 0 && i < n && j > 0 && j < n && k > 0 && k < n) {
          px += (2.0 * rng.uniform() - 1.0) * jitter * h;
          py += (2.0 * rng.uniform() - 1.0) * jitter * h;
          pz += (2.0 * rng.uniform() - 1.0) * jitter * h;
        }
        x[v] = px; y[v] = py; z[v] = pz;
      }
  std::memset(e2vx, 0, sizeof(int32_t) * (size_t)w * (size_t)(ne + nbf));
  // cell numbering permutation (new id -> generation-order id) when shuffle is requested
  std::vector<int64_t> slot((size_t)ne);
  for (int64_t e = 0; e < ne; ++e) slot[e] = e;
  if (shuffle) {
    Rng r2(seed ^ 0xC0FFEEull);
    for (int64_t e = ne - 1; e > 0; --e) std::swap(slot[e], slot[(int64_t)r2.below((uint64_t)e + 1)]);
  }
  // slot[] maps generation order -> stored position
  if (kind == CFDL_MESH_HEX) {
    int64_t g = 0;
    for (int k = 0; k < n; ++k)
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i, ++g) {
          int32_t* c = e2vx + 8 * slot[g];
          c[0] = (int32_t)vid(i, j, k, n);         c[1] = (int32_t)vid(i + 1, j, k, n);
          c[2] = (int32_t)vid(i + 1, j + 1, k, n); c[3] = (int32_t)vid(i, j + 1, k, n);
          c[4] = (int32_t)vid(i, j, k + 1, n);     c[5] = (int32_t)vid(i + 1, j, k + 1, n);
          c[6] = (int32_t)vid(i + 1, j + 1, k + 1, n); c[7] = (int32_t)vid(i, j + 1, k + 1, n);
        }
  } else {
    // Kuhn split: six tets per cube around the (0,0,0)-(1,1,1) diagonal, one per axis permutation.
    static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    static const int sign[6] = {+1, -1, -1, +1, +1, -1};
    int64_t g = 0;
    for (int k = 0; k < n; ++k)
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i)
          for (int t = 0; t < 6; ++t, ++g) {
            int p[3] = {i, j, k};
            int32_t v[4];
            v[0] = (int32_t)vid(p[0], p[1], p[2], n);
            for (int s = 0; s < 3; ++s) { p[perm[t][s]] += 1; v[s + 1] = (int32_t)vid(p[0], p[1], p[2], n); }
            if (sign[t] < 0) std::swap(v[1], v[2]);  // keep (r2-r1)x(r3-r1).(r4-r1) > 0
            int32_t* c = e2vx + 4 * slot[g];
            c[0] = v[0]; c[1] = v[1]; c[2] = v[2]; c[3] = v[3];
          }
  }
  // boundary sections: bottom(z=0) top(z=1) west(x=0) east(x=1) south(y=0) north(y=1)
  etype[0] = (kind == CFDL_MESH_HEX) ? 17 : 10;
  esec[0] = 1; esec[1] = (int32_t)ne;
  std::memset(names, ' ', 32 * 7);
  for (int s = 0; s < 7; ++s) std::memcpy(names + 32 * s, kSectionNames[s], std::strlen(kSectionNames[s]));
  int64_t b = ne;  // next free element slot (0-based)
  const int per = (kind == CFDL_MESH_HEX) ? 1 : 2;
  for (int s = 0; s < 6; ++s) {
    etype[s + 1] = (kind == CFDL_MESH_HEX) ? 7 : 5;
    esec[2 * (s + 1)] = (int32_t)(b + 1);
    const int axis = (s < 2) ? 2 : (s < 4 ? 0 : 1);  // fixed coordinate
    const int fixed = (s % 2 == 0) ? 0 : n;
    const int a1 = (axis == 0) ? 1 : 0, a2 = (axis == 2) ? 1 : 2;  // the two in-plane axes, ascending
    for (int q = 0; q < n; ++q)
      for (int p = 0; p < n; ++p) {
        int c[3];
        auto V = [&](int dp, int dq) {
          c[axis] = fixed; c[a1] = p + dp; c[a2] = q + dq;
          return (int32_t)vid(c[0], c[1], c[2], n);
        };
        int32_t v00 = V(0, 0), v10 = V(1, 0), v11 = V(1, 1), v01 = V(0, 1);
        if (per == 1) {
          int32_t* d = e2vx + (size_t)w * b++;
          d[0] = v00; d[1] = v10; d[2] = v11; d[3] = v01;
        } else {  // Kuhn faces are cut along the (0,0)-(1,1) diagonal of every cube face
          int32_t* d = e2vx + (size_t)w * b++;
          d[0] = v00; d[1] = v10; d[2] = v11;
          d = e2vx + (size_t)w * b++;
          d[0] = v00; d[1] = v11; d[2] = v01;
        }
      }
    esec[2 * (s + 1) + 1] = (int32_t)b;
  }
  return (b == ne + nbf) ? CFDL_OK : CFDL_ERR_INTERNAL;
}

// ---- raw mesh file -------------------------------------------------------------------------------
// What cell_input.f90:36-99 obtains from the CGNS library, in one little-endian binary file, for
// drivers built without CGNS/HDF5 (the reference needs a modified cgnslib 3.2.1): the Fortran side
// reads it with stream access (cfd-lite_b200/fortran/mod_rawmesh.f90) instead of cgns_db_open /
// cg_section_read_f / cgns_element_info / cgns_cell_data.  Layout:
//   char[8] "CFDLRAW1" | int64 nvx | int64 nelem | int32 nsec | int32 ne2vx_max |
//   nsec x { char[32] name (blank padded) | int32 etype | int32 first | int32 last } |
//   double x[nvx] | double y[nvx] | double z[nvx] | int32 e2vx[ne2vx_max * nelem]
#include <cstdio>
namespace {
const char kRawMagic[8] = {'C', 'F', 'D', 'L', 'R', 'A', 'W', '1'};
struct File {
  std::FILE* f;
  explicit File(std::FILE* p) : f(p) {}
  ~File() { if (f) std::fclose(f); }
};
bool put(std::FILE* f, const void* p, size_t n) { return n == 0 || std::fwrite(p, 1, n, f) == n; }
bool get(std::FILE* f, void* p, size_t n) { return n == 0 || std::fread(p, 1, n, f) == n; }
int raw_header(std::FILE* f, int64_t* nvx, int64_t* nelem, int32_t* nsec, int32_t* w) {
  char magic[8];
  if (!get(f, magic, 8) || std::memcmp(magic, kRawMagic, 8) != 0) return CFDL_ERR_ARG;
  if (!get(f, nvx, 8) || !get(f, nelem, 8) || !get(f, nsec, 4) || !get(f, w, 4)) return CFDL_ERR_ARG;
  if (*nvx < 0 || *nelem < 0 || *nsec < 0 || *nsec > 4096 || *w < 0 || *w > 64) return CFDL_ERR_ARG;
  return CFDL_OK;
}
}  // namespace

extern "C" int cfdl_rawmesh_write(const char* path, int64_t nvx, const double* x, const double* y, const double* z, int nsec,
                                  const int32_t* etype, const int32_t* esec, const char* names, int ne2vx_max, const int32_t* e2vx,
                                  int64_t nelem) {
  if (!path || !x || !y || !z || !etype || !esec || !names || !e2vx || nvx < 0 || nelem < 0 || nsec < 0 || ne2vx_max < 0) return CFDL_ERR_ARG;
  File fh(std::fopen(path, "wb"));
  if (!fh.f) return CFDL_ERR_ARG;
  const int32_t ns = nsec, w = ne2vx_max;
  bool ok = put(fh.f, kRawMagic, 8) && put(fh.f, &nvx, 8) && put(fh.f, &nelem, 8) && put(fh.f, &ns, 4) && put(fh.f, &w, 4);
  for (int s = 0; s < nsec && ok; ++s) {
    char nm[32];
    std::memset(nm, ' ', 32);
    for (int i = 0; i < 32 && names[32 * s + i] != 0; ++i) nm[i] = names[32 * s + i];
    ok = put(fh.f, nm, 32) && put(fh.f, &etype[s], 4) && put(fh.f, &esec[2 * s], 4) && put(fh.f, &esec[2 * s + 1], 4);
  }
  ok = ok && put(fh.f, x, 8 * (size_t)nvx) && put(fh.f, y, 8 * (size_t)nvx) && put(fh.f, z, 8 * (size_t)nvx) &&
       put(fh.f, e2vx, 4 * (size_t)ne2vx_max * (size_t)nelem);
  return ok ? CFDL_OK : CFDL_ERR_INTERNAL;
}

extern "C" int cfdl_rawmesh_sizes(const char* path, int64_t* nvx, int64_t* nelem, int* nsec, int* ne2vx_max) {
  if (!path || !nvx || !nelem || !nsec || !ne2vx_max) return CFDL_ERR_ARG;
  File fh(std::fopen(path, "rb"));
  if (!fh.f) return CFDL_ERR_ARG;
  int32_t ns = 0, w = 0;
  int rc = raw_header(fh.f, nvx, nelem, &ns, &w);
  *nsec = ns; *ne2vx_max = w;
  return rc;
}

extern "C" int cfdl_rawmesh_read(const char* path, double* x, double* y, double* z, int32_t* etype, int32_t* esec, char* names,
                                 int32_t* e2vx) {
  if (!path || !x || !y || !z || !etype || !esec || !names || !e2vx) return CFDL_ERR_ARG;
  File fh(std::fopen(path, "rb"));
  if (!fh.f) return CFDL_ERR_ARG;
  int64_t nvx = 0, nelem = 0;
  int32_t ns = 0, w = 0;
  int rc = raw_header(fh.f, &nvx, &nelem, &ns, &w);
  if (rc) return rc;
  bool ok = true;
  for (int s = 0; s < ns && ok; ++s)
    ok = get(fh.f, names + 32 * s, 32) && get(fh.f, &etype[s], 4) && get(fh.f, &esec[2 * s], 4) && get(fh.f, &esec[2 * s + 1], 4);
  ok = ok && get(fh.f, x, 8 * (size_t)nvx) && get(fh.f, y, 8 * (size_t)nvx) && get(fh.f, z, 8 * (size_t)nvx) &&
       get(fh.f, e2vx, 4 * (size_t)w * (size_t)nelem);
  return ok ? CFDL_OK : CFDL_ERR_ARG;
}
