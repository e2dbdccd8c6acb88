// Host-side preparation of the device-resident mesh: unpack the reference's packed
// connectivity, choose the device numbering (multicolour-major, optionally
// Morton within colour), renumber faces by owner, build ELL slot arrays and the level
// schedules that reproduce the reference's sequential Gauss-Seidel order on the GPU.
#pragma once
#include <cstdint>
#include <vector>

namespace cfdl {

// A sweep schedule: cells sorted by dependency level of a given sequential sweep order.
// Sweep-space index s in [0,N); values live in phi_s[0..H) (cells then halos) and a lagged
// copy phi_s[H..H+N) that cross-block neighbours read (block-Jacobi coupling of
// multi_subdomain_solver, src/modules/mod_solver.f90:157-169).
struct Schedule {
  int nlevels = 0;
  int nblocks = 1;
  std::vector<int32_t> lvl_ptr;    // nlevels+1
  std::vector<int32_t> s2c;        // sweep index -> device cell index
  std::vector<int32_t> nbs;        // K*Np, neighbour position in phi_s (see above)
  std::vector<int32_t> bpos;       // sweep index -> position in block order (0..N)
  std::vector<int32_t> blk_ptr;    // nblocks+1 offsets in block order
  std::vector<int32_t> lag_src;    // sweep indices whose value is copied to the lag area
};

struct Prep {
  int32_t N = 0, F = 0, B = 0, H = 0, Z = 0, K = 0, Np = 0, Fi = 0;
  int ncolors = 0;
  bool morton = false;
  std::vector<int32_t> row_ptr;    // N+1, 0-based CSR offsets in ORIGINAL cell order
  std::vector<int32_t> c2o, o2c;   // device cell <-> original cell (0-based)
  std::vector<int32_t> f2o, o2f;   // device face <-> original face (0-based)
  std::vector<int32_t> color_ptr;  // ncolors+1 (device cells are sorted by colour)
  std::vector<int32_t> ell_nb;     // K*Np device index of neighbour (cell, or N+halo), pad = self
  std::vector<int32_t> ell_fs;     // K*Np signed device face id +-(f+1); 0 = padding slot
  std::vector<uint8_t> nfc;        // N faces per cell
  std::vector<int32_t> face_a, face_b;  // per device face: reference owner / neighbour (device idx; halo for boundary)
  std::vector<int32_t> halo_cell, halo_face, halo_bc;  // per halo: interior device cell, device face, bc index
  std::vector<uint8_t> halo_slot;  // per halo: ELL slot k in its interior cell
  std::vector<int32_t> bc_kind;
  std::vector<double> bc_uvw;
  Schedule natural;                // solve_gs order 1..ne
  Schedule blocks;                 // multi_subdomain_solver order (empty when n_subdomains == 1)
  int n_subdomains = 1;
};

// returns 0 or a CFDL_ERR_* code (message via cfdl_last_error)
int prepare(Prep& p, int32_t ne, int32_t nf, int32_t nbf, const int32_t* ef2nb_idx,
            const int32_t* ef2nb_nb, const int32_t* ef2nb_fg, const int32_t* s2g, const int32_t* bs,
            const double* xc, const double* yc, const double* zc, int32_t nbc, const int32_t* bc_esec,
            const int32_t* bc_kind, const double* bc_uvw, int32_t n_subdomains,
            const int32_t* g2gf_p, const int32_t* g2gf_idx, int reorder_mode);

}  // namespace cfdl
