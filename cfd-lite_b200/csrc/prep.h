// Host-side preparation of the device-resident mesh: unpack the reference's packed
// connectivity, choose the device numbering (multicolour-major, optionally
// Morton within colour), renumber faces by owner, build ELL slot arrays and the level
// schedules that reproduce the reference's sequential Gauss-Seidel order on the GPU.
//
// Distributed use (one process per GPU): every rank passes the same GLOBAL mesh plus a
// cell->rank map; prepare() keeps the rank's owned cells, adds the neighbouring cells owned by
// other ranks as ghost cells, and derives matching send/receive lists on every rank without
// any communication (both sides order an interface by (colour, base order) of the global mesh).
//
// Device index spaces:   cells  [0, N) owned, [N, Nc) ghosts (Nc = N + G)
//                        halos  [Nc, Nc + B) physical boundary faces of owned cells
//                        faces  [0, Fi) between two cells, [Fi, F) boundary (face Fi+j <-> halo j)
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
  int32_t N = 0, G = 0, Nc = 0, F = 0, B = 0, H = 0, Z = 0, K = 0, Np = 0, Fi = 0;
  int32_t gN = 0, gF = 0, gB = 0, gZ = 0;  // sizes of the global (reference-numbered) mesh
  int ncolors = 0;
  bool morton = false;
  int rank = 0, nranks = 1;
  std::vector<int32_t> row_ptr;    // gN+1, 0-based CSR offsets in ORIGINAL (global) cell order
  std::vector<int32_t> c2o;        // Nc: device cell (owned then ghost) -> original cell (0-based)
  std::vector<int32_t> o2c;        // gN: original cell -> device cell or -1
  std::vector<int32_t> h2o;        // B: device halo -> original halo (0-based offset from gN)
  std::vector<int32_t> f2o;        // F: device face -> original face (0-based)
  std::vector<int32_t> fown;       // F: 1 when this rank reports the face on download
  std::vector<int32_t> loc_order;  // N: owned device cells in base (natural | Morton) order — neighbours in space are neighbours here
  std::vector<int32_t> color_ptr;  // ncolors+1 over owned cells (device cells are sorted by colour)
  std::vector<int32_t> color_if;   // per colour: leading cells of the colour that touch another rank (interface cells)
  // per owned device cell c: the peers' ghost slots that mirror it, tgt_nbr/tgt_pos[tgt_ptr[c]..tgt_ptr[c+1])
  // (neighbour index into nbr_rank, position inside that neighbour's (colour) receive slice)
  std::vector<int32_t> tgt_ptr, tgt_nbr, tgt_pos;
  std::vector<int32_t> ell_nb;     // K*Np device index of neighbour (cell, ghost or halo), pad = self
  std::vector<int32_t> ell_fs;     // K*Np signed device face id +-(f+1); 0 = padding slot
  // two colours, one rank: neighbour ids as 16-bit offsets inside the other colour (K*Np; nb16_ok when every
  // offset fits).  Slot k of red row c points to cell nred + c + nb16, of black row c to (c - nred) + nb16;
  // boundary and padding slots carry the offset of the row's first cell neighbour (their pc coefficient is zero).
  std::vector<int16_t> nb16;
  bool nb16_ok = false;
  int32_t color_dist = -1;         // two colours, one rank: largest |position of a neighbour in its colour - own position| (-1: unknown)
  std::vector<uint8_t> nfc;        // N faces per cell
  std::vector<uint8_t> ftouch;     // N: bit k set when slot k's cell-cell face is numbered from this cell (its first toucher)
  int32_t touch_end = 0;           // cells [touch_end, N) have ftouch == 0 (two-colour mesh: the whole second colour)
  std::vector<int32_t> face_a, face_b;  // per device face: reference owner / neighbour (device idx; halo for boundary)
  std::vector<int32_t> halo_cell, halo_face, halo_bc;  // per halo: interior device cell, device face, bc index
  std::vector<uint8_t> halo_slot;  // per halo: ELL slot k in its interior cell
  std::vector<int32_t> bc_kind;
  std::vector<double> bc_uvw;
  Schedule natural;                // solve_gs order 1..ne            (single-rank only)
  Schedule blocks;                 // multi_subdomain_solver order    (single-rank, n_subdomains > 1)
  int n_subdomains = 1;
  // interfaces with other ranks; lists are grouped by neighbour, inside a neighbour by colour
  std::vector<int32_t> nbr_rank;   // nnbr
  std::vector<int32_t> send_ptr;   // nnbr*ncolors + 1 offsets into send_cells (neighbour-major, colour-minor)
  std::vector<int32_t> send_cells; // owned device cells whose values the neighbour needs
  std::vector<int32_t> recv_ptr;   // nnbr*ncolors + 1 offsets into the ghost range (ghost index = N + offset)
  int32_t ref_cell_owner = 0;      // rank owning original cell 1 (pref = phic(1), mod_uvwp.f90:129)
};

// returns 0 or a CFDL_ERR_* code (message via cfdl_last_error); cell2rank (1..nranks per
// original cell) may be NULL when nranks == 1
int prepare(Prep& p, int32_t ne, int32_t nf, int32_t nbf, const int32_t* ef2nb_idx,
            const int32_t* ef2nb_nb, const int32_t* ef2nb_fg, const int32_t* s2g, const int32_t* bs,
            const double* xc, const double* yc, const double* zc, int32_t nbc, const int32_t* bc_esec,
            const int32_t* bc_kind, const double* bc_uvw, int32_t n_subdomains,
            const int32_t* g2gf_p, const int32_t* g2gf_idx, int reorder_mode,
            const int32_t* cell2rank, int32_t rank, int32_t nranks);

// the part of prepare() after unpacking (plain int32 connectivity; see prep.cpp)
int prepare_core(Prep& p, const std::vector<int32_t>& o_nb, const std::vector<int32_t>& o_fg, const std::vector<int32_t>& halo_e,
                 const std::vector<int32_t>& halo_lf, const double* xc, const double* yc, const double* zc, int32_t nbc,
                 const int32_t* bc_esec, const int32_t* bc_kind, const double* bc_uvw, int32_t n_subdomains, const int32_t* g2gf_p,
                 const int32_t* g2gf_idx, int reorder_mode, const int32_t* cell2rank, int32_t rank, int32_t nranks);

}  // namespace cfdl
