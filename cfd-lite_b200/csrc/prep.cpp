// Host-side mesh preparation (see prep.h).  Input = geometry_t/meshds_t arrays exactly as the
// reference's cell_input builds them (src/setup/mod_mg_lvl_uns.f90:283-433).
#include "prep.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include "cfdl_common.h"

namespace cfdl {
namespace {

inline uint64_t spread21(uint64_t v) {  // interleave helper for 3-D Morton keys
  v &= 0x1fffff;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

void morton_order(int32_t N, const double* xc, const double* yc, const double* zc, std::vector<int32_t>& order) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int32_t e = 0; e < N; ++e) {
    lo[0] = std::min(lo[0], xc[e]); hi[0] = std::max(hi[0], xc[e]);
    lo[1] = std::min(lo[1], yc[e]); hi[1] = std::max(hi[1], yc[e]);
    lo[2] = std::min(lo[2], zc[e]); hi[2] = std::max(hi[2], zc[e]);
  }
  double sc[3];
  for (int i = 0; i < 3; ++i) sc[i] = (hi[i] > lo[i]) ? 2097151.0 / (hi[i] - lo[i]) : 0.0;
  std::vector<std::pair<uint64_t, int32_t>> key((size_t)N);
  for (int32_t e = 0; e < N; ++e) {
    uint64_t a = (uint64_t)((xc[e] - lo[0]) * sc[0]), b = (uint64_t)((yc[e] - lo[1]) * sc[1]), c = (uint64_t)((zc[e] - lo[2]) * sc[2]);
    key[e] = {spread21(a) | (spread21(b) << 1) | (spread21(c) << 2), e};
  }
  std::sort(key.begin(), key.end());
  order.resize(N);
  for (int32_t i = 0; i < N; ++i) order[i] = key[i].second;
}

// Level schedule of a sequential sweep: `seq` lists original cells in sweep order (blocks
// concatenated), blk_ptr delimits the blocks.  A cell depends on the same-block neighbours
// that precede it; reversed levels serve the backward sweep.
void build_schedule(const Prep& p, const std::vector<int32_t>& o_nb, const std::vector<int32_t>& seq,
                    const std::vector<int32_t>& blk_ptr, Schedule& S) {
  const int32_t N = p.N, K = p.K, Np = p.Np;
  const int nb_blocks = (int)blk_ptr.size() - 1;
  S.nblocks = nb_blocks;
  S.blk_ptr = blk_ptr;
  std::vector<int32_t> rank((size_t)N), blockof((size_t)N), level((size_t)N, 0);
  for (int b = 0; b < nb_blocks; ++b)
    for (int32_t i = blk_ptr[b]; i < blk_ptr[b + 1]; ++i) { rank[seq[i]] = i; blockof[seq[i]] = b; }
  int32_t maxl = 0;
  for (int32_t i = 0; i < N; ++i) {
    int32_t e = seq[i], l = 0;
    for (int32_t idx = p.row_ptr[e]; idx < p.row_ptr[e + 1]; ++idx) {
      int32_t nb = o_nb[idx];
      if (nb < N && blockof[nb] == blockof[e] && rank[nb] < i) l = std::max(l, level[nb] + 1);
    }
    level[e] = l;
    maxl = std::max(maxl, l);
  }
  S.nlevels = maxl + 1;
  S.lvl_ptr.assign(S.nlevels + 1, 0);
  for (int32_t e = 0; e < N; ++e) S.lvl_ptr[level[e] + 1]++;
  for (int l = 0; l < S.nlevels; ++l) S.lvl_ptr[l + 1] += S.lvl_ptr[l];
  std::vector<int32_t> cur(S.lvl_ptr.begin(), S.lvl_ptr.end() - 1), c2s((size_t)N);
  S.s2c.resize(N);
  for (int32_t c = 0; c < N; ++c) {  // device order inside a level keeps gathers local
    int32_t s = cur[level[p.c2o[c]]]++;
    S.s2c[s] = c;
    c2s[c] = s;
  }
  S.nbs.assign((size_t)K * Np, 0);
  S.bpos.resize(N);
  std::vector<uint8_t> need_lag((size_t)N, 0);
  for (int32_t s = 0; s < N; ++s) {
    int32_t c = S.s2c[s], e = p.c2o[c];
    S.bpos[s] = rank[e];
    for (int k = 0; k < K; ++k) {
      int32_t nb = p.ell_nb[(size_t)k * Np + c], v;
      if (p.ell_fs[(size_t)k * Np + c] == 0) v = s;                 // padding slot, anb == 0
      else if (nb >= N) v = nb;                                      // physical-boundary halo
      else if (blockof[p.c2o[nb]] == blockof[e]) v = c2s[nb];
      else { v = p.H + c2s[nb]; need_lag[c2s[nb]] = 1; }             // other block: lagged copy
      S.nbs[(size_t)k * Np + s] = v;
    }
  }
  S.lag_src.clear();
  for (int32_t s = 0; s < N; ++s) if (need_lag[s]) S.lag_src.push_back(s);
}

}  // namespace

int prepare(Prep& p, int32_t ne, int32_t nf, int32_t nbf, const int32_t* ef2nb_idx,
            const int32_t* ef2nb_nb, const int32_t* ef2nb_fg, const int32_t* s2g, const int32_t* bs,
            const double* xc, const double* yc, const double* zc, int32_t nbc, const int32_t* bc_esec,
            const int32_t* bc_kind, const double* bc_uvw, int32_t n_subdomains,
            const int32_t* g2gf_p, const int32_t* g2gf_idx, int reorder_mode,
            const int32_t* cell2rank, int32_t rank, int32_t nranks) {
  if (ne < 1 || nf < 1 || nbf < 0) return fail(CFDL_ERR_ARG, "cfdl_create: bad sizes ne=%d nf=%d nbf=%d", ne, nf, nbf);
  if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !cell2rank)) return fail(CFDL_ERR_ARG, "cfdl_create: bad rank %d of %d", rank, nranks);
  if (nranks > 1 && n_subdomains > 1) return fail(CFDL_ERR_UNSUPPORTED, "cfdl_create: the pc block solver (n_subdomains>1) is single-GPU; on several GPUs the GPUs are the blocks");
  const int32_t gN = ne, gF = nf, gB = nbf, gH = ne + nbf;
  const int64_t Z64 = 2 * (int64_t)nf - nbf;
  if (Z64 > 0x7fffffff) return fail(CFDL_ERR_RANGE, "cfdl_create: 2nf-nbf exceeds int32");
  const int32_t gZ = (int32_t)Z64;
  p.gN = gN; p.gF = gF; p.gB = gB; p.gZ = gZ;
  p.rank = rank; p.nranks = nranks;
  p.n_subdomains = n_subdomains;
  if (ef2nb_idx[0] != 1 || ef2nb_idx[gN] - 1 != gZ) return fail(CFDL_ERR_MESH, "cfdl_create: ef2nb_idx does not span 2nf-nbf slots");
  p.row_ptr.resize((size_t)gN + 1);
  int K = 0;
  for (int32_t e = 0; e <= gN; ++e) p.row_ptr[e] = ef2nb_idx[e] - 1;
  for (int32_t e = 0; e < gN; ++e) {
    int len = p.row_ptr[e + 1] - p.row_ptr[e];
    if (len < 1 || len > 31) return fail(CFDL_ERR_MESH, "cfdl_create: cell %d has %d faces", e + 1, len);
    K = std::max(K, len);
  }
  p.K = K;
  // unpack (mod_util.f90:1428-1448)
  std::vector<int32_t> o_nb((size_t)gZ), o_fg((size_t)gZ);
  for (int32_t idx = 0; idx < gZ; ++idx) {
    uint32_t pk = (uint32_t)ef2nb_nb[idx];
    int32_t id = (int32_t)(pk >> 5), lf = (int32_t)(pk & 31u), fg = ef2nb_fg[idx];
    if (fg == 0 || std::abs(fg) > gF) return fail(CFDL_ERR_MESH, "cfdl_create: slot %d has face id %d", idx + 1, fg);
    if (lf > 0) { if (id < 1 || id > gN) return fail(CFDL_ERR_MESH, "cfdl_create: slot %d neighbour %d out of range", idx + 1, id); }
    else { if (id <= gN || id > gH || fg < 0) return fail(CFDL_ERR_MESH, "cfdl_create: slot %d halo %d out of range", idx + 1, id); }
    o_nb[idx] = id - 1;
    o_fg[idx] = fg;
  }
  // s2g consistency (owner side carries +fg; calc_mip/update_uvwp start from it)
  for (int32_t f = 0; f < gF; ++f) {
    uint32_t pk = (uint32_t)s2g[f];
    int32_t e = (int32_t)(pk >> 5) - 1, lf = (int32_t)(pk & 31u);
    if (e < 0 || e >= gN || lf < 1 || p.row_ptr[e] + lf - 1 >= p.row_ptr[e + 1] || o_fg[p.row_ptr[e] + lf - 1] != f + 1)
      return fail(CFDL_ERR_MESH, "cfdl_create: s2g(%d) is not the owner slot of the face", f + 1);
  }
  // halo -> (interior cell, local face); the sign bit of bs is a flag, every consumer takes abs
  // (mod_mg_lvl_uns.f90:421-433)
  std::vector<int32_t> halo_e((size_t)gB), halo_lf((size_t)gB);
  for (int32_t j = 0; j < gB; ++j) {
    uint32_t pk = (uint32_t)std::abs(bs[j]);
    halo_e[j] = (int32_t)(pk >> 5) - 1;
    halo_lf[j] = (int32_t)(pk & 31u);
  }
  return prepare_core(p, o_nb, o_fg, halo_e, halo_lf, xc, yc, zc, nbc, bc_esec, bc_kind, bc_uvw, n_subdomains, g2gf_p, g2gf_idx,
                      reorder_mode, cell2rank, rank, nranks);
}

// Everything after unpacking: works on plain int32 arrays, so meshes beyond the reference's
// 2^26 (cell<<5|face) packing limit can be prepared too (cfdl_create_structured_hex).
//   o_nb[slot]  0-based neighbour cell, or gN + halo index for a boundary slot
//   o_fg[slot]  signed 1-based global face id (+ on the owner = lower-numbered cell)
//   p.gN, p.gF, p.gB, p.gZ, p.K, p.row_ptr, p.rank, p.nranks, p.n_subdomains are already set
int prepare_core(Prep& p, const std::vector<int32_t>& o_nb, const std::vector<int32_t>& o_fg, const std::vector<int32_t>& halo_e,
                 const std::vector<int32_t>& halo_lf, const double* xc, const double* yc, const double* zc, int32_t nbc,
                 const int32_t* bc_esec, const int32_t* bc_kind, const double* bc_uvw, int32_t n_subdomains, const int32_t* g2gf_p,
                 const int32_t* g2gf_idx, int reorder_mode, const int32_t* cell2rank, int32_t rank, int32_t nranks) {
  const int32_t gN = p.gN, gF = p.gF, gB = p.gB, gH = p.gN + p.gB;
  const int K = p.K;
  // ---- global base order (natural | Morton) and global greedy colouring --------------------
  std::vector<int32_t> base((size_t)gN), brank((size_t)gN);
  std::iota(base.begin(), base.end(), 0);
  p.morton = false;
  if (reorder_mode != 0) {
    std::vector<int32_t> mo;
    morton_order(gN, xc, yc, zc, mo);
    bool use = (reorder_mode == 1);
    if (reorder_mode == 2) {  // auto: adopt Morton only when the given numbering has poor locality
      std::vector<int32_t> mrank((size_t)gN);
      for (int32_t i = 0; i < gN; ++i) mrank[mo[i]] = i;
      double dn = 0, dm = 0;
      for (int32_t e = 0; e < gN; ++e)
        for (int32_t idx = p.row_ptr[e]; idx < p.row_ptr[e + 1]; ++idx)
          if (o_nb[idx] < gN) { dn += std::abs((double)o_nb[idx] - e); dm += std::abs((double)mrank[o_nb[idx]] - mrank[e]); }
      use = dn > 4.0 * dm;
    }
    if (use) { base.swap(mo); p.morton = true; }
  }
  for (int32_t i = 0; i < gN; ++i) brank[base[i]] = i;
  std::vector<int8_t> color((size_t)gN, -1);
  int ncol = 0;
  for (int32_t i = 0; i < gN; ++i) {
    int32_t e = base[i];
    uint32_t used = 0;
    for (int32_t idx = p.row_ptr[e]; idx < p.row_ptr[e + 1]; ++idx)
      if (o_nb[idx] < gN && color[o_nb[idx]] >= 0) used |= 1u << color[o_nb[idx]];
    int c = 0;
    while (used & (1u << c)) ++c;
    color[e] = (int8_t)c;
    ncol = std::max(ncol, c + 1);
  }
  p.ncolors = ncol;
  auto owner = [&](int32_t e) -> int32_t { return nranks > 1 ? cell2rank[e] - 1 : 0; };
  if (nranks > 1)
    for (int32_t e = 0; e < gN; ++e)
      if (cell2rank[e] < 1 || cell2rank[e] > nranks) return fail(CFDL_ERR_ARG, "cfdl_create: cell2rank(%d)=%d outside 1..%d", e + 1, cell2rank[e], nranks);
  p.ref_cell_owner = owner(0);
  // ---- owned cells: colour-major, base order inside a colour ---------------------------------
  p.color_ptr.assign(ncol + 1, 0);
  int32_t N = 0;
  for (int32_t e = 0; e < gN; ++e) if (owner(e) == rank) { p.color_ptr[color[e] + 1]++; ++N; }
  if (N < 1) return fail(CFDL_ERR_ARG, "cfdl_create: rank %d owns no cell", rank);
  for (int c = 0; c < ncol; ++c) p.color_ptr[c + 1] += p.color_ptr[c];
  p.o2c.assign(gN, -1);
  p.c2o.resize(N);
  p.loc_order.clear();
  p.loc_order.reserve(N);
  p.color_if.assign(ncol, 0);
  {
    // inside a colour the interface cells (those with a neighbour owned by another rank) come
    // first, so a pass can compute and ship them before it sweeps the interior
    std::vector<uint8_t> is_if;
    if (nranks > 1) {
      is_if.assign((size_t)gN, 0);
      for (int32_t e = 0; e < gN; ++e) {
        if (owner(e) != rank) continue;
        for (int32_t idx = p.row_ptr[e]; idx < p.row_ptr[e + 1]; ++idx)
          if (o_nb[idx] < gN && owner(o_nb[idx]) != rank) { is_if[e] = 1; break; }
        if (is_if[e]) p.color_if[color[e]]++;
      }
    }
    std::vector<int32_t> cur_if(p.color_ptr.begin(), p.color_ptr.end() - 1), cur_in(ncol);
    for (int c = 0; c < ncol; ++c) cur_in[c] = p.color_ptr[c] + p.color_if[c];
    for (int32_t i = 0; i < gN; ++i) {
      int32_t e = base[i];
      if (owner(e) != rank) continue;
      int32_t c = (nranks > 1 && is_if[e]) ? cur_if[color[e]]++ : cur_in[color[e]]++;
      p.c2o[c] = e; p.o2c[e] = c;
      p.loc_order.push_back(c);
    }
  }
  // ---- ghost cells (owned by other ranks, face-adjacent to an owned cell) and interface lists -
  struct Key { int32_t r, c, b, e; };
  auto key_less = [](const Key& a, const Key& b) { return a.r != b.r ? a.r < b.r : (a.c != b.c ? a.c < b.c : a.b < b.b); };
  std::vector<Key> ghosts, sends;
  if (nranks > 1) {
    std::vector<uint8_t> is_ghost((size_t)gN, 0);
    for (int32_t c = 0; c < N; ++c) {
      int32_t e = p.c2o[c];
      uint64_t sent_to = 0;  // neighbour ranks this cell was already listed for (nranks <= 64 checked below)
      for (int32_t idx = p.row_ptr[e]; idx < p.row_ptr[e + 1]; ++idx) {
        int32_t nb = o_nb[idx];
        if (nb >= gN || owner(nb) == rank) continue;
        if (!is_ghost[nb]) { is_ghost[nb] = 1; ghosts.push_back({owner(nb), color[nb], brank[nb], nb}); }
        const int32_t r = owner(nb);
        if (r < 64) { if (sent_to >> r & 1) continue; sent_to |= uint64_t(1) << r; }
        sends.push_back({r, color[e], brank[e], e});
      }
    }
    if (nranks > 64) return fail(CFDL_ERR_UNSUPPORTED, "cfdl_create: more than 64 ranks");
    std::sort(ghosts.begin(), ghosts.end(), key_less);
    std::sort(sends.begin(), sends.end(), key_less);
  }
  const int32_t G = (int32_t)ghosts.size(), Nc = N + G;
  p.N = N; p.G = G; p.Nc = Nc;
  p.Np = (N + 31) / 32 * 32;
  const int32_t Np = p.Np;
  p.c2o.resize(Nc);
  for (int32_t g = 0; g < G; ++g) { p.c2o[N + g] = ghosts[g].e; p.o2c[ghosts[g].e] = N + g; }
  p.nbr_rank.clear();
  for (const Key& k : ghosts) if (p.nbr_rank.empty() || p.nbr_rank.back() != k.r) p.nbr_rank.push_back(k.r);
  {
    std::vector<int32_t> chk;
    for (const Key& k : sends) if (chk.empty() || chk.back() != k.r) chk.push_back(k.r);
    if (chk != p.nbr_rank) return fail(CFDL_ERR_INTERNAL, "cfdl_create: send and receive neighbour sets differ");
  }
  const int nnbr = (int)p.nbr_rank.size();
  p.recv_ptr.assign((size_t)nnbr * ncol + 1, 0);
  p.send_ptr.assign((size_t)nnbr * ncol + 1, 0);
  p.send_cells.resize(sends.size());
  {
    auto slot_of = [&](const Key& k) { return (size_t)(std::lower_bound(p.nbr_rank.begin(), p.nbr_rank.end(), k.r) - p.nbr_rank.begin()) * ncol + k.c; };
    for (const Key& k : ghosts) p.recv_ptr[slot_of(k) + 1]++;
    for (const Key& k : sends) p.send_ptr[slot_of(k) + 1]++;
    for (size_t i = 0; i < (size_t)nnbr * ncol; ++i) { p.recv_ptr[i + 1] += p.recv_ptr[i]; p.send_ptr[i + 1] += p.send_ptr[i]; }
    for (size_t i = 0; i < sends.size(); ++i) p.send_cells[i] = p.o2c[sends[i].e];
    // the same lists seen from the cell: which ghost slots on which neighbours mirror cell c
    p.tgt_ptr.assign((size_t)N + 1, 0);
    for (size_t i = 0; i < sends.size(); ++i) p.tgt_ptr[p.send_cells[i] + 1]++;
    for (int32_t c = 0; c < N; ++c) p.tgt_ptr[c + 1] += p.tgt_ptr[c];
    p.tgt_nbr.resize(sends.size()); p.tgt_pos.resize(sends.size());
    std::vector<int32_t> fill(p.tgt_ptr.begin(), p.tgt_ptr.end() - 1);
    for (int r = 0; r < nnbr; ++r)
      for (int c = 0; c < ncol; ++c)
        for (int32_t s = p.send_ptr[(size_t)r * ncol + c]; s < p.send_ptr[(size_t)r * ncol + c + 1]; ++s) {
          const int32_t at = fill[p.send_cells[s]]++;
          p.tgt_nbr[at] = r;
          p.tgt_pos[at] = s - p.send_ptr[(size_t)r * ncol + c];
        }
  }
  // ---- halos: physical boundary faces of owned cells, in original halo order ----------------
  p.h2o.clear();
  std::vector<int32_t> o2h((size_t)gB, -1);
  for (int32_t j = 0; j < gB; ++j) {
    const int32_t e = halo_e[j], lf = halo_lf[j];
    if (e < 0 || e >= gN || lf < 1 || lf > p.row_ptr[e + 1] - p.row_ptr[e]) return fail(CFDL_ERR_MESH, "cfdl_create: bs(%d) invalid", gN + 1 + j);
    if (o_nb[p.row_ptr[e] + lf - 1] != gN + j) return fail(CFDL_ERR_MESH, "cfdl_create: bs(%d) does not point back to its halo", gN + 1 + j);
    if (owner(e) == rank) { o2h[j] = (int32_t)p.h2o.size(); p.h2o.push_back(j); }
  }
  const int32_t B = (int32_t)p.h2o.size(), H = Nc + B;
  p.B = B; p.H = H;
  // ---- faces touching an owned cell: cell-cell faces first (first-touch order), then boundary -
  // A face belongs to the first device cell that touches it; faces are then numbered SLOT-MAJOR
  // over those cells (all slot-0 faces in cell order, then slot 1, ...), so that for a fixed slot
  // consecutive cells read consecutive faces — on a two-colour mesh every face is first touched
  // by a colour-0 cell, and the other colour meets them through consecutive neighbours as well.
  std::vector<int32_t> o2f((size_t)gF, -1);
  p.f2o.clear(); p.face_a.clear(); p.face_b.clear(); p.fown.clear();
  {
    std::vector<int32_t> toucher((size_t)gF, -1);
    p.ftouch.assign((size_t)N, 0);
    p.touch_end = 0;
    for (int32_t c = 0; c < N; ++c) {
      int32_t e = p.c2o[c];
      for (int32_t idx = p.row_ptr[e]; idx < p.row_ptr[e + 1]; ++idx) {
        int32_t f = std::abs(o_fg[idx]) - 1;
        if (o_nb[idx] < gN && toucher[f] < 0) toucher[f] = c;
      }
    }
    for (int k = 0; k < K; ++k)
      for (int32_t c = 0; c < N; ++c) {
        int32_t e = p.c2o[c];
        if (k >= p.row_ptr[e + 1] - p.row_ptr[e]) continue;
        const int32_t idx = p.row_ptr[e] + k;
        int32_t nb = o_nb[idx], f = std::abs(o_fg[idx]) - 1;
        if (nb >= gN || toucher[f] != c || o2f[f] >= 0) continue;
        p.ftouch[c] |= (uint8_t)(1u << k);
        p.touch_end = std::max(p.touch_end, c + 1);
        o2f[f] = (int32_t)p.f2o.size();
        p.f2o.push_back(f);
        const int32_t nbd = p.o2c[nb];
        if (nbd < 0) return fail(CFDL_ERR_INTERNAL, "cfdl_create: neighbour cell without device index");
        const bool plus = o_fg[idx] > 0;
        p.face_a.push_back(plus ? c : nbd);
        p.face_b.push_back(plus ? nbd : c);
        p.fown.push_back((plus ? c : nbd) < N ? 1 : 0);
      }
  }
  p.Fi = (int32_t)p.f2o.size();
  p.halo_cell.assign(B, -1); p.halo_face.assign(B, -1); p.halo_bc.assign(B, -1); p.halo_slot.assign(B, 0);
  for (int32_t jl = 0; jl < B; ++jl) {
    const int32_t j = p.h2o[jl];
    const int32_t e = halo_e[j], lf = halo_lf[j], idx = p.row_ptr[e] + lf - 1, f = o_fg[idx] - 1;
    if (o2f[f] != -1) return fail(CFDL_ERR_MESH, "cfdl_create: boundary face %d used twice", f + 1);
    o2f[f] = p.Fi + jl;
    p.f2o.push_back(f); p.face_a.push_back(p.o2c[e]); p.face_b.push_back(Nc + jl); p.fown.push_back(1);
    p.halo_cell[jl] = p.o2c[e]; p.halo_slot[jl] = (uint8_t)(lf - 1); p.halo_face[jl] = p.Fi + jl;
  }
  p.F = (int32_t)p.f2o.size();
  if (nranks == 1 && (p.F != gF || p.Fi != gF - gB)) return fail(CFDL_ERR_MESH, "cfdl_create: %d faces found, expected %d", p.F, gF);
  int64_t zloc = 0;
  for (int32_t c = 0; c < N; ++c) zloc += p.row_ptr[p.c2o[c] + 1] - p.row_ptr[p.c2o[c]];
  p.Z = (int32_t)zloc;
  // ---- ELL slot arrays (owned rows) ------------------------------------------------------------
  p.ell_nb.assign((size_t)K * Np, 0); p.ell_fs.assign((size_t)K * Np, 0); p.nfc.assign(N, 0);
  for (int32_t c = 0; c < Np; ++c)
    for (int k = 0; k < K; ++k) p.ell_nb[(size_t)k * Np + c] = std::min(c, N - 1);
  for (int32_t c = 0; c < N; ++c) {
    int32_t e = p.c2o[c];
    p.nfc[c] = (uint8_t)(p.row_ptr[e + 1] - p.row_ptr[e]);
    for (int32_t idx = p.row_ptr[e]; idx < p.row_ptr[e + 1]; ++idx) {
      int k = idx - p.row_ptr[e];
      int32_t nb = o_nb[idx], fdev = o2f[std::abs(o_fg[idx]) - 1];
      int32_t nbd = (nb < gN) ? p.o2c[nb] : Nc + o2h[nb - gN];
      p.ell_nb[(size_t)k * Np + c] = nbd;
      p.ell_fs[(size_t)k * Np + c] = (o_fg[idx] > 0) ? (fdev + 1) : -(fdev + 1);
    }
  }
  // ---- 16-bit neighbour offsets for the pc passes (two colours, one rank) ------------------------
  p.nb16.clear();
  p.nb16_ok = false;
  p.color_dist = -1;
  if (nranks == 1 && ncol == 2) {
    const int32_t nred = p.color_ptr[1];
    int64_t dist = 0;
    bool proper = true;
    for (int32_t c = 0; c < N && proper; ++c) {
      const bool red = c < nred;
      for (int k = 0; k < p.nfc[c]; ++k) {
        const int32_t nb = p.ell_nb[(size_t)k * Np + c];
        if (nb >= N) continue;
        if ((nb < nred) == red) { proper = false; break; }
        dist = std::max<int64_t>(dist, std::llabs((int64_t)(red ? nb - nred : nb) - (red ? c : c - nred)));
      }
    }
    if (proper && dist < INT32_MAX) p.color_dist = (int32_t)dist;
    p.nb16.assign((size_t)K * Np, 0);
    bool ok = true;
    for (int32_t c = 0; c < N && ok; ++c) {
      const bool red = c < nred;
      const int32_t cl = red ? c : c - nred;
      int32_t d[8], first = INT32_MIN;
      for (int k = 0; k < K; ++k) {
        d[k] = INT32_MIN;
        const int32_t nb = p.ell_nb[(size_t)k * Np + c];
        if (k >= p.nfc[c] || nb >= N) continue;          // padding or boundary slot
        if ((nb < nred) == red) { ok = false; break; }   // a neighbour of the same colour: not a proper two-colouring
        d[k] = (red ? nb - nred : nb) - cl;
        if (d[k] < -32768 || d[k] > 32767) { ok = false; break; }
        if (first == INT32_MIN) first = d[k];
      }
      if (first == INT32_MIN) ok = false;                // a cell without cell neighbours
      for (int k = 0; k < K && ok; ++k) p.nb16[(size_t)k * Np + c] = (int16_t)(d[k] == INT32_MIN ? first : d[k]);
    }
    p.nb16_ok = ok;
    if (!ok) p.nb16.clear();
  }
  // ---- boundary conditions -------------------------------------------------------------------
  p.bc_kind.assign(bc_kind, bc_kind + nbc);
  p.bc_uvw.assign(bc_uvw, bc_uvw + 3 * (size_t)nbc);
  for (int32_t i = 0; i < nbc; ++i) {
    if (bc_kind[i] < CFDL_BC_WALL || bc_kind[i] > CFDL_BC_SYMMETRY) return fail(CFDL_ERR_UNSUPPORTED, "cfdl_create: bc kind %d", bc_kind[i]);
    int32_t a = bc_esec[2 * i], b = bc_esec[2 * i + 1];
    if (a < gN + 1 || b > gH || a > b + 1) return fail(CFDL_ERR_ARG, "cfdl_create: bc %d halo range [%d,%d] invalid", i, a, b);
    for (int32_t h = a; h <= b; ++h) if (o2h[h - gN - 1] >= 0) p.halo_bc[o2h[h - gN - 1]] = i;
  }
  // ---- level schedules reproducing the reference's sequential sweeps (single rank) -----------
  if (nranks == 1) {
    {
      std::vector<int32_t> seq((size_t)N), bp = {0, N};
      std::iota(seq.begin(), seq.end(), 0);
      build_schedule(p, o_nb, seq, bp, p.natural);
    }
    if (n_subdomains > 1) {
      if (!g2gf_p || !g2gf_idx) return fail(CFDL_ERR_ARG, "cfdl_create: n_subdomains>1 needs g2gf_p and g2gf_idx");
      std::vector<int32_t> seq((size_t)N), bp((size_t)n_subdomains + 1);
      std::vector<uint8_t> seen((size_t)N, 0);
      for (int b = 0; b <= n_subdomains; ++b) bp[b] = g2gf_idx[b] - 1;
      if (bp[0] != 0 || bp[n_subdomains] != N) return fail(CFDL_ERR_ARG, "cfdl_create: g2gf_idx must span 1..ne+1");
      for (int32_t i = 0; i < N; ++i) {
        int32_t e = g2gf_p[i] - 1;
        if (e < 0 || e >= N || seen[e]) return fail(CFDL_ERR_ARG, "cfdl_create: g2gf_p is not a permutation of the cells");
        seen[e] = 1; seq[i] = e;
      }
      build_schedule(p, o_nb, seq, bp, p.blocks);
    }
  }
  return CFDL_OK;
}

}  // namespace cfdl
