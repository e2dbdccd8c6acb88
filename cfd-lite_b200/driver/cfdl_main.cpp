// C++ re-creation of the reference's driver loop (src/main.f90:24-93) on top of the C ABI of
// libcfdl.so — the host side a Fortran build would provide through fortran/mod_gpu_bridge.f90.
// The container has no Fortran compiler and no CGNS library, so the mesh comes from the
// synthetic generator (the stand-in for cell_input.f90's CGNS read) instead of a .cgns file.
//
//   cfdl_main <hex|tet> <n> [solver=parity|mcsgs|pcg] [ntstep=10] [ncoef=3] [n_subdomains=1]
//
// Prints the reference's residual table (main.f90:49, mod_solver.f90:6,184,325) and the
// cell-iterations/s of the loop.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "cfdl.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc__ = (call);                                                               \
    if (rc__) { std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, cfdl_last_error()); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s <hex|tet|file:mesh.raw> <n> [solver] [ntstep] [ncoef] [n_subdomains]\n", argv[0]); return 2; }
  const int kind = !std::strcmp(argv[1], "tet") ? CFDL_MESH_TET : CFDL_MESH_HEX;
  const int n = std::atoi(argv[2]);
  const std::string solver = argc > 3 ? argv[3] : "parity";
  const int ntstep = argc > 4 ? std::atoi(argv[4]) : 10;   // phys_t defaults, mod_physics.f90:15-19
  const int ncoef = argc > 5 ? std::atoi(argv[5]) : 3;
  const int nsub = argc > 6 ? std::atoi(argv[6]) : 1;
  const double dt = 0.01;
  const int nit = 100;

  // the mesh: a raw mesh file (file:<path>, the CGNS-free input of cfdl_rawmesh_write / mod_rawmesh.f90) or a synthetic cube
  int64_t nvx, ne64 = 0, nbf64 = 0;
  int nsec, w;
  std::vector<double> x, y, z;
  std::vector<int32_t> e2vx, etype, esec;
  std::vector<char> names;
  int64_t faces2 = 0;  // sum of faces per 3-D cell
  if (!std::strncmp(argv[1], "file:", 5)) {
    int64_t nelem;
    CHECK(cfdl_rawmesh_sizes(argv[1] + 5, &nvx, &nelem, &nsec, &w));
    x.resize(nvx); y.resize(nvx); z.resize(nvx); e2vx.resize((size_t)w * nelem); etype.resize(nsec); esec.resize(2 * nsec); names.resize(32 * nsec);
    CHECK(cfdl_rawmesh_read(argv[1] + 5, x.data(), y.data(), z.data(), etype.data(), esec.data(), names.data(), e2vx.data()));
  } else {
    CHECK(cfdl_meshgen_sizes(kind, n, &nvx, &ne64, &nbf64, &nsec, &w));
    x.resize(nvx); y.resize(nvx); z.resize(nvx); e2vx.resize((size_t)w * (ne64 + nbf64)); etype.resize(nsec); esec.resize(2 * nsec); names.resize(32 * nsec);
    CHECK(cfdl_meshgen_fill(kind, n, kind == CFDL_MESH_TET ? 0.2 : 0.0, kind == CFDL_MESH_TET, 12345, x.data(), y.data(), z.data(),
                            e2vx.data(), etype.data(), esec.data(), names.data()));
  }
  ne64 = nbf64 = 0;
  for (int s = 0; s < nsec; ++s) {  // cell_input.f90:58-71 (faces per cell: mod_util.f90:168)
    const int64_t cnt = esec[2 * s + 1] - esec[2 * s] + 1;
    const int t = etype[s];
    if (t >= 10 && t <= 20) { ne64 += cnt; faces2 += cnt * (t == 17 ? 6 : t == 10 ? 4 : 5); }
    else nbf64 += cnt;
  }
  const int32_t ne = (int32_t)ne64, nbf = (int32_t)nbf64;
  const int32_t nf = (int32_t)((faces2 + nbf) / 2);
  const int64_t Z = 2 * (int64_t)nf - nbf, H = (int64_t)ne + nbf;
  std::vector<int32_t> idx(ne + 1), nb(Z), fg(Z), s2g(nf), bs(nbf);
  std::vector<double> xc(H), yc(H), zc(H), aip(3 * (size_t)nf), rip(3 * (size_t)nf), vol(ne);
  CHECK(cfdl_mesh_build(nvx, x.data(), y.data(), z.data(), nsec, etype.data(), esec.data(), w, e2vx.data(), ne, nf, nbf, idx.data(),
                        nb.data(), fg.data(), s2g.data(), bs.data(), xc.data(), yc.data(), zc.data(), aip.data(), rip.data(), vol.data()));
  // BCs of construct_uvwp (mod_uvwp.f90:73-78): lid on 'top', no-slip elsewhere, section order
  std::vector<int32_t> bc_esec, bc_kind;
  std::vector<double> bc_uvw;
  for (int s = 0; s < nsec; ++s) {
    if (etype[s] >= 10) continue;
    std::string nm(names.data() + 32 * s, 32);
    nm.erase(nm.find_last_not_of(' ') + 1);
    bc_esec.push_back(esec[2 * s]); bc_esec.push_back(esec[2 * s + 1]);
    const bool lid = std::string("top").find(nm) != std::string::npos && !nm.empty();
    bc_kind.push_back(lid ? CFDL_BC_LID : CFDL_BC_WALL);
    bc_uvw.push_back(lid ? 1.0 : 0.0); bc_uvw.push_back(0.0); bc_uvw.push_back(0.0);
  }
  std::vector<double> rho(ne, 5.0), mu(ne, 0.01);  // init_properties, mod_properties.f90:86-87
  std::vector<int32_t> g2gf_p, g2gf_idx;
  if (nsub > 1) {
    std::vector<int32_t> cell2sub(ne);
    g2gf_p.resize(ne); g2gf_idx.resize(nsub + 1);
    CHECK(cfdl_partition_rcb(ne, xc.data(), yc.data(), zc.data(), vol.data(), nsub, cell2sub.data(), g2gf_p.data(), g2gf_idx.data()));
  }
  cfdl_handle h = nullptr;
  CHECK(cfdl_create(&h, ne, nf, nbf, idx.data(), nb.data(), fg.data(), s2g.data(), bs.data(), xc.data(), yc.data(), zc.data(),
                    aip.data(), rip.data(), vol.data(), rho.data(), mu.data(), (int32_t)bc_kind.size(), bc_esec.data(), bc_kind.data(),
                    bc_uvw.data(), nsub, nsub > 1 ? g2gf_p.data() : nullptr, nsub > 1 ? g2gf_idx.data() : nullptr, 0));
  CHECK(cfdl_set_option(h, "solver", solver == "mcsgs" ? CFDL_SOLVER_MCSGS : solver == "pcg" ? CFDL_SOLVER_PCG : CFDL_SOLVER_PARITY));

  std::printf("     %-16s %5s %15s%9s   %9s   %9s\n", "solve eqn", "nit", "residual(rms)", "initial", "final", "max");
  static const char* eqn[4] = {"u", "v", "w", "pc"};
  const auto t0 = std::chrono::steady_clock::now();
  for (int tstep = 1; tstep <= ntstep; ++tstep) {
    for (int icoef = 1; icoef <= ncoef; ++icoef) {
      double hist[16];
      CHECK(cfdl_update_boundaries(h));
      CHECK(cfdl_solve_uvwp(h, dt, nit, hist));
      for (int k = 0; k < 4; ++k)
        std::printf("     %-16s %5d %15s%9.3E   %9.3E   %9.3E\n", eqn[k], (int)hist[4 * k], "", hist[4 * k + 1], hist[4 * k + 2], hist[4 * k + 3]);
    }
    CHECK(cfdl_update_time(h));
    std::printf("------------------------------------time step(%5d)\n", tstep);
  }
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("%d cells, %d SIMPLE iterations in %.4f s: %.4g cell-iterations/s (%s)\n", ne, ntstep * ncoef, sec,
              (double)ne * ntstep * ncoef / sec, solver.c_str());
  CHECK(cfdl_destroy(h));
  return 0;
}
