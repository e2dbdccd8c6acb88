/* cfdl.h — C ABI of the B200-native SIMPLE hot path of CFD-Lite.
 *
 * The reference has no FFI registry for this path: the boundary is the set of Fortran call
 * sites inside solve_uvwp (src/equations/mod_uvwp.f90:95-134) and the two per-iteration
 * calls in src/main.f90:52,63.  Every entry point below names the reference routine it
 * replaces.  Conventions follow the reference's own bind(C) precedent
 * (src/VTK/mod_vtk.f90:3-29): plain `double*` / `int32_t*`, arrays passed as their first
 * element, nothing retained beyond the call except what cfdl_create copies to the device.
 *
 * Index conventions at this boundary are the reference's:
 *   - cells 1..ne, halo (boundary-face) cells ne+1..ne+nbf, faces 1..nf, all 1-based;
 *   - ef2nb(:,1) packs (neighbour<<5)|neighbour_local_face, local_face==0 <=> halo
 *     (src/modules/mod_util.f90:1428-1448); ef2nb(:,2) is the signed global face id;
 *   - vectors (aip, rip, gu, gv, gw, gp, gpc) are AoS xyz.
 * All functions return 0 on success, a CFDL_ERR_* code otherwise; cfdl_last_error() gives
 * the message.  The library never calls abort()/exit().  There is NO CPU fallback: every
 * compute entry point fails with CFDL_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef CFDL_H
#define CFDL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  CFDL_OK = 0,
  CFDL_ERR_ARG = 1,      /* bad argument */
  CFDL_ERR_RANGE = 2,    /* size exceeds an index limit */
  CFDL_ERR_CUDA = 3,     /* CUDA runtime error / no device */
  CFDL_ERR_NCCL = 4,     /* NCCL error / library not loadable */
  CFDL_ERR_MESH = 5,     /* inconsistent connectivity (reference: `stop` in find_element_nb) */
  CFDL_ERR_INTERNAL = 6,
  CFDL_ERR_UNSUPPORTED = 7,
  CFDL_ERR_COMM = 8      /* a wait on another rank's flag word ran into its time limit (a rank died or never launched) */
};

/* boundary-condition kinds == the reference's BC callbacks (mod_uvwp.f90:493-570) */
enum {
  CFDL_BC_WALL = 0,     /* dirichlet0: u=v=w=0, bc_type 'dirichlet' */
  CFDL_BC_LID = 1,      /* lid: (u,v,w)=bc_uvw (reference: 1,0,0), 'dirichlet' */
  CFDL_BC_SYMMETRY = 2  /* symmetry: mirrored velocity, 'zero_flux' (mod_uvwp.f90:517-546; assembly :266-268) */
  /* NOT offered: the reference's `inlet` / `outlet` callbacks (mod_uvwp.f90:571-606).  They set bc_type 'inlet' /
     'outlet_p', which neither `select case` of the assembly handles (calc_coef_uvw :254-269 knows 'dirichlet' and
     'zero_flux' only, calc_coef_p :355-360 likewise): a boundary face of such a section adds the values `d`, `f`
     left over from whatever face was processed before it — undefined behaviour in the reference itself, no BC the
     reference binds uses them (construct_uvwp :73-78), so there is nothing to be faithful to.  cfdl_create rejects
     any other kind with CFDL_ERR_UNSUPPORTED. */
};

/* linear-solver modes */
enum {
  CFDL_SOLVER_PARITY = 0, /* exact reference order: natural-order SGS (solve_gs) for u,v,w and,
                             for pc, solve_gs (n_subdomains==1) or the block-SGS of
                             multi_subdomain_solver (n_subdomains>1); level-scheduled on the GPU */
  CFDL_SOLVER_MCSGS = 1,  /* multicolour symmetric Gauss-Seidel: same update formula, same
                             stopping rule, colour order instead of natural order */
  CFDL_SOLVER_PCG = 2     /* pc: Jacobi-preconditioned conjugate gradients with the reference's stopping rule (not a
                             restatement of solve_gs: same interface, fewer matrix passes); u,v,w: MCSGS; partitioned handles:
                             ghost exchange of the search direction + all-reduced dot products per iteration */
};

/* field selectors for upload/download and the per-routine entry points */
enum {
  CFDL_F_U = 0, CFDL_F_V, CFDL_F_W, CFDL_F_P,         /* (ne+nbf) */
  CFDL_F_U0, CFDL_F_V0, CFDL_F_W0, CFDL_F_PC,         /* (ne+nbf) */
  CFDL_F_GU, CFDL_F_GV, CFDL_F_GW, CFDL_F_GP, CFDL_F_GPC, /* 3*(ne+nbf), AoS */
  CFDL_F_MIP, CFDL_F_MIP0,                            /* (nf) */
  CFDL_F_BU, CFDL_F_BV, CFDL_F_BW, CFDL_F_D, CFDL_F_DC, /* (ne) */
  CFDL_F_AP, CFDL_F_B,                                /* (ne)  shared ap / b of phys_t */
  CFDL_F_ANB,                                         /* (2nf-nbf) CSR order of ef2nb */
  /* energy_t (mod_energy.f90:5-10: t, phi = enthalpy cp t, phi0) and scalar_t (mod_scalar.f90:5-12: phi, phi0);
     allocated by cfdl_energy_init / cfdl_scalar_init */
  CFDL_F_T, CFDL_F_H, CFDL_F_H0, CFDL_F_S, CFDL_F_S0, /* (ne+nbf) */
  CFDL_F_GT, CFDL_F_GH, CFDL_F_GS,                    /* 3*(ne+nbf), AoS: gt, grad of the energy equation, grad of the scalar */
  CFDL_F_COUNT
};

/* equation selector replacing the Fortran `cname` string (mod_solver.f90:268-269) */
enum { CFDL_EQ_U = 0, CFDL_EQ_V = 1, CFDL_EQ_W = 2, CFDL_EQ_PC = 3, CFDL_EQ_E = 4 /* energy */, CFDL_EQ_S = 5 /* scalar */ };

typedef struct cfdl_handle_s* cfdl_handle;

const char* cfdl_last_error(void);
int cfdl_version(void);
/* number of usable CUDA devices (0 on a CPU-only box; never an error) */
int cfdl_device_count(void);

/* ---- lifecycle: replaces construct_physics / construct_uvwp / construct_subdomains
 *      (mod_physics.f90:52-75, mod_uvwp.f90:20-84, mod_subdomains.f90:18-160).
 *  Mesh arrays are geometry_t + meshds_t as built by cell_input:
 *    ef2nb_idx(ne+1), ef2nb_nb(Z) = ef2nb(:,1), ef2nb_fg(Z) = ef2nb(:,2), Z = 2nf-nbf,
 *    s2g(nf), bs(nbf) (halo ne+1.. first), xc,yc,zc(ne+nbf), aip(3nf), rip(3nf), vol(ne),
 *    rho,mu(ne).
 *  BCs: bc_esec(2,nbc) halo ranges in 2-D section order (mod_eqn_setup.f90:60), bc_kind(nbc),
 *    bc_uvw(3,nbc).
 *  Subdomains (pc block solver): n_subdomains, and when >1 g2gf_p(ne) = cells sorted by block
 *    in the reference's own (unstable-sort) order and g2gf_idx(n_subdomains+1)
 *    (mod_mg_lvl_uns.f90:883-902); both 1-based.  The library never recomputes the RCB.
 *  device: CUDA device ordinal.  Fields start at the reference's initial state (all zero,
 *  mod_uvwp.f90:57-69) including mip from calc_mip(.false.) (:81-82). */
int cfdl_create(cfdl_handle* out, int32_t ne, int32_t nf, int32_t nbf,
                const int32_t* ef2nb_idx, const int32_t* ef2nb_nb, const int32_t* ef2nb_fg,
                const int32_t* s2g, const int32_t* bs,
                const double* xc, const double* yc, const double* zc,
                const double* aip, const double* rip, const double* vol,
                const double* rho, const double* mu,
                int32_t nbc, const int32_t* bc_esec, const int32_t* bc_kind, const double* bc_uvw,
                int32_t n_subdomains, const int32_t* g2gf_p, const int32_t* g2gf_idx,
                int32_t device);
int cfdl_destroy(cfdl_handle h);

/* Options (key, value); unknown keys return CFDL_ERR_ARG.  What a caller may want to set:
 *   "solver"          CFDL_SOLVER_PARITY (the reference's sweep order, one GPU) | CFDL_SOLVER_MCSGS (multicolour order) |
 *                     CFDL_SOLVER_PCG (conjugate gradients for pc)
 *   "pcg_precond"     1 multicolour-SSOR preconditioner on two-colour meshes (default), 0 Jacobi
 *   "uvw_fused"       1 the three momentum solves side by side (default), 0 one after the other (same bits)
 *   "p2p"             1 peer-to-peer ghost exchange once cfdl_comm_ipc_connect has run (default), 0 NCCL send/recv
 *   "rbq"             1 the pc solve as one persistent launch where it pays (default), 0 one launch per colour pass
 *   "rbq_counter"     -1 chunks handed out from a counter where one chunk per CTA would be too long (default), 0 never, 1 always
 *   "rbq_l2_fraction" largest share of the L2 the value arrays may take for the persistent form (default 0.55)
 *   "rbq_prefetch"    trips ahead the constants are prefetched into the L2 in the persistent form (-1 by mesh size, 0 off)
 *   "statics"         1 face statics precomputed at creation (default; needed by cfdl_solve_energy), 0 geometry re-evaluated
 *   "profile"         0 off, 1 one CUDA-event pair per launch, 2 one per batch of passes (bench.py's phase times)
 * The remaining keys ("fused", "pdl", "pdl_rows", "rb_idx16", "pc_sumap", "grad_variant", "mip_hoist", "occ_grids",
 * "ctas_per_sm", "rbq_lmax", "rbq_lbig", "rbq_cap", "rbq_ctas", "reset_counters") select measured alternatives of single
 * kernels for tests and tuning experiments; every setting gives the same results (tests/test_gpu_parity.py).
 * On partitioned handles the rbq_* options must be set before cfdl_comm_ipc_handle, which fixes the chunk geometry. */
int cfdl_set_option(cfdl_handle h, const char* key, double value);
int cfdl_get_info(cfdl_handle h, const char* key, double* value);
/* device cell numbering (colour-major, optionally Morton inside a colour): c2o[i] = 1-based
 * reference cell stored at device position i; colour c occupies [color_ptr[c], color_ptr[c+1]).
 * color_ptr has ncolors+1 entries ("ncolors" via cfdl_get_info); either pointer may be NULL. */
int cfdl_get_cell_order(cfdl_handle h, int32_t* c2o, int32_t* color_ptr);

/* ---- instrumentation (bench.py): CUDA events on the library's stream, pinned host buffers.
 * set_option keys: "solver", "profile" (0 off; 1: an event pair around every launch of the profiled
 *                  kernels; 2: an event pair around every batch of back-to-back solver passes),
 *                  "reset_counters" (any value); tuning switches: "fused", "pdl", "pdl_rows", "p2p",
 *                  "statics", "mip_variant", "occ_grids", "ctas_per_sm";
 *                  "autotune" (1 default: bit-identical kernel variants are timed on first use and the
 *                  fastest kept; 0: round-1 defaults; 2: forget what was measured), and the switches
 *                  that pin a choice by hand: "uvw_variant" (2..14), "grad_variant" (0..3, -1 measured),
 *                  "coef_p_variant" (0..5, -1), "mip_fast" (0/1, -1), "correct_fast" (0/1),
 *                  "uvw_fused" (momentum solves side by side: 0/1, -1 measured), "rb_persistent" (fused
 *                  passes of a batch in one cooperative launch: 0/1, -1 measured), "rb_keep_mb" (megabytes of the pc
 *                  coefficients hinted to stay in L2 across passes: >= 0 pinned, -1 measured), "rb_idx16" (16-bit
 *                  neighbour offsets in the pc passes: 0/1, -1 measured), "rb_wave" (pass teams of the temporally
 *                  blocked pc solve, 0 = not pinned) with "rb_wave_block" (iterations per launch) and "rb_wave_rows"
 *                  (rows per thread and chunk), "pc_sumap" (0/1).
 * get_info keys:   "launches" (kernels launched since reset), "prof_ms_<k>" / "prof_n_<k>" with
 *                  k in sgs, residual, coef_uvw, coef_p, mip, grad, levels, pcg, sgs3 (u,v,w side-by-side passes);
 *                  "tuned_<r>" (chosen variant, -1 = not measured), "tuned_<r>_n", "tuned_<r>_cand<i>",
 *                  "tuned_<r>_ms<i>" with r in uvw, grad3, grad1, coef_p, mip, uvw_solve, rb_persistent;
 *                  "ncolors", "morton", "nlevels_natural", "nlevels_blocks", "ell_width", "num_sms". */
int cfdl_timer_record(cfdl_handle h, int32_t slot);                 /* slot 0..3 */
int cfdl_timer_elapsed_ms(cfdl_handle h, int32_t slot_begin, int32_t slot_end, double* ms);
int cfdl_host_alloc(void** ptr, uint64_t bytes);                    /* page-locked host memory */
int cfdl_host_free(void* ptr);
int cfdl_host_register(void* ptr, uint64_t bytes);                  /* page-lock caller-owned memory (Fortran allocatables) */
int cfdl_host_unregister(void* ptr);

/* ---- host <-> device field sync (for write_vtubin, main.f90:79,89) */
int cfdl_upload_field(cfdl_handle h, int field, const double* host);
int cfdl_download_field(cfdl_handle h, int field, double* host);

/* ---- whole-step path (device-resident; no host transfers besides the 16-double history) */
/* update_boundaries, mod_physics.f90:38-50 -> BC callbacks mod_uvwp.f90:493-570 */
int cfdl_update_boundaries(cfdl_handle h);
/* solve_uvwp, mod_uvwp.f90:95-134.  hist[4][4] (row = u,v,w,pc; col = it,res_i,res_f,res_max)
 * is what the reference prints per solve (mod_solver.f90:184,325); may be NULL. */
int cfdl_solve_uvwp(cfdl_handle h, double dt, int32_t nit, double* hist);
/* update_time, mod_physics.f90:101-112 */
int cfdl_update_time(cfdl_handle h);
/* The energy (enthalpy) and passive-scalar equations of the reference (single-GPU and partitioned handles:
 * tc / cp are given for every cell of the global mesh in the reference numbering, on every rank).
 *   cfdl_energy_init   construct_energy, mod_energy.f90:14-48: t = 273, phi = cp t, phi0 = phi, gradients 0; tc / cp are
 *                      per-cell arrays (ne, reference numbering) or NULL for init_properties' 5 and 1000
 *                      (mod_properties.f90:88-89).  From then on cfdl_update_boundaries also runs the energy callbacks
 *                      (lid: 373 K on the CFDL_BC_LID section, dirichlet0: 273 K elsewhere, mod_energy.f90:173-212) and
 *                      cfdl_update_time copies phi0 = phi (mod_physics.f90:110).
 *   cfdl_solve_energy  solve_energy, mod_energy.f90:59-80 (the call main.f90:59 has commented out): gradients of t and
 *                      phi, calc_coef_energy, solve_gs('e'), calc_temperature; out4 = (it, res_i, res_f, res_max).
 *   cfdl_scalar_init   construct_scalar, mod_scalar.f90:16-46, with this mesh's boundary sections: bc_value[i] is the
 *                      Dirichlet value on section i (dirichlet0 / dirichlet1; NULL: all 0); dcoef and vel are the
 *                      type's components (reference defaults 1 and (0,0,-100)).
 *   cfdl_solve_scalar  solve_scalar, mod_scalar.f90:56-72.
 * Fields: CFDL_F_T, CFDL_F_H (phi), CFDL_F_H0, CFDL_F_GT, CFDL_F_GH; CFDL_F_S, CFDL_F_S0, CFDL_F_GS.  The shared
 * ap / anb / b arrays hold the last assembled equation, as in the reference. */
int cfdl_energy_init(cfdl_handle h, const double* tc, const double* cp);
int cfdl_solve_energy(cfdl_handle h, double dt, int32_t nit, double* out4);
int cfdl_scalar_init(cfdl_handle h, double dcoef, const double* vel, const double* bc_value);
int cfdl_solve_scalar(cfdl_handle h, double dt, int32_t nit, double* out4);
/* write_vtubin + vtu_data, src/modules/mod_vtu_output.f90:6-326: the reference's binary .vtu file of one equation
 * (equation 0 = uvwp: u, v, w, p, gpc, 'mip'; 1 = energy: enthalpy, grad_enthalpy, temperature, grad_temperature; 2 = scalar:
 * phi, grad), byte for byte what the reference writes to <projPath>VTK/output-<iout>/<name>.vtu — `path` is that file
 * (its directory must exist).  The vertex coordinates and the element table are the arrays cell_input read from the mesh file
 * (geom%x,y,z; mg%esec(2,nsec), mg%etype(nsec), mg%e2vx(ne2vx_max, nelem), 1-based vertex ids).  For uvwp the reference's side
 * effect is kept: CFDL_F_DC is overwritten with the per-cell sum of the signed face mass fluxes (:114-123), which is what the
 * 'mip' array of the file holds.  Single-GPU handles. */
int cfdl_write_vtu(cfdl_handle h, const char* path, int32_t equation, int32_t nvx, const double* x, const double* y, const double* z,
                   int32_t nsec, const int32_t* esec, const int32_t* etype, int32_t ne2vx_max, const int32_t* e2vx);
/* the main.f90:50-63 loop: ntstep x (ncoef x (update_boundaries; solve_uvwp); update_time).
 * hist (ntstep*ncoef*16 doubles) may be NULL. */
int cfdl_run(cfdl_handle h, double dt, int32_t nit, int32_t ntstep, int32_t ncoef, double* hist);

/* One SIMPLE iteration for a driver whose HOST arrays stay authoritative (the reference's situation:
 * uvwp_t lives in Fortran-owned memory): [update_boundaries, mod_physics.f90:38-50, when apply_bcs != 0]
 * + solve_uvwp, mod_uvwp.f90:95-134, with the listed input fields (CFDL_F_* ids) uploaded first and
 * the listed output fields written back — the same as cfdl_upload_field x n_in, cfdl_update_boundaries,
 * cfdl_solve_uvwp, cfdl_download_field x n_out, except that transfers run beside the computation
 * where the iteration allows it: mip0 (first read by calc_mip) travels while the momentum equations
 * are assembled and solved; u, v, w and gu, gv, gw are final before the pressure-correction solve
 * (the reference's velocity correction is disabled, mod_uvwp.f90:387-391) and are copied back
 * during it.  Use page-locked host arrays (cfdl_host_alloc) for the overlap to take place.
 * local_numbering = 0: host arrays in the reference numbering (single-GPU handles);
 * local_numbering = 1: partition-local arrays as in cfdl_upload_field_local (any handle).
 * The call returns when every transfer has completed. */
int cfdl_step_host(cfdl_handle h, double dt, int32_t nit, int32_t apply_bcs, int32_t local_numbering,
                   int32_t n_in, const int32_t* in_fields, const double* const* in_ptrs,
                   int32_t n_out, const int32_t* out_fields, double* const* out_ptrs, double* hist);

/* Checkpoint / restart (the reference has none): the state that carries over between SIMPLE
 * iterations — u,v,w,p,u0,v0,w0,gu,gv,gw,gp,mip,mip0 — in the reference numbering on one GPU, as
 * one partition-local file per rank (path + ".r<rank>of<nranks>") on several.  A run continued from
 * a checkpoint reproduces the uninterrupted run bit for bit. */
int cfdl_checkpoint_write(cfdl_handle h, const char* path);
int cfdl_checkpoint_read(cfdl_handle h, const char* path);

/* ---- per-routine path on the handle's device-resident state (one reference routine each) */
int cfdl_calc_coef_uvw(cfdl_handle h, double dt);                 /* mod_uvwp.f90:161-286 */
int cfdl_calc_mip(cfdl_handle h, int32_t l_rhie_chow, double dt); /* mod_uvwp.f90:438-490 */
int cfdl_calc_coef_p(cfdl_handle h);                              /* mod_uvwp.f90:289-368 */
int cfdl_adjust_pc(cfdl_handle h);                                /* mod_uvwp.f90:129-130,136-158 */
int cfdl_update_uvwp(cfdl_handle h);                              /* mod_uvwp.f90:370-436 */
/* calc_grad, mod_solver.f90:40-81: phi in {U,V,W,P,PC} -> grad in {GU,GV,GW,GP,GPC} */
int cfdl_calc_grad(cfdl_handle h, int phi_field, int grad_field);
/* solve_gs / solve, mod_solver.f90:255-327,329-344 with the handle's ap/anb and the rhs of
 * equation eq (bu,bv,bw or b); out4 = it,res_i,res_f,res_max */
int cfdl_solve_eq(cfdl_handle h, int eq, int32_t nit, double* out4);

/* ---- stand-alone drop-ins with host arrays, argument order of the flat Fortran signatures.
 * The mesh comes from the handle (connectivity never changes between calls). */
/* calc_grad(phi,grad,...) mod_solver.f90:40 */
int cfdl_host_calc_grad(cfdl_handle h, const double* phi, double* grad);
/* solve_gs(cname,phi,ap,anb,b,...,nit) mod_solver.f90:255; eq selects omega */
int cfdl_host_solve_gs(cfdl_handle h, int eq, double* phi, const double* ap, const double* anb,
                       const double* b, int32_t nit, double* out4);
/* calc_residual(phi,ap,anb,b,...,res,res_max) mod_solver.f90:230 */
int cfdl_host_calc_residual(cfdl_handle h, const double* phi, const double* ap, const double* anb,
                            const double* b, double* res, double* res_max);
/* solve('pc',subdomain,intf,...) mod_solver.f90:329: multi_subdomain_solver when the handle
 * was created with n_subdomains>1, else solve_gs */
int cfdl_host_solve(cfdl_handle h, int eq, double* phi, const double* ap, const double* anb,
                    const double* b, int32_t nit, double* out4);

/* The assembly routines with host arrays: the reference passes derived types (uvwp_t, geometry_t,
 * properties_t); flattened here to the arrays each routine reads and writes, in the reference
 * numbering (cells+halos H, faces F, CSR slots Z, cells N).  Mesh, rho, mu and the BC table come
 * from the handle.  Single-GPU handles only. */
/* calc_coef_uvw(eqn,prop,geom,dt) mod_uvwp.f90:161-286: in u,v,w,u0,v0,w0 [H], gu,gv,gw,gp [3H],
 * mip [F]; out ap [N], anb [Z], bu,bv,bw,d,dc [N] */
int cfdl_host_calc_coef_uvw(cfdl_handle h, double dt, const double* u, const double* v, const double* w,
                            const double* u0, const double* v0, const double* w0,
                            const double* gu, const double* gv, const double* gw, const double* gp,
                            const double* mip, double* ap, double* anb,
                            double* bu, double* bv, double* bw, double* d, double* dc);
/* calc_mip(eqn,prop,geom,dt,l_rhie_chow) mod_uvwp.f90:438-490: in u,v,w,u0,v0,w0,p [H], gp [3H],
 * d [N], mip0 [F]; mip [F] in/out (only cell-cell faces are written, :456) */
int cfdl_host_calc_mip(cfdl_handle h, int32_t l_rhie_chow, double dt, const double* u, const double* v,
                       const double* w, const double* u0, const double* v0, const double* w0,
                       const double* p, const double* gp, const double* d, const double* mip0,
                       double* mip);
/* calc_coef_p(eqn,prop,geom) mod_uvwp.f90:289-368: in dc [N], mip [F]; out ap [N], anb [Z], b [N] */
int cfdl_host_calc_coef_p(cfdl_handle h, const double* dc, const double* mip,
                          double* ap, double* anb, double* b);
/* pref=phic(1); adjust_pc(phic,pref,...) mod_uvwp.f90:129-130,136-158: pc [H] in/out */
int cfdl_host_adjust_pc(cfdl_handle h, double* pc);
/* update_uvwp(eqn,prop,geom) mod_uvwp.f90:370-436: in pc [H], gpc [3H], dc [N]; p [H], gp [3H],
 * mip [F] in/out */
int cfdl_host_update_uvwp(cfdl_handle h, const double* pc, const double* gpc, const double* dc,
                          double* p, double* gp, double* mip);

/* ---- multi-GPU: one process per GPU (the reference is one process; launch P copies of the
 * driver, e.g. under mpirun/torchrun).  Every rank passes the same GLOBAL mesh, in the format
 * of cfdl_create, plus cell2rank(ne) in 1..nranks (e.g. from cfdl_partition_rcb) and its own
 * 0-based rank.  The library keeps the rank's cells, adds the adjacent cells of other ranks as
 * ghost cells and derives matching send/receive lists on every rank without communication.
 * Ghost values move by NCCL send/recv, residual norms by NCCL all-reduce, pc(1) by broadcast —
 * the GPU analogue of update_halos (mod_subdomains.f90:191-212) and of the residual loop of
 * multi_subdomain_solver (mod_solver.f90:144-150).  Solver: multicolour SGS with a global
 * colouring and a ghost refresh after every colour sweep, i.e. exactly the single-GPU MCSGS
 * iteration.  Field upload/download keep the global reference numbering: upload reads the
 * rank's entries of the global host array, download writes only the entries the rank owns. */
int cfdl_create_distributed(cfdl_handle* out, int32_t ne, int32_t nf, int32_t nbf,
                            const int32_t* ef2nb_idx, const int32_t* ef2nb_nb, const int32_t* ef2nb_fg,
                            const int32_t* s2g, const int32_t* bs,
                            const double* xc, const double* yc, const double* zc,
                            const double* aip, const double* rip, const double* vol,
                            const double* rho, const double* mu,
                            int32_t nbc, const int32_t* bc_esec, const int32_t* bc_kind, const double* bc_uvw,
                            const int32_t* cell2rank, int32_t rank, int32_t nranks, int32_t device);
/* The synthetic n^3 lid-driven cavity (BASELINE.json configs 2, 3 and 5) created without the
 * reference's packed (id<<5)|face int32 arrays, whose 2^26 id limit (ef2nb/bs/s2g
 * of mod_mg_lvl_uns.f90) stops at 406^3: the mesh cfdl_meshgen_fill(HEX, n, 0, 0) +
 * cfdl_mesh_build + cfdl_default cavity BCs would give (same cell/halo numbering, same
 * geometry bits, constant rho/mu, top section = lid (1,0,0)), generated analytically on each
 * rank and partitioned by recursive coordinate bisection (x, y, z in turn).  nranks must be a
 * power of two, n <= 700.  Differences from cfdl_create_distributed: host-side FACE fields
 * (CFDL_F_MIP, CFDL_F_MIP0) are numbered x-normal faces, then y, then z, not in the
 * reference's face order; everything else, including comm_init / ipc_connect, is the same. */
int cfdl_create_structured_hex(cfdl_handle* out, int32_t n, double rho, double mu,
                               int32_t rank, int32_t nranks, int32_t device);
/* The same cavity cut into nranks slabs along z (any nranks <= nz), the slowest index of the cell numbering: every
 * rank's interface cells are whole planes at the two ends of its own range.  nz = n (or 0) is the unit cube; any other
 * nz gives n x n x nz cells of the same edge 1/n, i.e. a cavity of depth nz/n with the lid on its top (z) face —
 * bench.py's weak-scaling series stacks one 128^3 block per GPU that way, so every GPU keeps the same work.  With such a partition the pc solve
 * runs as one persistent launch per rank whose chunks synchronise with the neighbouring chunks only, across
 * NVLink too (kernels_rbq.inc); with partitions that scatter the interface through the numbering (x slabs, octants)
 * the library falls back to one launch per pass.  Same results either way. */
int cfdl_create_structured_hex_slabs(cfdl_handle* out, int32_t n, int32_t nz, double rho, double mu,
                                     int32_t rank, int32_t nranks, int32_t device);
/* The arrays cfdl_create_structured_hex works from, for inspection and CPU tests (any output may
 * be NULL): nb/fg 6 n^3 slots (0-based neighbour cell, or n^3 + halo offset; signed 1-based face
 * id in the numbering described above), xc/yc/zc n^3 + 6 n^2, vol n^3, aip/rip 3 per face,
 * cell2rank n^3 entries in 1..nranks. */
int cfdl_structured_hex_arrays(int32_t n, int32_t nranks, int32_t* nb, int32_t* fg,
                               double* xc, double* yc, double* zc, double* vol,
                               double* aip, double* rip, int32_t* cell2rank);
/* 128-byte NCCL unique id, created on rank 0 and broadcast by the caller (MPI / torch) */
int cfdl_comm_unique_id(uint8_t id[128]);
/* collective over all ranks; must precede any compute call on a distributed handle */
int cfdl_comm_init(cfdl_handle h, const uint8_t id[128], int32_t rank, int32_t nranks);
/* optional peer-to-peer mode (GPUs of one node, NVLink): every rank exports the slab holding
 * u, v, w, pc (64-byte cudaIpcMemHandle), the caller gathers the nranks handles (rank order)
 * and every rank connects.  Afterwards the fused two-colour solver passes store interface values
 * straight into the neighbours' ghost cells and combine residual norms through peer memory —
 * no NCCL call inside a solver iteration.  set_option("p2p", 0) switches back to NCCL. */
int cfdl_comm_ipc_handle(cfdl_handle h, uint8_t handle[64]);
int cfdl_comm_ipc_connect(cfdl_handle h, const uint8_t* handles /* nranks * 64 bytes */);
/* partition-local host I/O (device numbering: owned cells, ghosts, halos): the bytes a rank
 * actually needs to move per step in a distributed host driver */
int cfdl_field_local_size(cfdl_handle h, int field, int64_t* n);
int cfdl_upload_field_local(cfdl_handle h, int field, const double* host);
int cfdl_download_field_local(cfdl_handle h, int field, double* host);
/* host-only view of the partition a rank would get (tests, tooling): owned and ghost cells in
 * device order (1-based global ids; arrays of capacity ne), neighbour ranks (capacity nranks),
 * send_ptr/recv_ptr (capacity nranks+1), send_cells (capacity ne), owned_color_ptr (capacity 33).
 * Any output pointer may be NULL. */
int cfdl_partition_plan(int32_t ne, int32_t nf, int32_t nbf, const int32_t* ef2nb_idx,
                        const int32_t* ef2nb_nb, const int32_t* ef2nb_fg, const int32_t* s2g,
                        const int32_t* bs, const double* xc, const double* yc, const double* zc,
                        const int32_t* cell2rank, int32_t nranks, int32_t rank,
                        int32_t* n_owned, int32_t* owned, int32_t* n_ghost, int32_t* ghost,
                        int32_t* n_nbr, int32_t* nbr_rank, int32_t* send_ptr, int32_t* send_cells,
                        int32_t* recv_ptr, int32_t* ncolors, int32_t* owned_color_ptr);

/* ---- host-side mesh tooling (no GPU needed) ------------------------------------------- */
enum { CFDL_MESH_HEX = 0, CFDL_MESH_TET = 1 };
/* synthetic unit-cube meshes standing in for the CGNS file read by cell_input.f90:36-99 */
int cfdl_meshgen_sizes(int kind, int n, int64_t* nvx, int64_t* ne, int64_t* nbf, int* nsec,
                       int* ne2vx_max);
int cfdl_meshgen_fill(int kind, int n, double jitter, int shuffle, uint64_t seed,
                      double* x, double* y, double* z, int32_t* e2vx, int32_t* etype,
                      int32_t* esec, char* names);
/* raw mesh file: what cell_input.f90:36-99 obtains from the CGNS library (vertex coordinates,
 * sections with name / element type / element range, element->vertex lists) in one little-endian
 * binary file, so that a driver can be built without CGNS/HDF5 — the Fortran reader is
 * cfd-lite_b200/fortran/mod_rawmesh.f90.  names: 32 characters per section; nelem = 3-D + 2-D
 * elements; e2vx has ne2vx_max entries per element as in mg_lvl%e2vx. */
int cfdl_rawmesh_write(const char* path, int64_t nvx, const double* x, const double* y, const double* z,
                       int nsec, const int32_t* etype, const int32_t* esec, const char* names,
                       int ne2vx_max, const int32_t* e2vx, int64_t nelem);
int cfdl_rawmesh_sizes(const char* path, int64_t* nvx, int64_t* nelem, int* nsec, int* ne2vx_max);
int cfdl_rawmesh_read(const char* path, double* x, double* y, double* z, int32_t* etype,
                      int32_t* esec, char* names, int32_t* e2vx);
/* fast connectivity + geometry build producing exactly the arrays of find_element_nb,
 * calc_aip_xyzip_uns and calc_vol_cv_centers_uns (mod_mg_lvl_uns.f90:283-488,
 * calc_aip_xyzip.f90, calc_vol_cv_centers.f90), including the reference's face numbering. */
int cfdl_mesh_build(int64_t nvx, const double* x, const double* y, const double* z,
                    int nsec, const int32_t* etype, const int32_t* esec, int ne2vx_max,
                    const int32_t* e2vx, int32_t ne, int32_t nf, int32_t nbf,
                    int32_t* ef2nb_idx, int32_t* ef2nb_nb, int32_t* ef2nb_fg, int32_t* s2g,
                    int32_t* bs, double* xc, double* yc, double* zc, double* aip, double* rip,
                    double* vol);
/* The same arrays computed on the GPU: face records bucketed by the smallest vertex of their face (atomics +
 * scan), one thread per vertex pairs the records of its bucket in (element, face) order — which is the reference's face
 * numbering order, so no sort is needed — then the geometry kernels with the reference's formulas.  Bit for bit the output of
 * cfdl_mesh_build; inputs and outputs are host arrays.  Fails with CFDL_ERR_CUDA without a device (no CPU path). */
int cfdl_mesh_build_gpu(int32_t device, int64_t nvx, const double* x, const double* y, const double* z,
                    int nsec, const int32_t* etype, const int32_t* esec, int ne2vx_max,
                    const int32_t* e2vx, int32_t ne, int32_t nf, int32_t nbf,
                    int32_t* ef2nb_idx, int32_t* ef2nb_nb, int32_t* ef2nb_fg, int32_t* s2g,
                    int32_t* bs, double* xc, double* yc, double* zc, double* aip, double* rip,
                    double* vol);

/* the reference's recursive-coordinate-bisection block decomposition (generate_seeds/grow/
 * split_leaf, mod_agglomeration.f90:380-561) and its block-local cell order (the unstable
 * quicksort of mod_mg_lvl_uns.f90:883-902): cell2sub(ne) in 1..P, g2gf_p(ne) cells sorted by
 * block, g2gf_idx(P+1); all 1-based, any output may be NULL. */
int cfdl_partition_rcb(int32_t ne, const double* xc, const double* yc, const double* zc,
                       const double* vol, int32_t P, int32_t* cell2sub, int32_t* g2gf_p,
                       int32_t* g2gf_idx);

#ifdef __cplusplus
}
#endif
#endif
