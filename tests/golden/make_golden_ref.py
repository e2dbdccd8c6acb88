#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: golden vectors from the REFERENCE'S OWN SOURCE TEXT.

No Fortran compiler exists in this image or on the GPU box (profiles/r02_fortran_probe_*.log; oracle/build_ref.sh is the
recipe for a box that has one), so the reference cannot be run as a binary.  This script executes the unmodified files
under /root/reference/src with oracle/f90run/f90py.py (statement-by-statement translation to Python, IEEE binary64,
evaluation order as written) and stores what the reference computes as fixtures:

    construct_physics (construct_subdomains, init_properties, construct_uvwp incl. make_bc and the initial calc_mip,
    construct_energy), then the loop of src/main.f90:50-63: update_boundaries + solve_uvwp (calc_coef_uvw, solve_gs x3,
    calc_grad x4, calc_mip, calc_coef_p, solve / multi_subdomain_solver, adjust_pc, update_uvwp) + update_time.

The mesh set-up is the reference's too: cell_input (src/setup/cell_input.f90) is followed statement by statement with
only its CGNS library calls (:36-93, cgnslib is not available) replaced by the arrays of the synthetic mesh file --
add_meshds / find_element_nb (connectivity), calc_aip_xyzip_uns, calc_vol_cv_centers_uns (geometry) and
add_transformation_bt (order of the cells inside the subdomains, incl. the unstable qsort_key_nRec) run from the reference's
source, and so does generate_seeds (mod_mg_lvl_uns.f90:95-117: which subdomain a cell belongs to -- the threaded binary tree
of mod_agglomeration.f90 over the doubly linked lists of mod_dll, with their defined assignments, structure-constructor
generics and functions that rebind their pointer dummies).  The fixtures therefore hold the reference's connectivity,
geometry, subdomain membership (setup_gf2g) and subdomain order.

Each fixture tests/golden/ref_<case>.npz holds the per-solve records the reference prints (name, it, res_i, res_f, res_max
-- taken from the argument list of its `write(*,oformat)` statements at full precision) and the final u, v, w, p, gp, mip,
pc, plus the intermediate matrices of the last iteration.  tests/test_oracle_vs_reference_source.py compares the C++ oracle
with them; this script needs /root/reference and is therefore not run by the test-suite.

usage: python tests/golden/make_golden_ref.py [case ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "f90run"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
import f90py  # noqa: E402

REF = os.environ.get("CFDL_REFERENCE", "/root/reference")
FILES = ["modules/mod_util.f90", "setup/mod_meshds_uns.f90", "setup/mod_agglomeration.f90", "setup/mod_mg_lvl_uns.f90",
         "setup/calc_aip_xyzip.f90", "setup/calc_vol_cv_centers.f90",
         "modules/mod_properties.f90", "modules/mod_eqn_setup.f90", "modules/mod_subdomains.f90", "modules/mod_solver.f90",
         "equations/mod_scalar.f90", "equations/mod_energy.f90", "modules/mod_multiphase.f90", "equations/mod_uvwp.f90",
         "modules/mod_physics.f90"]

# case name -> (mesh kind, n, jitter, shuffle, n_subdomains, ntstep, ncoef, boundary-condition overrides {section: routine}, dt)
# (dt = 0.01 is the reference's default, mod_physics.f90:15; the large-dt case makes the momentum solves iterate)
CASES = {
    "hex6_sub1": (0, 6, 0.0, False, 1, 1, 3, {}, 0.01),
    "hex6_sub4": (0, 6, 0.0, False, 4, 1, 3, {}, 0.01),
    "hex8_sub4": (0, 8, 0.0, False, 4, 2, 3, {}, 0.01),      # the reference's defaults: n_subdomains = 4, ncoef = 3
    "hex10_sub1": (0, 10, 0.0, False, 1, 2, 3, {}, 0.01),
    "hex5_jit_sub2": (0, 5, 0.15, False, 2, 1, 2, {}, 0.01),  # jittered vertices: non-orthogonal faces, deferred correction terms
    "hex6_dt5_sub1": (0, 6, 0.1, False, 1, 1, 3, {}, 5.0),
    "tet2_sub1": (1, 2, 0.2, True, 1, 1, 2, {}, 0.01),
    "tet3_sub2": (1, 3, 0.2, True, 2, 1, 3, {}, 0.01),
    "tet3_sub4": (1, 3, 0.2, True, 4, 1, 2, {}, 0.01),
    "tet4_sub4": (1, 4, 0.2, True, 4, 1, 2, {}, 0.01),        # 384 tets: uneven subdomains out of the bisection (threshold relaxation)
    "hex7_jit_dt5_sub4": (0, 7, 0.15, False, 4, 1, 3, {}, 5.0),  # block solver + iterating momentum solves on a non-orthogonal mesh
    "hex6_sub3": (0, 6, 0.0, False, 3, 1, 2, {}, 0.01),       # a subdomain count that is not a power of two
    "hex5_sym_sub1": (0, 5, 0.0, False, 1, 1, 2, {"east": "symmetry", "west": "symmetry"}, 0.01),
    "hex5_symjit_sub4": (0, 5, 0.15, False, 4, 1, 2, {"south": "symmetry", "bottom": "symmetry"}, 0.01),
    # the energy (enthalpy) and scalar equations next to uvwp: solve_energy is the call main.f90:59 has commented out,
    # solve_scalar the one at :56; the scalar's boundary values are 1 on 'top' and 'east', 0 elsewhere (dirichlet1 / dirichlet0)
    "hex6_energy_scalar_sub1": (0, 6, 0.1, False, 1, 1, 3, {}, 0.01),
    "tet3_energy_scalar_sub2": (1, 3, 0.2, True, 2, 1, 2, {}, 0.01),
}
EXTRA_EQUATIONS = ("hex6_energy_scalar_sub1", "tet3_energy_scalar_sub2")
SCALAR_ONES = ("top", "east")


def world():
    return f90py.World([os.path.join(REF, "src", f) for f in FILES])


def reference_cell_input(w, raw, n_subdomains, gf2g=None):
    """cell_input (src/setup/cell_input.f90:9-167) with the CGNS calls of :36-93 replaced by the raw mesh arrays; every
    `call` below is the reference's own routine executed from its source"""
    ns = w.ns
    g = ns["T_geometry_t"]()
    mg = g.mg
    nsec, nvx, ne, nbf, m = int(raw["nsec"]), int(raw["nvx"]), int(raw["ne"]), int(raw["nbf"]), int(raw["ne2vx_max"])
    names = raw["names"].decode()
    # :44-83 (cg_section_read_f, cgns_element_info, cgns_cell_data)
    mg.sectionname = np.array([names[32 * i:32 * (i + 1)].rstrip() for i in range(nsec)], dtype=object)
    mg.esec = np.asfortranarray(np.array(raw["esec"], dtype=np.int64).reshape(nsec, 2).T)
    mg.etype = np.array(raw["etype"], dtype=np.int64)
    mg.ne2vx = np.zeros(nsec, dtype=np.int64)
    element_nface, element_nvx = ns[w.pyname("element_nface")], ns[w.pyname("element_nvx")]
    cvs, vx2e_size = [0, 0, 0], 0
    for sec in range(nsec):
        cnt = int(mg.esec[1, sec] - mg.esec[0, sec] + 1)
        t = int(mg.etype[sec])
        if 10 <= t <= 20:
            cvs[0] += cnt
            cvs[1] += int(element_nface[t - 1]) * cnt
        else:
            cvs[2] += cnt
        vx2e_size += int(element_nvx[t - 1]) * cnt
    cvs[1] = (cvs[1] + cvs[2]) // 2
    assert cvs[0] == ne and cvs[2] == nbf
    nf, nelem = cvs[1], ne + nbf
    mg.e2vx = np.array(raw["e2vx"], dtype=np.int64)
    mg.nsec, mg.ne2vx_max, mg.nvx, mg.nelem, mg.nbndry, mg.nfaces = nsec, m, nvx, nelem, nbf, nf
    g.x, g.y, g.z = (np.array(raw[k], dtype=np.float64) for k in ("x", "y", "z"))
    # :99 onwards, verbatim
    w.get("add_meshds")(mg, nsec, g.x, g.y, g.z, vx2e_size)
    g.rip, g.aip = np.zeros(3 * nf), np.zeros(3 * nf)
    w.get("calc_aip_xyzip_uns")(g.x, g.y, g.z, g.rip, g.aip, mg.e2vx, mg.etype, nelem, nbf, nvx, m, mg.fine_lvl.s2g, nsec, mg.esec, nf)
    g.xc, g.yc, g.zc, g.vol = np.zeros(nelem), np.zeros(nelem), np.zeros(nelem), np.zeros(nelem - nbf)
    w.get("calc_vol_cv_centers_uns")(g.xc, g.yc, g.zc, g.vol, g.x, g.y, g.z, mg.e2vx, mg.fine_lvl.gs2nb, mg.fine_lvl.gs2nb_idx, mg.etype,
                                     nvx, nelem, nbf, nf, m, nsec, mg.esec, g.rip, g.aip)
    if n_subdomains > 1:
        # :131-153 ('Oct-Tree Isotropic', nl = (n_subdomains, 1)): generate_seeds -- the threaded-tree bisection of
        # mod_agglomeration.f90 over the doubly linked lists of mod_dll (mod_util.f90:2064-2290) -- decides which subdomain a
        # cell belongs to, add_transformation_bt (with the reference's unstable qsort_key_nRec) the order inside the
        # subdomains; both are executed from the reference's source.  (The second add_meshds builds the coarse level's
        # connectivity, which the hot path never reads.)
        nl = np.zeros(10, dtype=np.int64)
        nl[0], nl[1] = n_subdomains, 1
        w.get("generate_seeds")(mg, nl, 2, nelem - nbf, g.vol, g.xc, g.yc, g.zc)
        if gf2g is not None:  # (what the oracle's restatement says, for the caller's information only)
            g.gf2g_matches_oracle = bool(np.array_equal(np.asarray(mg.gf2g[0].p)[:nelem - nbf], np.asarray(gf2g)))
        w.get("add_transformation_bt")(mg)
    g.ef2nb, g.ef2nb_idx = mg.fine_lvl.gs2nb, mg.fine_lvl.gs2nb_idx
    g.ne, g.nf, g.nbf, g.nvx = int(mg.fine_lvl.ng), int(mg.fine_lvl.ns), int(mg.fine_lvl.nbs), nvx
    return g


def make_geom(w, raw, oc):
    """geometry_t (src/setup/mod_mg_lvl_uns.f90:46-58) filled with the set-up arrays in the reference's packed format"""
    ns = w.ns
    ne, nf, nbf = oc.ne, oc.nf, oc.nbf
    g = ns["T_geometry_t"]()
    g.ne, g.nf, g.nbf, g.nvx = ne, nf, nbf, int(raw["nvx"])
    for k in ("xc", "yc", "zc", "aip", "rip", "vol"):
        setattr(g, k, np.array(oc[k], dtype=np.float64))
    g.x, g.y, g.z = (np.array(raw[k], dtype=np.float64) for k in ("x", "y", "z"))
    g.ef2nb_idx = np.array(oc["ef2nb_idx"], dtype=np.int64)
    g.ef2nb = np.asfortranarray(np.stack([np.array(oc["ef2nb_nb"], dtype=np.int64), np.array(oc["ef2nb_fg"], dtype=np.int64)], axis=1))
    mg = g.mg
    fine = ns["T_meshds_t"]()
    fine.ng, fine.ns, fine.nbs = ne, nf, nbf
    fine.s2g = np.array(oc["s2g"], dtype=np.int64)
    bs = np.zeros(ne + nbf, dtype=np.int64)  # allocate(meshds%bs(ng+1:ng+nbs)), mod_meshds_uns.f90:55: indexed by the halo's element number
    bs[ne:] = oc["bs"]
    fine.bs = bs
    mg.fine_lvl = fine
    mg.cur_lvl = fine
    nsec = int(raw["nsec"])
    names = raw["names"].decode()
    mg.nsec = nsec
    mg.sectionname = np.array([names[32 * i:32 * (i + 1)].rstrip() for i in range(nsec)], dtype=object)
    mg.esec = np.asfortranarray(np.array(raw["esec"], dtype=np.int64).reshape(nsec, 2).T)
    mg.etype = np.array(raw["etype"], dtype=np.int64)
    bnd = [s + 1 for s in range(nsec) if mg.esec[0, s] > ne]  # 2-D sections = c2b interfaces, in file order (map_sec2intf)
    mg.nintf_c2b = len(bnd)
    mg.intf2sec = np.array(bnd, dtype=np.int64)
    mg.nelem, mg.nbndry, mg.nfaces, mg.nvx = ne + nbf, nbf, nf, int(raw["nvx"])
    if len(oc["g2gf_p"]):
        mg.g2gf[0].p = np.array(oc["g2gf_p"], dtype=np.int64)
        mg.g2gf[0].idx = np.array(oc["g2gf_idx"], dtype=np.int64)
        mg.gf2g[0].p = np.array(oc["gf2g"], dtype=np.int64)
    return g


def run_case(name, w=None, verbose=True):
    import cfdl
    import oracle
    kind, n, jitter, shuffle, nsub, ntstep, ncoef, bcs, dt = CASES[name]
    raw = cfdl.meshgen(kind, n, jitter=jitter, shuffle=shuffle, seed=12345)
    oc = oracle.OracleCase(raw, n_subdomains=nsub)
    w = w or world()
    ns = w.ns
    geom = reference_cell_input(w, raw, nsub, oc["gf2g"] if nsub > 1 else None)
    setup = dict(ef2nb_nb=geom.ef2nb[:, 0], ef2nb_fg=geom.ef2nb[:, 1], ef2nb_idx=geom.ef2nb_idx, s2g=geom.mg.fine_lvl.s2g[:geom.nf],
                 bs=geom.mg.fine_lvl.bs[geom.ne:geom.ne + geom.nbf], xc=geom.xc, yc=geom.yc, zc=geom.zc, aip=geom.aip, rip=geom.rip, vol=geom.vol)
    if nsub > 1:
        setup.update(g2gf_p=geom.mg.g2gf[0].p, g2gf_idx=geom.mg.g2gf[0].idx, gf2g=np.asarray(geom.mg.gf2g[0].p)[:geom.ne])
    phys = ns["T_phys_t"]()
    phys.n_subdomains = nsub
    phys.ntstep, phys.ncoef, phys.dt = ntstep, ncoef, dt
    t0 = time.time()
    # construct_energy (mod_energy.f90:33) forms eqn%t*prop%cp from arrays of ne+nbf and ne entries: a non-conforming
    # expression whose tail reads past prop%cp in a compiled run.  The energy equation is constructed but never solved
    # (main.f90:57) and nothing of it reaches uvwp, so for that one call cp is extended with its own constant.
    key = w.resolve("construct_energy")
    orig = w.get("construct_energy")

    def construct_energy_padded(g, mip, prop):
        cp = prop.cp
        prop.cp = np.concatenate([cp, np.full(g.nbf, cp[0])])
        try:
            return orig(g, mip, prop)
        finally:
            prop.cp = cp

    ns[key] = construct_energy_padded
    w.get("construct_physics")(phys, geom)
    ns[key] = orig
    for sec, routine in bcs.items():  # the reference hard-wires the cavity's BCs in construct_uvwp; other kinds by rebinding the pointer
        for bc in phys.uvwp.bcs:
            if bc.name.strip() == sec:
                bc.coef = w.get(routine, "mod_uvwp")
    del ns["_records"][:]
    update_boundaries, solve_uvwp, update_time = w.get("update_boundaries"), w.get("solve_uvwp"), w.get("update_time")
    extra = name in EXTRA_EQUATIONS
    if extra:
        solve_energy, solve_scalar = w.get("solve_energy"), w.get("solve_scalar")
        # construct_scalar (mod_scalar.f90:16-46) statement by statement, with this mesh's section names instead of the
        # hard-wired ones of another mesh ('fp-connect', 'out_wall', 'outlet')
        sc = ns["T_scalar_t"]()
        length = geom.ne + geom.nbf
        sc.name = "scalar"
        sc.phi, sc.phi0, sc.grad = np.zeros(length), np.zeros(length), np.zeros(3 * length)
        sc.bcs = f90py._newarr((geom.mg.nintf_c2b,), "type", ns["T_bc_t"])
        make_bc = w.get("make_bc")
        for sname in [x.strip() for x in geom.mg.sectionname[1:]]:
            bcp = make_bc(sc, geom, sname)
            bcp.coef = w.get("dirichlet1" if sname in SCALAR_ONES else "dirichlet0", "mod_scalar")
        phys.scalar = sc
    for tstep in range(ntstep):  # src/main.f90:50-63
        for icoef in range(ncoef):
            update_boundaries(phys, geom)
            if extra:  # mod_physics.f90:45, commented out there
                for bc in phys.scalar.bcs:
                    bc.coef(bc, geom, phys.scalar, phys.prop)
                solve_scalar(phys.scalar, phys.prop, geom, phys.dt, phys.nit, phys.ap, phys.anb, phys.b, phys.phic)  # main.f90:56
            solve_uvwp(phys.uvwp, phys.prop, geom, phys.dt, phys.nit, phys.ap, phys.anb, phys.b, phys.phic, phys.subdomain, phys.intf, phys.n_subdomains)
            if extra:
                solve_energy(phys.energy, phys.prop, geom, phys.dt, phys.nit, phys.ap, phys.anb, phys.b, phys.phic)  # main.f90:59
        update_time(phys)
        if extra:
            phys.scalar.phi0[...] = phys.scalar.phi  # mod_physics.f90:104, commented out there
    rec_all = [r for r in ns["_records"] if len(r) == 5]
    rec = [r for r in rec_all if r[0].strip() in ("u", "v", "w", "pc")]
    hist = np.array([[float(r[1]), float(r[2]), float(r[3]), float(r[4])] for r in rec]).reshape(ntstep * ncoef, 4, 4)
    names = [r[0].strip() for r in rec[:4]]
    assert names == ["u", "v", "w", "pc"], names
    e = phys.uvwp
    out = dict(hist=hist, u=e.u, v=e.v, w=e.w, p=e.p, gp=e.gp, gpc=e.gpc, gu=e.gu, gv=e.gv, gw=e.gw, mip=e.mip, mip0=e.mip0, u0=e.u0,
               d=e.d, dc=e.dc, bu=e.bu, bv=e.bv, bw=e.bw, ap=phys.ap, anb=phys.anb, b=phys.b, pc=phys.phic,
               case=np.array([kind, n, nsub, ntstep, ncoef]), dt=np.array(dt), jitter=np.array(jitter), shuffle=np.array(shuffle),
               bc_sections=np.array(list(bcs.keys()), dtype="U16"), bc_routines=np.array(list(bcs.values()), dtype="U16"))
    out.update({"setup_" + k: v for k, v in setup.items()})
    if extra:
        en = phys.energy
        out.update(hist_e=np.array([[float(x) for x in r[1:]] for r in rec_all if r[0].strip() == "e"]),
                   hist_s=np.array([[float(x) for x in r[1:]] for r in rec_all if r[0].strip() == "scalar"]),
                   t=en.t, gt=en.gt, h=en.phi, h0=en.phi0, gh=en.grad, s=sc.phi, s0=sc.phi0, gs=sc.grad,
                   scalar_ones=np.array(SCALAR_ONES, dtype="U16"))
    if verbose:
        print("%-14s ne=%d nf=%d: %d SIMPLE iterations of the reference source in %.1f s; iterations %s" %
              (name, oc.ne, oc.nf, ntstep * ncoef, time.time() - t0, hist[:, :, 0].astype(int).tolist()))
    return {k: np.array(v) for k, v in out.items()}


def main():
    cases = sys.argv[1:] or list(CASES)
    w = world()
    for c in cases:
        out = run_case(c, w)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % c), **out)


if __name__ == "__main__":
    main()
