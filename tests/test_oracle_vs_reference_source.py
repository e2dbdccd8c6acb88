"""The pin of the oracle to the reference: tests/golden/ref_*.npz hold what the reference's OWN SOURCE TEXT computes —
the unmodified files under /root/reference/src executed statement by statement by oracle/f90run/f90py.py
(tests/golden/make_golden_ref.py; no Fortran compiler exists here or on the GPU box, profiles/r02_fortran_probe_*.log).

CPU: the C++ oracle must reproduce those fixtures BIT FOR BIT — the set-up arrays of cell_input (connectivity of
find_element_nb, geometry of calc_aip_xyzip_uns / calc_vol_cv_centers_uns, subdomain membership of generate_seeds (the
threaded-tree bisection of mod_agglomeration.f90), subdomain order of add_transformation_bt), the
per-solve records (it, res_i, res_f, res_max) of solve_gs / multi_subdomain_solver and every field and matrix of solve_uvwp
(src/equations/mod_uvwp.f90:95-134) — on hex and tet meshes, orthogonal and jittered, 1 / 2 / 4 subdomains, wall / lid /
symmetry boundaries, small and large time steps.  The product's host-side set-up (cfdl_mesh_build, cfdl_partition_rcb) is
compared with the same set-up arrays.
GPU: the CUDA path through the C ABI against the same fixtures, parity mode, at north_star's 1e-10.
The fixtures travel with the repository; nothing here reads /root/reference."""
import glob
import os

import numpy as np
import pytest

from conftest import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(HERE, "golden", "ref_*.npz")))
FIELDS = ("u", "v", "w", "p", "gp", "gpc", "gu", "gv", "gw", "mip", "mip0", "u0", "d", "dc", "bu", "bv", "bw", "ap", "anb", "b")
SETUP = ("ef2nb_idx", "ef2nb_nb", "ef2nb_fg", "s2g", "bs", "xc", "yc", "zc", "aip", "rip", "vol")
BC_KIND = {"symmetry": 2, "lid": 1, "dirichlet0": 0}


def load(name):
    g = np.load(os.path.join(HERE, "golden", "ref_%s.npz" % name))
    kind, n, nsub, ntstep, ncoef = (int(x) for x in g["case"])
    return g, dict(kind=kind, n=n, nsub=nsub, ntstep=ntstep, ncoef=ncoef, dt=float(g["dt"]), jitter=float(g["jitter"]), shuffle=bool(g["shuffle"]),
                   bcs=dict(zip((str(s) for s in g["bc_sections"]), (str(r) for r in g["bc_routines"]))))


def section_index(raw, section):
    names = raw["names"].decode()
    secs = [names[32 * i:32 * (i + 1)].strip() for i in range(int(raw["nsec"]))]
    return secs.index(section) - 1  # boundary sections follow the volume section


def scalar_bc_values(raw, g):
    names = raw["names"].decode()
    secs = [names[32 * i:32 * (i + 1)].strip() for i in range(1, int(raw["nsec"]))]
    ones = set(str(x) for x in g["scalar_ones"])
    return np.array([1.0 if s in ones else 0.0 for s in secs])


def test_fixtures_present():
    assert len(FIXTURES) >= 10, FIXTURES


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_equals_reference_source_bit_for_bit(cfdl, oracle, name):
    g, kw = load(name)
    raw = cfdl.meshgen(kw["kind"], kw["n"], jitter=kw["jitter"], shuffle=kw["shuffle"], seed=12345)
    oc = oracle.OracleCase(raw, n_subdomains=kw["nsub"])
    oc.set_param("dt", kw["dt"])
    for sec, routine in kw["bcs"].items():
        oc.set_bc(section_index(raw, sec), BC_KIND[routine])
    for k in SETUP:
        assert np.array_equal(oc[k], g["setup_" + k]), "set-up array %s differs from the reference source" % k
    if kw["nsub"] > 1:
        assert np.array_equal(oc["gf2g"], g["setup_gf2g"]), "subdomain membership differs from the reference's generate_seeds"
        assert np.array_equal(oc["g2gf_p"], g["setup_g2gf_p"]) and np.array_equal(oc["g2gf_idx"], g["setup_g2gf_idx"])
    extra = "hist_e" in g.files  # energy (mod_energy.f90) and scalar (mod_scalar.f90) equations next to uvwp
    if extra:
        oc.construct_energy()
        oc.construct_scalar(bc_value=scalar_bc_values(raw, g))
        hist, he, hs = [], [], []
        for ts in range(kw["ntstep"]):
            for ic in range(kw["ncoef"]):
                oc.update_boundaries()
                hs.append(oc.solve_scalar())
                hist.append(oc.solve_uvwp())
                he.append(oc.solve_energy())
            oc.update_time()
        hist = np.array(hist)
        assert np.array_equal(np.array(he), g["hist_e"]), "energy: residual records differ from the reference source"
        assert np.array_equal(np.array(hs), g["hist_s"]), "scalar: residual records differ from the reference source"
        for f in ("t", "gt", "h", "h0", "gh", "s", "s0", "gs"):
            assert np.array_equal(oc[f], g[f]), "%s differs from the reference source: max rel %.3e" % (f, rel_err(oc[f], g[f]))
    else:
        hist, _ = oc.run(kw["ntstep"], kw["ncoef"])
    assert np.array_equal(hist[:, :, 0], g["hist"][:, :, 0]), "iteration counts differ from the reference source"
    assert np.array_equal(hist, g["hist"]), "residual history differs from the reference source: max rel %.3e" % rel_err(hist, g["hist"])
    for f in FIELDS:
        assert np.array_equal(oc[f], g[f]), "%s differs from the reference source: max rel %.3e" % (f, rel_err(oc[f], g[f]))
    assert np.array_equal(oc["phic"], g["pc"])


@pytest.mark.parametrize("name", FIXTURES)
def test_product_mesh_build_equals_reference_source(cfdl, name):
    """cfdl_mesh_build (sort-based face matching + geometry) and cfdl_partition_rcb's block order: no GPU, no oracle"""
    g, kw = load(name)
    raw = cfdl.meshgen(kw["kind"], kw["n"], jitter=kw["jitter"], shuffle=kw["shuffle"], seed=12345)
    geom = cfdl.mesh_build(raw)
    for k in SETUP:
        assert np.array_equal(np.asarray(geom[k]), g["setup_" + k]), "cfdl_mesh_build: %s differs from the reference source" % k
    if kw["nsub"] > 1:
        c2r, p, idx = cfdl.partition_rcb(geom, kw["nsub"])
        assert np.array_equal(c2r, g["setup_gf2g"]), "cfdl_partition_rcb: membership differs from the reference's generate_seeds"
        assert np.array_equal(p, g["setup_g2gf_p"]) and np.array_equal(idx, g["setup_g2gf_idx"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_equals_reference_source(cfdl, name):
    """No oracle involved: mesh from the product's own builder; subdomain order from the product's own RCB."""
    g, kw = load(name)
    raw = cfdl.meshgen(kw["kind"], kw["n"], jitter=kw["jitter"], shuffle=kw["shuffle"], seed=12345)
    geom = cfdl.mesh_build(raw)
    esec, kind, uvw = cfdl.default_bcs(raw)
    kind = np.array(kind)
    for sec, routine in kw["bcs"].items():
        kind[section_index(raw, sec)] = BC_KIND[routine]
    bcs = (esec, kind, uvw)
    if kw["nsub"] > 1:
        _, p, idx = cfdl.partition_rcb(geom, kw["nsub"])
        s = cfdl.Solver(geom, bcs, n_subdomains=kw["nsub"], g2gf_p=p, g2gf_idx=idx)
    else:
        s = cfdl.Solver(geom, bcs)
    try:
        s.set_option("solver", cfdl.SOLVER_PARITY)
        if "hist_e" in g.files:  # energy and scalar equations next to uvwp, in the order of the fixture's loop
            s.energy_init()
            s.scalar_init(bc_value=scalar_bc_values(raw, g))
            hist, he, hs = [], [], []
            for ts in range(kw["ntstep"]):
                for ic in range(kw["ncoef"]):
                    s.update_boundaries()
                    hs.append(s.solve_scalar(kw["dt"], 100))
                    hist.append(s.solve_uvwp(kw["dt"], 100))
                    he.append(s.solve_energy(kw["dt"], 100))
                s.update_time()
            hist, he, hs = np.array(hist), np.array(he), np.array(hs)
            assert np.array_equal(he[:, 0], g["hist_e"][:, 0]) and np.array_equal(hs[:, 0], g["hist_s"][:, 0])
            assert rel_err(he[:, 1:3], g["hist_e"][:, 1:3]) < 1e-10 and rel_err(hs[:, 1:3], g["hist_s"][:, 1:3]) < 1e-10
            for f in ("t", "gt", "h", "h0", "gh", "s", "s0", "gs"):
                assert rel_err(s.download(f), g[f]) < 1e-10, f
        else:
            hist = s.run(dt=kw["dt"], nit=100, ntstep=kw["ntstep"], ncoef=kw["ncoef"])
        assert np.array_equal(hist[:, :, 0], g["hist"][:, :, 0])
        # residual norms: summation order of the device reduction; everything else is the reference's arithmetic
        assert rel_err(hist[:, :, 1:3], g["hist"][:, :, 1:3]) < 1e-10
        for f in ("u", "v", "w", "p", "gp", "mip", "gu", "gv", "gw", "d", "dc"):
            assert rel_err(s.download(f)[:len(g[f])], g[f]) < 1e-10, f
    finally:
        s.close()
