"""GPU parity tests: every CUDA routine of the hot path, called through the C ABI
(include/cfdl.h via cfd-lite_b200/python/cfdl.py), against the CPU oracle on the same inputs.

Tolerances (FP64): north_star asks 1e-10 relative on residual history and final fields for
the box case.  Per-routine checks use 1e-12; because the kernels follow the reference's
operation order without FMA contraction most of them are in fact bit-identical, which the
tests report through `exact` asserts where that is guaranteed by construction.
"""
import numpy as np
import pytest

import conftest
from conftest import make_case, make_solver, rel_err, same_history

pytestmark = pytest.mark.gpu

TOL_ROUTINE = 1e-12
TOL_RUN = 1e-10

CASES = {
    "hex8_sub4": dict(kind=0, n=8, n_subdomains=4),
    "hex6_jitter": dict(kind=0, n=6, jitter=0.25, n_subdomains=1),
    "tet4_shuffled": dict(kind=1, n=4, jitter=0.2, shuffle=True, n_subdomains=1),
    "tet3_sub2": dict(kind=1, n=3, jitter=0.2, shuffle=True, n_subdomains=2),
    # symmetry planes (mod_uvwp.f90:517-546, bc_type 'zero_flux' in calc_coef_uvw :266-268) on two opposite and one
    # adjacent side of a jittered mesh (non-axis-aligned normals), walls + lid elsewhere
    "hex6_symmetry": dict(kind=0, n=6, jitter=0.2, n_subdomains=1, symmetry=(2, 3, 4)),
    "tet3_symmetry": dict(kind=1, n=3, jitter=0.2, shuffle=True, n_subdomains=1, symmetry=(0, 5)),
}


@pytest.fixture(scope="module", params=list(CASES))
def case(request, cfdl, oracle):
    raw, oc, geom = make_case(cfdl, oracle, **CASES[request.param])
    s = make_solver(cfdl, raw, oc, geom)
    yield request.param, raw, oc, geom, s
    s.close()


STATE = "u v w p u0 v0 w0 gu gv gw gp gpc mip mip0".split()


def randomize(oc, s, seed=7):
    """Same random state on both sides (halo entries included)."""
    rng = np.random.default_rng(seed)
    for name in STATE:
        a = oc[name]
        a[:] = rng.standard_normal(a.size) * (0.01 if name.startswith("mip") else 1.0)
        s.upload(name, a)
    pc = oc["phic"]
    pc[:] = rng.standard_normal(pc.size)
    s.upload("pc", pc)


def check(name, got, want, tol=TOL_ROUTINE):
    err = rel_err(got, want)
    assert err <= tol, "%s: rel err %.3e > %.1e" % (name, err, tol)
    return err


def test_update_boundaries(case):
    _, raw, oc, geom, s = case
    randomize(oc, s)
    oc.update_boundaries()
    s.update_boundaries()
    for f in ("u", "v", "w", "p", "mip"):
        assert np.array_equal(s.download(f), oc[f]), f


def test_calc_coef_uvw(case):
    _, raw, oc, geom, s = case
    randomize(oc, s, seed=11)
    oc.update_boundaries(); s.update_boundaries()
    oc.calc_coef_uvw()
    try:
        # statics 1: precomputed face statics + reciprocal quotients, locality order; statics 0: the reference's form — same bits
        for statics in (1, 0):
            s.set_option("statics", statics)
            for f in ("ap", "anb", "bu", "bv", "bw", "d", "dc"):
                s.upload(f, np.full(s.field_size(f), 7.5))  # stale values must be overwritten
            s.calc_coef_uvw(dt=0.01)
            for f in ("ap", "anb", "bu", "bv", "bw", "d", "dc"):
                got = s.download(f)
                check(f, got, oc[f])
                assert np.array_equal(got, oc[f]), "%s (statics %d) is not bit-identical: %.3e" % (f, statics, rel_err(got, oc[f]))
    finally:
        s.set_option("statics", 1)


def test_calc_grad(case):
    _, raw, oc, geom, s = case
    randomize(oc, s, seed=13)
    for phi, g in (("u", "gu"), ("v", "gv"), ("w", "gw"), ("p", "gp"), ("pc", "gpc")):
        want = oc.calc_grad(oc["phic" if phi == "pc" else phi])
        s.calc_grad(phi, g)
        got = s.download(g)
        check(g, got[:3 * oc.ne], want[:3 * oc.ne])
    # stand-alone drop-in with host arrays
    phi = np.random.default_rng(3).standard_normal(oc.ne + oc.nbf)
    check("host_calc_grad", s.host_calc_grad(phi)[:3 * oc.ne], oc.calc_grad(phi)[:3 * oc.ne])


def test_lsq_gradient_is_exact_for_linear_fields(case):
    """KAT lsq-linear (mod_solver.f90:48-79): exact gradient of a linear field on any mesh."""
    _, raw, oc, geom, s = case
    g = np.array([0.3, -1.7, 2.2])
    phi = 0.5 + g[0] * oc["xc"] + g[1] * oc["yc"] + g[2] * oc["zc"]
    got = s.host_calc_grad(phi)[:3 * oc.ne].reshape(-1, 3)
    assert np.abs(got - g).max() < 1e-9


@pytest.mark.parametrize("rhie_chow", [False, True])
def test_calc_mip(case, rhie_chow):
    _, raw, oc, geom, s = case
    randomize(oc, s, seed=17)
    oc.calc_coef_uvw(); s.calc_coef_uvw(dt=0.01)  # provides d
    mip_in = oc["mip"].copy()
    oc.calc_mip(rhie_chow)
    try:
        # precomputed face geometry (faces visited from the cells that number them, reciprocal quotients)
        # vs geometry recomputed in a per-face kernel as the reference does: same bits
        for statics in (1, 0):
            s.set_option("statics", statics)
            s.upload("mip", mip_in)  # every variant starts from the same faces (boundary ones are not written)
            s.calc_mip(rhie_chow, dt=0.01)
            got = s.download("mip")
            check("mip", got, oc["mip"])
            assert np.array_equal(got, oc["mip"]), "mip (statics=%d) not bit-identical: %.3e" % (statics, rel_err(got, oc["mip"]))
    finally:
        s.set_option("statics", 1)


def test_calc_coef_p(case):
    _, raw, oc, geom, s = case
    randomize(oc, s, seed=19)
    oc.update_boundaries(); s.update_boundaries()
    oc.calc_coef_uvw(); s.calc_coef_uvw(dt=0.01)
    oc.calc_coef_p()
    try:
        for statics in (0, 1):
            s.set_option("statics", statics)
            s.calc_coef_p()
            for f in ("ap", "anb", "b"):
                got = s.download(f)
                check(f, got, oc[f])
                assert np.array_equal(got, oc[f]), "%s (statics=%d) not bit-identical" % (f, statics)
    finally:
        s.set_option("statics", 1)
    # KAT pc-matrix: ap == sum anb (singular Neumann operator), anb symmetric
    ap, anb = s.download("ap"), s.download("anb")
    idx = geom["ef2nb_idx"] - 1
    rows = np.add.reduceat(anb, idx[:-1])
    assert np.abs(rows - ap).max() <= 1e-12 * np.abs(ap).max()


def test_adjust_pc_and_update_uvwp(case):
    _, raw, oc, geom, s = case
    for statics in (1, 0):
        randomize(oc, s, seed=23)
        s.set_option("statics", statics)
        try:
            oc.calc_coef_uvw(); s.calc_coef_uvw(dt=0.01)  # provides dc
            oc.adjust_pc(); s.adjust_pc()
            assert np.array_equal(s.download("pc"), oc["phic"])
            oc["gpc"][:] = oc.calc_grad(oc["phic"]); s.calc_grad("pc", "gpc")
            oc.update_uvwp(); s.update_uvwp()
            for f in ("p", "gp", "mip"):
                got = s.download(f)[:len(oc[f])]
                check(f, got, oc[f])
                if f == "mip":
                    assert np.array_equal(got, oc[f]), "mip (statics=%d) not bit-identical" % statics
        finally:
            s.set_option("statics", 1)


def assembled_system(oc):
    rng = np.random.default_rng(29)
    for name in STATE:
        a = oc[name]
        a[:] = rng.standard_normal(a.size) * (0.01 if name.startswith("mip") else 0.1)
    oc.update_boundaries()
    oc.calc_coef_uvw()
    return oc["ap"].copy(), oc["anb"].copy(), oc["bu"].copy(), oc["u"].copy()


def test_solve_gs_parity_mode(case, oracle):
    """Level-scheduled natural-order SGS == the sequential sweeps of solve_gs, bit for bit."""
    _, raw, oc, geom, s = case
    ap, anb, b, phi0 = assembled_system(oc)
    s.set_option("solver", 0)
    for eq, is_pc in ((0, False), (3, True)):
        want_phi, want = oracle.flat_solve_gs(is_pc, phi0, ap, anb, b, geom["ef2nb_idx"], geom["ef2nb_nb"], nit=7)
        got_phi, got = s.host_solve_gs(eq, phi0, ap, anb, b, nit=7)
        assert got[0] == want[0], (got, want)
        assert np.array_equal(got_phi[:oc.ne], want_phi[:oc.ne]), "phi differs by %.3e" % rel_err(got_phi[:oc.ne], want_phi[:oc.ne])
        check("res_i", got[1], want[1], 1e-12)
        check("res_f", got[2], want[2], 1e-12)
        if want[0] > 0:
            check("res_max", got[3], want[3], 1e-12)


def test_calc_residual(case, oracle):
    _, raw, oc, geom, s = case
    ap, anb, b, phi0 = assembled_system(oc)
    want = oracle.flat_calc_residual(phi0, ap, anb, b, geom["ef2nb_idx"], geom["ef2nb_nb"])
    got = s.host_calc_residual(phi0, ap, anb, b)
    check("res", got[0], want[0], 1e-12)
    check("res_max", got[1], want[1], 1e-12)


def test_solve_dispatcher_block_sgs(case):
    """solve('pc',...) == multi_subdomain_solver for n_subdomains>1 (incl. its residual quirks)."""
    name, raw, oc, geom, s = case
    rng = np.random.default_rng(31)
    for nm in STATE:
        a = oc[nm]
        a[:] = rng.standard_normal(a.size) * (0.01 if nm.startswith("mip") else 0.1)
    oc.update_boundaries(); oc.calc_coef_uvw(); oc.calc_coef_p()
    ap, anb, b = oc["ap"].copy(), oc["anb"].copy(), oc["b"].copy()
    phi0 = np.zeros(oc.ne + oc.nbf)
    s.set_option("solver", 0)
    want_phi, want = oc.solve(True, ap, anb, b, phi0, nit=20)
    got_phi, got = s.host_solve(3, phi0, ap, anb, b, nit=20)
    assert got[0] == want[0], (got, want)
    assert np.array_equal(got_phi[:oc.ne], want_phi[:oc.ne]), "phi differs by %.3e" % rel_err(got_phi[:oc.ne], want_phi[:oc.ne])
    for k in (1, 2, 3):
        check("stat%d" % k, got[k], want[k], 1e-12)


def test_mcsgs_equals_solve_gs_on_colour_permuted_system(case, oracle):
    """Fast mode = the reference's solve_gs applied to the colour-major renumbered system."""
    _, raw, oc, geom, s = case
    ap, anb, b, phi0 = assembled_system(oc)
    c2o, cp = s.cell_order()
    ne = oc.ne
    o2c = np.zeros(ne, np.int64)
    o2c[c2o - 1] = np.arange(ne)
    idx = geom["ef2nb_idx"].astype(np.int64) - 1
    nb_packed = geom["ef2nb_nb"].astype(np.int64)
    lens = np.diff(idx)
    p_idx = np.concatenate([[0], np.cumsum(lens[c2o - 1])])
    p_nb = np.zeros_like(nb_packed)
    p_anb = np.zeros_like(anb)
    for c in range(ne):
        e = c2o[c] - 1
        sl = slice(idx[e], idx[e + 1])
        nbid = nb_packed[sl] >> 5
        lf = nb_packed[sl] & 31
        new = np.where(lf > 0, o2c[np.minimum(nbid, ne) - 1] + 1, nbid)
        p_nb[p_idx[c]:p_idx[c + 1]] = (new << 5) | lf
        p_anb[p_idx[c]:p_idx[c + 1]] = anb[sl]
    perm_phi = np.concatenate([phi0[c2o - 1], phi0[ne:]])
    s.set_option("solver", 1)
    try:
        # fused=1: two-colour meshes run the fused red/black passes (two updates + residual from one
        # evaluation of the neighbour sum); fused=0: one launch per colour + a residual pass.  Both
        # must equal the sequential reference sweeps on the permuted system bit for bit.
        for fused in (1, 0):
            s.set_option("fused", fused)
            for eq, is_pc in ((0, False), (3, True)):
                for nit in (1, 2, 9):
                    want_phi, want = oracle.flat_solve_gs(is_pc, perm_phi, ap[c2o - 1], p_anb, b[c2o - 1], (p_idx + 1).astype(np.int32),
                                                          p_nb.astype(np.int32), nit=nit)
                    got_phi, got = s.host_solve_gs(eq, phi0, ap, anb, b, nit=nit)
                    assert got[0] == want[0], (fused, eq, nit, got, want)
                    assert np.array_equal(got_phi[c2o - 1], want_phi[:ne]), "phi differs by %.3e" % rel_err(got_phi[c2o - 1], want_phi[:ne])
                    check("res_i", got[1], want[1], 1e-12)
                    check("res_f", got[2], want[2], 1e-12)
                    if want[0] > 0:
                        check("res_max", got[3], want[3], 1e-12)
    finally:
        s.set_option("fused", 1)
        s.set_option("solver", 0)


def test_run_parity(case):
    """main.f90:50-63 schedule (2 steps x 3 iterations): history and fields within 1e-10."""
    name, raw, oc, geom, s0 = case
    from conftest import make_case as mk
    import cfdl as cf
    import oracle as orc
    raw, oc, geom = mk(cf, orc, **CASES[name])
    s = make_solver(cf, raw, oc, geom)
    try:
        want_hist, _ = oc.run(2, 3)
        got_hist = s.run(dt=0.01, nit=100, ntstep=2, ncoef=3)
        assert np.array_equal(got_hist[:, :, 0], want_hist[:, :, 0]), "iteration counts differ"
        for k, nm in ((1, "res_i"), (2, "res_f")):
            assert rel_err(got_hist[:, :, k], want_hist[:, :, k]) <= TOL_RUN, nm
        ran = want_hist[:, :, 0] > 0  # res_max is undefined in the reference when no iteration ran
        assert rel_err(got_hist[:, :, 3][ran], want_hist[:, :, 3][ran]) <= TOL_RUN
        for f in ("u", "v", "w", "p", "mip"):
            check(f, s.download(f), oc[f], TOL_RUN)
    finally:
        s.close()


def test_box_standin_full_schedule(cfdl, oracle):
    """Config #1: box stand-in (32^3 hex), reference defaults: n_subdomains=4, dt=0.01,
    nit=100, 10 time steps x 3 iterations.  Residual history and final u,v,w,p to 1e-10."""
    raw, oc, geom = make_case(cfdl, oracle, kind=0, n=32, n_subdomains=4)
    s = make_solver(cfdl, raw, oc, geom)
    try:
        ntstep = 3 if conftest.EMULATED else 10  # the host emulation runs the first 9 of the 30 iterations
        want_hist, _ = oc.run(ntstep, 3)
        got_hist = s.run(dt=0.01, nit=100, ntstep=ntstep, ncoef=3)
        assert np.array_equal(got_hist[:, :, 0], want_hist[:, :, 0])
        assert rel_err(got_hist[:, :, 1:3], want_hist[:, :, 1:3]) <= TOL_RUN
        for f in ("u", "v", "w", "p"):
            check(f, s.download(f), oc[f], TOL_RUN)
    finally:
        s.close()


def test_create_rejects_bad_mesh(cfdl, oracle):
    raw, oc, geom = make_case(cfdl, oracle, kind=0, n=3)
    bad = dict(geom)
    bad["ef2nb_fg"] = geom["ef2nb_fg"].copy()
    bad["ef2nb_fg"][0] = 0
    with pytest.raises(cfdl.CfdlError):
        cfdl.Solver(bad, oc.bc_table())


def test_overlapped_passes_equal_serialised_passes(cfdl):
    """Fused passes launched with programmatic dependent launch (a pass starts while the previous
    one drains, option pdl=1) against the same passes fully serialised (pdl=0) and against one
    launch per colour (fused=0): identical bits over ~100-iteration pc solves on a mesh large
    enough (64^3) for every SM to hold several CTAs of consecutive passes at once."""
    raw = cfdl.meshgen(cfdl.MESH_HEX, 28 if conftest.EMULATED else 64)  # (launch overlap does not exist under emulation)
    geom = cfdl.mesh_build(raw)
    out = []
    for opts in ({"pdl": 1}, {"pdl": 0, "rbq": 0}, {"pdl": 1, "rbq": 0}, {"fused": 0}):
        s = cfdl.Solver(geom, cfdl.default_bcs(raw))
        s.set_option("solver", cfdl.SOLVER_MCSGS)
        for k, v in opts.items():
            s.set_option(k, v)
        hist = s.run(dt=0.01, nit=100, ntstep=2, ncoef=2)
        out.append((hist, {f: s.download(f) for f in ("u", "v", "w", "p", "pc", "mip")}))
        s.close()
    assert out[0][0][:, 3, 0].max() >= (30 if conftest.EMULATED else 50), "pc solves too short to exercise the overlap"
    for hist, fields in out[1:]:
        assert np.array_equal(hist[:, :, 0], out[0][0][:, :, 0])
        for f in fields:
            assert np.array_equal(fields[f], out[0][1][f]), f


@pytest.mark.parametrize("solver", ["parity", "mcsgs"])
def test_step_host_equals_separate_calls(case, cfdl, solver):
    """cfdl_step_host (uploads, update_boundaries, solve_uvwp, downloads in one call, transfers beside
    the computation) against the same sequence of separate C-ABI calls: identical bits."""
    _, raw, oc, geom, s = case
    ins = "u v w p u0 v0 w0 gu gv gw gp mip mip0".split()
    outs = "u v w p gu gv gw gp gpc mip".split()
    rng = np.random.default_rng(23)
    state = {k: rng.standard_normal(s.field_size(k)) * (0.01 if k.startswith("mip") else 0.1) for k in ins}
    s.set_option("solver", cfdl.SOLVER_PARITY if solver == "parity" else cfdl.SOLVER_MCSGS)
    try:
        for k in ins:
            s.upload(k, state[k])
        s.update_boundaries()
        want_hist = s.solve_uvwp(0.01, 20)
        want = {k: s.download(k) for k in outs}
        # scramble the device state so that the second path must really transfer everything
        for k in ins:
            s.upload(k, np.full(s.field_size(k), 3.25))
        bufs_in = {k: cfdl.PinnedBuffer(s.field_size(k)) for k in ins}
        bufs_out = {k: cfdl.PinnedBuffer(s.field_size(k)) for k in outs}
        for k in ins:
            bufs_in[k].array[:] = state[k]
        got_hist = s.step_host({k: bufs_in[k].array for k in ins}, {k: bufs_out[k].array for k in outs}, dt=0.01, nit=20)
        assert same_history(got_hist, want_hist)
        for k in outs:
            assert np.array_equal(bufs_out[k].array, want[k]), k
        # a second call on the same handle reuses the transfer streams and staging area
        got_hist2 = s.step_host({k: bufs_in[k].array for k in ins}, {k: bufs_out[k].array for k in outs}, dt=0.01, nit=20)
        assert same_history(got_hist2, want_hist)
        for k in outs:
            assert np.array_equal(bufs_out[k].array, want[k]), k
        # fields outside the usual lists, incl. anb (CSR <-> ELL: the unstaged path)
        extra = {k: np.zeros(s.field_size(k)) for k in ("anb", "ap", "dc", "bu")}
        outs3 = {k: bufs_out[k].array for k in outs}
        outs3.update(extra)
        s.step_host({k: bufs_in[k].array for k in ins}, outs3, dt=0.01, nit=20)
        for k in extra:
            assert np.array_equal(extra[k], s.download(k)), k
        for b in list(bufs_in.values()) + list(bufs_out.values()):
            b.free()
    finally:
        s.set_option("solver", cfdl.SOLVER_PARITY)


def test_calc_grad_variants_keep_the_bits(case):
    """grad_variant 1 (inverse least-squares matrix and weights precomputed once with the reference's
    expressions) against variant 0 (everything rebuilt per call, as the reference does): same bits, for
    the single-field kernel and for the fused u,v,w pass inside solve_uvwp."""
    _, raw, oc, geom, s = case
    randomize(oc, s, seed=29)
    got = {}
    try:
        for variant in (0, 1):
            s.set_option("grad_variant", variant)
            s.calc_grad("p", "gp")
            got[variant, "gp"] = s.download("gp")[:3 * oc.ne]
            s.calc_grad("pc", "gpc")
            got[variant, "gpc"] = s.download("gpc")[:3 * oc.ne]
        want = oc.calc_grad(oc["p"])[:3 * oc.ne]
        check("gp", got[0, "gp"], want)
        for f in ("gp", "gpc"):
            for variant in (1,):
                assert np.array_equal(got[0, f], got[variant, f]), (f, variant)
        # the fused three-field pass runs inside solve_uvwp: two iterations from the same state
        hist = {}
        for variant in (0, 1):
            s.set_option("grad_variant", variant)
            randomize(oc, s, seed=31)
            s.update_boundaries()
            hist[variant] = s.solve_uvwp(0.01, 5)
            for f in ("gu", "gv", "gw", "gp", "mip", "p"):
                got[variant, f] = s.download(f)
        for variant in (1,):
            assert same_history(hist[0], hist[variant])
            for f in ("gu", "gv", "gw", "gp", "mip", "p"):
                assert np.array_equal(got[0, f], got[variant, f]), (f, variant)
    finally:
        s.set_option("grad_variant", 1)


def test_calc_coef_p_overwrites_stale_values(case):
    """calc_coef_p on the face statics (paired colour order on two-colour meshes, locality order otherwise,
    quotient by dr.n through the stored reciprocal) and in the reference's form: every entry of ap, anb, b is
    rewritten and equals the oracle's bits."""
    _, raw, oc, geom, s = case
    randomize(oc, s, seed=37)
    oc.update_boundaries(); s.update_boundaries()
    oc.calc_coef_uvw(); s.calc_coef_uvw(dt=0.01)  # provides dc
    oc.calc_coef_p()
    try:
        for statics in (1, 0):
            s.set_option("statics", statics)
            for f in ("ap", "anb", "b"):
                s.upload(f, np.full(s.field_size(f), 7.5))  # stale values must be overwritten
            s.calc_coef_p()
            for f in ("ap", "anb", "b"):
                assert np.array_equal(s.download(f), oc[f]), (statics, f)
    finally:
        s.set_option("statics", 1)


def _numpy_pcg(ne, idx, nb_packed, ap, anb, b, phi0, nit, red=None):
    """Preconditioned CG with the reference's residual definition and stopping rule, in plain numpy on the
    reference-format (CSR, packed neighbour ids) system: the model of solver=pcg.  red = None: Jacobi; red = boolean
    mask of the first colour: two-colour symmetric Gauss-Seidel, z = (D - U)^-1 D (D - L)^-1 r with red before black."""
    rows = np.repeat(np.arange(ne), np.diff(idx))
    cols = (nb_packed >> 5) - 1  # cells and halos share the id space of phi
    x = phi0.copy()

    def nbsum(v):
        return np.bincount(rows, weights=anb * v[cols], minlength=ne)

    def precond(r):
        if red is None:
            return r / ap
        z = np.zeros_like(x)
        z[:ne][red] = (r / ap)[red]                       # y_r
        zb = (r + nbsum(z)) / ap                           # black rows gather the red y
        z[:ne][~red] = zb[~red]
        zr = z[:ne] + nbsum(z) / ap                        # red rows gather the black z
        z[:ne][red] = zr[red]
        return z[:ne].copy()

    r = b + nbsum(x) - ap * x[:ne]
    res_i = np.sqrt(np.sum(r * r) / ne)
    res_f, res_max, it = res_i, 0.0, 0
    p = np.zeros_like(x)
    z = precond(r)
    p[:ne] = z
    rz = np.sum(r * z)
    while it < nit and res_f > res_i / 10.0:
        q = ap * p[:ne] - nbsum(p)
        pq = np.sum(p[:ne] * q)
        if not pq > 0.0:
            break
        alpha = rz / pq
        x[:ne] += alpha * p[:ne]
        r = r - alpha * q
        it += 1
        res_f, res_max = np.sqrt(np.sum(r * r) / ne), np.abs(r).max()
        z = precond(r)
        rz_new = np.sum(r * z)
        p[:ne] = z + (rz_new / rz) * p[:ne]
        rz = rz_new
    return x, (it, res_i, res_f, res_max)


def test_pcg_solves_the_pc_system_like_its_numpy_model(case, cfdl):
    """solver=pcg (conjugate gradients for pc; not a reference algorithm, see kernels_pcg.inc): same
    iterates as the numpy model up to summation order (1e-9), same iteration counts and stopping rule,
    and the residual really drops by the factor the reference asks for."""
    name, raw, oc, geom, s = case
    rng = np.random.default_rng(41)
    for nm in STATE:
        a = oc[nm]
        a[:] = rng.standard_normal(a.size) * (0.01 if nm.startswith("mip") else 0.1)
    oc.update_boundaries(); oc.calc_coef_uvw(); oc.calc_mip(True); oc.calc_coef_p()
    ap, anb, b = oc["ap"].copy(), oc["anb"].copy(), oc["b"].copy()
    b -= b.mean()  # consistent right-hand side for the singular Neumann system (true in a run: boundary fluxes are zero)
    phi0 = np.zeros(oc.ne + oc.nbf)
    idx = geom["ef2nb_idx"].astype(np.int64) - 1
    nbp = geom["ef2nb_nb"].astype(np.int64)
    s.set_option("solver", cfdl.SOLVER_PCG)
    # the library's first colour on a two-colour mesh: the cells at even graph distance from cell 1 (greedy colouring in natural order)
    red = None
    if int(s.get_info("ncolors")) == 2:
        rows = np.repeat(np.arange(oc.ne), np.diff(idx))
        cols = (nbp >> 5) - 1
        dist = np.full(oc.ne, -1)
        dist[0], frontier = 0, [0]
        while frontier:
            nxt = []
            for c in frontier:
                for j in cols[idx[c]:idx[c + 1]]:
                    if j < oc.ne and dist[j] < 0:
                        dist[j] = dist[c] + 1
                        nxt.append(j)
            frontier = nxt
        red = dist % 2 == 0
    try:
        iters = {}
        for precond in ((1, 0) if red is not None else (0,)):
            s.set_option("pcg_precond", precond)
            for nit in (1, 3, 12, 400):
                want_phi, want = _numpy_pcg(oc.ne, idx, nbp, ap, anb, b, phi0, nit, red if precond else None)
                got_phi, got = s.host_solve(3, phi0, ap, anb, b, nit=nit)
                assert got[0] == want[0], (precond, nit, got, want)
                scale = max(np.abs(want_phi[:oc.ne]).max(), 1e-300)
                assert np.abs(got_phi[:oc.ne] - want_phi[:oc.ne]).max() / scale < 1e-9, (precond, nit)
                for k in (1, 2, 3):
                    if want[k] != 0.0:
                        assert abs(got[k] - want[k]) <= 1e-9 * abs(want[k]), (precond, nit, k, got, want)
            assert got[2] <= got[1] / 10.0 and got[0] < 400  # converged by the reference's criterion, not by the cap
            iters[precond] = got[0]
        if red is not None:
            assert iters[1] <= iters[0], iters  # the Gauss-Seidel preconditioner never needs more iterations than Jacobi
        s.set_option("pcg_precond", 1)
        # momentum equations are not symmetric: solver=pcg keeps them on MCSGS
        s.set_option("solver", cfdl.SOLVER_MCSGS)
        ap_u, anb_u, b_u, phi_u = assembled_system(oc)
        ref_phi, ref = s.host_solve_gs(0, phi_u, ap_u, anb_u, b_u, nit=5)
        s.set_option("solver", cfdl.SOLVER_PCG)
        pcg_phi, pcg = s.host_solve_gs(0, phi_u, ap_u, anb_u, b_u, nit=5)
        assert np.array_equal(ref_phi, pcg_phi) and np.array_equal(ref, pcg)
    finally:
        s.set_option("solver", cfdl.SOLVER_PARITY)


def test_momentum_solves_side_by_side_keep_the_bits(case, cfdl):
    """uvw_fused=1 (u, v, w in one set of passes — fused two-colour passes, or one launch per colour on
    meshes with more colours — matrix rows read once per pass, kernels_rb3.inc) against uvw_fused=0 (three separate solves as the reference orders them): every
    equation stops at its own iteration count and all fields are bit-identical."""
    name, raw, oc, geom, s = case
    s.set_option("solver", cfdl.SOLVER_MCSGS)
    try:
        for seed, nit in ((43, 100), (47, 2), (53, 1)):
            res = {}
            for fused in (0, 1):
                s.set_option("uvw_fused", fused)
                randomize(oc, s, seed=seed)
                s.update_boundaries()
                h1 = s.solve_uvwp(0.01, nit)
                s.update_boundaries()
                h2 = s.solve_uvwp(0.01, nit)
                res[fused] = (h1, h2, {f: s.download(f) for f in ("u", "v", "w", "p", "gu", "gv", "gw", "mip")})
            for i in (0, 1):
                a, b = res[0][i], res[1][i]
                assert np.array_equal(a[:, 0], b[:, 0]), (seed, nit, a[:, 0], b[:, 0])  # iteration counts
                assert np.array_equal(a[:, 1], b[:, 1])                                  # opening residuals: same launch geometry
                assert np.allclose(a[:, 2:], b[:, 2:], rtol=1e-12, atol=0.0)
            for f, v in res[0][2].items():
                assert np.array_equal(v, res[1][2][f]), (seed, nit, f)
    finally:
        s.set_option("uvw_fused", 1)
        s.set_option("solver", cfdl.SOLVER_PARITY)


def test_pc_passes_rebuilding_the_diagonal_keep_the_bits(case, cfdl):
    """pc_sumap=1: the fused pc passes do not read ap but rebuild it as the slot-order sum of the row's
    coefficients — how calc_coef_p forms it — against pc_sumap=0: identical history and fields; a
    caller-supplied matrix (host drop-in) whose diagonal is NOT that sum must still be honoured."""
    name, raw, oc, geom, s = case
    if int(s.get_info("ncolors")) != 2:
        pytest.skip("fused two-colour passes only")
    s.set_option("solver", cfdl.SOLVER_MCSGS)
    try:
        res = {}
        for flag in (0, 1):
            s.set_option("pc_sumap", flag)
            randomize(oc, s, seed=59)
            hs = []
            for _ in range(2):
                s.update_boundaries()
                hs.append(s.solve_uvwp(0.01, 30))
            res[flag] = (np.array(hs), {f: s.download(f) for f in ("u", "p", "pc", "gpc", "mip")})
        assert same_history(res[0][0], res[1][0])
        for f, v in res[0][1].items():
            assert np.array_equal(v, res[1][1][f]), f
        # host drop-in with a diagonally dominant matrix (ap != sum anb): the flag must not apply
        s.set_option("pc_sumap", 1)
        s.update_boundaries(); s.solve_uvwp(0.01, 3)  # leaves the device matrix in the "written by calc_coef_p" state
        ap, anb, b, phi0 = assembled_system(oc)
        want_phi, want = None, None
        for flag in (0, 1):
            s.set_option("pc_sumap", flag)
            got_phi, got = s.host_solve_gs(3, phi0, ap, anb, b, nit=4)
            if want_phi is None:
                want_phi, want = got_phi, got
        assert np.array_equal(got_phi, want_phi) and np.array_equal(got, want)
    finally:
        s.set_option("pc_sumap", 1)
        s.set_option("solver", cfdl.SOLVER_PARITY)


def test_persistent_pc_solve_keeps_the_bits(case, cfdl):
    """The pc solve as ONE persistent launch with neighbour-only synchronisation (kernels_rbq.inc: red values as
    {newest, mid} pairs, two rows per thread, 128-bit loads, a fixed block of iterations per launch with the
    stopping rule applied to the recorded residuals, a block redone when the rule fired inside it) against the
    pass-by-pass kernels, with 16-bit neighbour offsets and with 32-bit ids: same fields, same iteration
    counts; the RMS norms are summed per chunk instead of per grid-stride CTA and agree to rounding.  The
    iteration caps 1, 2, 7 exercise the estimate-too-short / rule-fires-inside-the-block paths, 40 and 100 the
    long solves; several CTAs per SM are forced so that even the small meshes are cut into many chunks."""
    name, raw, oc, geom, s = case
    if int(s.get_info("ncolors")) != 2:
        pytest.skip("fused two-colour passes only")
    s.set_option("solver", cfdl.SOLVER_MCSGS)
    try:
        res = {}
        # third entry: more chunks than CTAs (chunks of 64 rows handed out in order from a counter to at most 2 CTAs)
        combos = [(0, 1, 0), (1, 1, 0), (1, 0, 0), (0, 0, 0), (1, 1, 1), (1, 0, 1)]
        for combo in combos:
            s.set_option("rbq", combo[0])
            s.set_option("rb_idx16", combo[1])
            s.set_option("rbq_lmax", 64 if combo[2] else 0)
            s.set_option("rbq_lbig", 64 if combo[2] else 0)
            s.set_option("rbq_cap", 2 if combo[2] else 0)
            randomize(oc, s, seed=61)
            hs = []
            for nit in (1, 2, 7, 40, 100, 3):
                s.update_boundaries()
                hs.append(s.solve_uvwp(0.01, nit))
            res[combo] = (np.array(hs), {f: s.download(f) for f in ("u", "v", "w", "p", "pc", "mip")})
            if combo[0]:
                assert int(s.get_info("rbq_active")) == 1 and int(s.get_info("rbq_refused")) == 0
                if combo[2] and s.ne >= 512:
                    assert int(s.get_info("rbq_grid")) == 2 and int(s.get_info("rbq_chunks")) > 2 and int(s.get_info("rbq_chunk_rows")) == 64
        ref_hist, ref_fields = res[combos[0]]
        for combo, (hist, fields) in res.items():
            assert same_history(hist, ref_hist), (combo, hist, ref_hist)
            for f, v in fields.items():
                assert np.array_equal(v, ref_fields[f]), (combo, f)
    finally:
        s.set_option("rbq", 1)
        s.set_option("rb_idx16", 1)
        for k in ("rbq_lmax", "rbq_lbig", "rbq_cap"):
            s.set_option(k, 0)
        s.set_option("solver", cfdl.SOLVER_PARITY)


@pytest.mark.parametrize("solver", ["parity", "mcsgs"])
def test_restart_from_checkpoint_reproduces_the_run(case, cfdl, tmp_path, solver):
    """cfdl_checkpoint_write / _read: 2 time steps + checkpoint + 2 more == fresh handle + read + 2."""
    name, raw, oc, geom, s = case
    mode = cfdl.SOLVER_PARITY if solver == "parity" else cfdl.SOLVER_MCSGS
    path = str(tmp_path / "state.ckp")
    a = make_solver(cfdl, raw, oc, geom)
    b = make_solver(cfdl, raw, oc, geom)
    try:
        a.set_option("solver", mode); b.set_option("solver", mode)
        a.run(dt=0.01, nit=30, ntstep=2, ncoef=2, want_hist=False)
        a.checkpoint_write(path)
        ha = a.run(dt=0.01, nit=30, ntstep=2, ncoef=2)
        b.checkpoint_read(path)
        hb = b.run(dt=0.01, nit=30, ntstep=2, ncoef=2)
        assert same_history(ha, hb)
        for f in ("u", "v", "w", "p", "gp", "mip", "mip0", "u0"):
            assert np.array_equal(a.download(f), b.download(f)), f
        # a truncated file, a file with a foreign mesh id and a file of another mesh size are refused,
        # and a refused file leaves the handle's state as it was
        good = open(path, "rb").read()
        before = {f: b.download(f) for f in ("u", "p", "mip", "mip0")}
        for bad in (good[: len(good) - 4096], good[:48] + bytes(8) + good[56:], good[:8] + (12345).to_bytes(8, "little") + good[16:],
                    good + b"x"):
            with open(path, "wb") as fh:
                fh.write(bad)
            with pytest.raises(cfdl.CfdlError):
                b.checkpoint_read(path)
            for f, v in before.items():
                assert np.array_equal(b.download(f), v), f
    finally:
        a.close(); b.close()
