"""The assembly routines with host arrays (cfdl_host_calc_coef_uvw, _calc_mip, _calc_coef_p,
_adjust_pc, _update_uvwp): the flattened form of the reference's derived-type signatures
(SURVEY §8b).  Each is upload -> the routine tested in test_gpu_parity.py -> download, so the
results must again equal the oracle's bit for bit.

Confirmed on a B200 in round 2 (profiles/r02_call1_pytest_gpu.log: the four former non-strict xfails XPASSed),
since then plain tests."""
import numpy as np
import pytest

from conftest import make_case, make_solver, rel_err

pytestmark = pytest.mark.gpu

STATE = ("u", "v", "w", "p", "u0", "v0", "w0", "gu", "gv", "gw", "gp", "mip", "mip0")


@pytest.fixture(scope="module", params=["hex", "tet"])
def case(request, cfdl, oracle):
    kw = dict(kind=0, n=7) if request.param == "hex" else dict(kind=1, n=3, jitter=0.2, shuffle=True)
    raw, oc, geom = make_case(cfdl, oracle, **kw)
    s = make_solver(cfdl, raw, oc, geom)
    rng = np.random.default_rng(5)
    for name in STATE:
        a = oc[name]
        a[:] = rng.standard_normal(a.size) * (0.01 if name.startswith("mip") else 1.0)
    oc["phic"][:] = rng.standard_normal(oc["phic"].size)
    oc.update_boundaries()  # halo values and boundary fluxes as the BC callbacks leave them
    yield oc, s
    s.close()


def test_host_assembly_chain(case):
    oc, s = case
    f = {k: oc[k].copy() for k in STATE}
    oc.calc_coef_uvw()
    got = s.host_calc_coef_uvw(f, dt=0.01)
    for k, ok in (("ap", "ap"), ("anb", "anb"), ("bu", "bu"), ("bv", "bv"), ("bw", "bw"), ("d", "d"), ("dc", "dc")):
        assert np.array_equal(got[k], oc[ok]), k
    f["d"], f["dc"] = got["d"], got["dc"]
    oc.calc_mip(True)
    mip = s.host_calc_mip(f, rhie_chow=True, dt=0.01)
    assert np.array_equal(mip, oc["mip"])
    oc.calc_coef_p()
    gp_ = s.host_calc_coef_p(f["dc"], mip)
    for k in ("ap", "anb", "b"):
        assert np.array_equal(gp_[k], oc[k]), k
    pc_in = oc["phic"].copy()
    oc.adjust_pc()
    pc = s.host_adjust_pc(pc_in)
    assert np.array_equal(pc, oc["phic"])
    oc["gpc"][:] = oc.calc_grad(oc["phic"])
    p_in, gp_in = oc["p"].copy(), oc["gp"].copy()
    oc.update_uvwp()
    p, gp, mip2 = s.host_update_uvwp(pc, oc["gpc"], f["dc"], p_in, gp_in, mip)
    assert rel_err(p, oc["p"]) <= 1e-12 and rel_err(gp, oc["gp"]) <= 1e-12
    assert np.array_equal(mip2, oc["mip"])


def test_coef_uvw_on_statics_keeps_the_bits(case):
    """calc_coef_uvw on the face statics (ten quotients per face from two stored reciprocals + FMA corrections,
    locality order) and in the reference's form (statics = 0) against the oracle: same bits."""
    oc, s = case
    for name in STATE:
        s.upload(name, oc[name])
    oc.calc_coef_uvw()
    try:
        for statics in (1, 0):
            s.set_option("statics", statics)
            s.calc_coef_uvw(dt=0.01)
            for f in ("ap", "anb", "bu", "bv", "bw", "d", "dc"):
                assert np.array_equal(s.download(f), oc[f]), (statics, f)
    finally:
        s.set_option("statics", 1)
