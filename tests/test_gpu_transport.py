"""GPU parity of the energy and scalar equations (kernels_transport.cu) against the CPU oracle, whose
restatement of mod_energy.f90 / mod_scalar.f90 is pinned bit for bit to the reference's source text
(tests/test_oracle_vs_reference_source.py).  Random states on both sides: the assembled matrices and right-hand sides are
compared bit for bit (the kernels follow the reference's operation order), the solved fields to 1e-12 in both solver
modes (parity mode = the reference's natural-order sweeps; multicolour mode solves the same system to the same stopping
rule, so only the assembly and the stopping behaviour are asserted there)."""
import numpy as np
import pytest

from conftest import make_case, make_solver, rel_err

pytestmark = pytest.mark.gpu

CASES = {
    "hex7_jitter": dict(kind=0, n=7, jitter=0.2, n_subdomains=1),
    "tet3_shuffled": dict(kind=1, n=3, jitter=0.2, shuffle=True, n_subdomains=1),
    "hex6_sub4": dict(kind=0, n=6, n_subdomains=4),
}


@pytest.fixture(scope="module", params=list(CASES))
def case(request, cfdl, oracle):
    raw, oc, geom = make_case(cfdl, oracle, **CASES[request.param])
    s = make_solver(cfdl, raw, oc, geom)
    yield request.param, raw, oc, geom, s
    s.close()


def test_energy_and_scalar_equations(case, cfdl):
    name, raw, oc, geom, s = case
    rng = np.random.default_rng(11)
    tc = 5.0 * (1.0 + 0.3 * rng.random(oc.ne))
    cp = 1000.0 * (1.0 + 0.2 * rng.random(oc.ne))
    oc.construct_energy()
    oc["tc"][:] = tc
    oc["cp"][:] = cp
    oc["h"][:oc.ne] = oc["t"][:oc.ne] * cp
    oc["h0"][:] = oc["h"]
    bcv = rng.integers(0, 2, oc.nbc).astype(np.float64)
    oc.construct_scalar(dcoef=0.7, vel=(3.0, -2.0, 5.0), bc_value=bcv)
    s.energy_init(tc=tc, cp=cp)
    s.scalar_init(dcoef=0.7, vel=(3.0, -2.0, 5.0), bc_value=bcv)
    # a velocity / mass-flux state for the advection terms, and perturbed temperatures / scalars
    for f in ("u", "v", "w", "p", "mip"):
        a = oc[f]
        a[:] = rng.standard_normal(a.size) * (0.01 if f == "mip" else 1.0)
        s.upload(f, a)
    for f, scale in (("t", 300.0), ("h", 3.0e5), ("h0", 3.0e5), ("s", 1.0), ("s0", 1.0)):
        a = oc[f]
        a[:] = scale * (1.0 + 0.05 * rng.standard_normal(a.size))
        s.upload(f, a)
    for mode in (cfdl.SOLVER_PARITY, cfdl.SOLVER_MCSGS):
        s.set_option("solver", mode)
        for it in range(2):
            oc.update_boundaries(); s.update_boundaries()
            for f in ("t", "h", "s"):
                assert np.array_equal(s.download(f), oc[f]), (name, "boundary callbacks", f)
            # scalar
            want = oc.solve_scalar()
            ap, anb, b = oc["ap"].copy(), oc["anb"].copy(), oc["b"].copy()
            got = s.solve_scalar(0.01, 100)
            assert np.array_equal(s.download("ap"), ap) and np.array_equal(s.download("anb"), anb) and np.array_equal(s.download("b"), b), (name, "calc_coef_scalar")
            assert rel_err(s.download("gs"), oc["gs"]) < 1e-12
            if mode == cfdl.SOLVER_PARITY:
                assert got[0] == want[0] and rel_err(got[1:3], want[1:3]) < 1e-10
                assert rel_err(s.download("s"), oc["s"]) < 1e-12
            else:
                assert np.isfinite(got).all() and (got[0] == 100 or got[2] <= got[1] / 10.0)
                s.upload("s", oc["s"])  # continue from the same state
            # energy
            want = oc.solve_energy()
            ap, anb, b = oc["ap"].copy(), oc["anb"].copy(), oc["b"].copy()
            got = s.solve_energy(0.01, 100)
            assert np.array_equal(s.download("ap"), ap) and np.array_equal(s.download("anb"), anb) and np.array_equal(s.download("b"), b), (name, "calc_coef_energy")
            assert rel_err(s.download("gt"), oc["gt"]) < 1e-12 and rel_err(s.download("gh"), oc["gh"]) < 1e-12
            if mode == cfdl.SOLVER_PARITY:
                assert got[0] == want[0] and rel_err(got[1:3], want[1:3]) < 1e-10
                assert rel_err(s.download("h"), oc["h"]) < 1e-12 and rel_err(s.download("t"), oc["t"]) < 1e-12
            else:
                assert np.isfinite(got).all() and (got[0] == 100 or got[2] <= got[1] / 10.0)
                s.upload("h", oc["h"]); s.upload("t", oc["t"])
            oc.update_time(); s.update_time()
            assert np.array_equal(s.download("h0"), s.download("h")) and np.array_equal(s.download("s0"), s.download("s"))
    s.set_option("solver", cfdl.SOLVER_PARITY)


def test_transport_fields_need_their_init(cfdl, oracle):
    raw, oc, geom = make_case(cfdl, oracle, kind=0, n=4)
    s = make_solver(cfdl, raw, oc, geom)
    try:
        with pytest.raises(cfdl.CfdlError):
            s.download("t")
        with pytest.raises(cfdl.CfdlError):
            s.solve_energy()
        with pytest.raises(cfdl.CfdlError):
            s.solve_scalar()
    finally:
        s.close()
