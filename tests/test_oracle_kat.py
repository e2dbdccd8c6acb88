"""Known-answer tests that pin the CPU oracle (oracle/*.cpp).

The reference ships no tests, golden vectors or runnable case for this path
and cannot be compiled here (no Fortran compiler); since round 2 the oracle is pinned to the reference's source text
executed by an interpreter (tests/test_oracle_vs_reference_source.py).  These are, in addition, the analytic properties the reference code satisfies by
construction; each test names the reference lines it follows.
"""
import numpy as np
import pytest

from conftest import make_case


def unpack(nb):
    nb = nb.astype(np.int64)
    return nb >> 5, nb & 31


@pytest.fixture(scope="module", params=["hex5", "hex4_jitter", "tet3"])
def mesh(request, cfdl, oracle):
    kw = {"hex5": dict(kind=0, n=5), "hex4_jitter": dict(kind=0, n=4, jitter=0.3),
          "tet3": dict(kind=1, n=3, jitter=0.2, shuffle=True)}[request.param]
    raw, oc, geom = make_case(cfdl, oracle, **kw)
    return request.param, raw, oc, geom


def test_geometry_closure(mesh):
    """calc_aip_xyzip.f90:25-72, calc_vol_cv_centers.f90:31-55: closed cells, unit volume,
    halo centre == face centroid."""
    _, raw, oc, g = mesh
    ne, idx = g["ne"], g["ef2nb_idx"] - 1
    fg = g["ef2nb_fg"].astype(np.int64)
    a = g["aip"].reshape(-1, 3)[np.abs(fg) - 1] * np.sign(fg)[:, None]
    closure = np.add.reduceat(a, idx[:-1], axis=0)
    assert np.abs(closure).max() < 1e-14
    assert abs(g["vol"].sum() - 1.0) < 1e-13
    assert g["vol"].min() > 0
    nbid, lf = unpack(g["ef2nb_nb"])
    halo = lf == 0
    r = g["rip"].reshape(-1, 3)[np.abs(fg[halo]) - 1]
    h = nbid[halo] - 1
    assert np.array_equal(np.c_[g["xc"][h], g["yc"][h], g["zc"][h]], r)


def test_connectivity_invariants(mesh):
    """SURVEY App. C (mod_mg_lvl_uns.f90:351-433): reciprocity, ownership, boundary back-pointers."""
    _, raw, oc, g = mesh
    ne, nf, nbf = g["ne"], g["nf"], g["nbf"]
    idx = g["ef2nb_idx"].astype(np.int64)
    nbid, lf = unpack(g["ef2nb_nb"])
    fg = g["ef2nb_fg"].astype(np.int64)
    for e in range(1, ne + 1):
        for k in range(idx[e - 1], idx[e]):
            s = k - 1
            if lf[s] > 0:
                r = idx[nbid[s] - 1] + lf[s] - 1 - 1
                assert nbid[r] == e and lf[r] == k - idx[e - 1] + 1 and fg[r] == -fg[s]
                assert (fg[s] > 0) == (e < nbid[s])
            else:
                assert ne < nbid[s] <= ne + nbf and fg[s] > 0
                be, bl = unpack(np.abs(g["bs"][nbid[s] - ne - 1:nbid[s] - ne]))
                assert be[0] == e and bl[0] == k - idx[e - 1] + 1
    oe, ol = unpack(g["s2g"])
    slots = idx[oe - 1] + ol - 1 - 1
    assert np.array_equal(fg[slots], np.arange(1, nf + 1))
    cnt = np.bincount(np.abs(fg), minlength=nf + 1)[1:]
    assert cnt.sum() == 2 * nf - nbf and set(cnt.tolist()) <= {1, 2} and (cnt == 1).sum() == nbf


def test_lsq_gradient_exact_for_linear_field(mesh, oracle):
    """mod_solver.f90:48-79: LSQ gradient of a linear field is exact on any mesh."""
    _, raw, oc, g = mesh
    gr = np.array([1.25, -0.5, 3.0])
    phi = 2.0 + gr[0] * g["xc"] + gr[1] * g["yc"] + gr[2] * g["zc"]
    out = oracle.flat_calc_grad(phi, g["xc"], g["yc"], g["zc"], g["ef2nb_idx"], g["ef2nb_nb"])
    assert np.abs(out[:3 * g["ne"]].reshape(-1, 3) - gr).max() < 1e-10


def test_momentum_coefficients_uniform_hex(cfdl, oracle):
    """mod_uvwp.f90:161-286 on the uniform orthogonal hex mesh, against closed forms:
    d = mu*h, ap = sum(d + max(f,0)) + rho*h^3/dt, deferred correction exactly zero, and
    SIMPLEC dc = dt/rho (KATs ortho-zero and simplec)."""
    n = 5
    raw, oc, g = make_case(cfdl, oracle, kind=0, n=n)
    h, rho, mu, dt = 1.0 / n, 5.0, 0.01, 0.01
    rng = np.random.default_rng(5)
    for f in ("u", "v", "w", "u0", "v0", "w0", "gp"):
        oc[f][:] = rng.standard_normal(oc[f].size)
    oc["mip"][:] = 0.01 * rng.standard_normal(oc.nf)
    # constant velocity gradients: secondary-stress sums cancel on interior cells, defc == 0
    for f, val in (("gu", (0.3, -0.2, 0.1)), ("gv", (0.0, 0.4, 0.7)), ("gw", (-0.6, 0.2, 0.5))):
        oc[f][:] = np.tile(val, oc.ne + oc.nbf)
    oc.update_boundaries()
    oc.calc_coef_uvw()
    idx = g["ef2nb_idx"] - 1
    nbid, lf = unpack(g["ef2nb_nb"])
    fg = g["ef2nb_fg"].astype(np.int64)
    f_in = -np.sign(fg) * oc["mip"][np.abs(fg) - 1]
    interior = lf > 0
    anb_want = np.where(interior, mu * h + np.maximum(f_in, 0.0), 2 * mu * h)  # boundary: d = mu*A/(h/2)
    assert np.abs(oc["anb"] - anb_want).max() < 1e-15 + 1e-12 * np.abs(anb_want).max()
    ap_want = np.add.reduceat(anb_want, idx[:-1]) + rho * h ** 3 / dt
    assert np.abs(oc["ap"] - ap_want).max() < 1e-12 * ap_want.max()
    assert np.abs(oc["dc"] - dt / rho).max() < 1e-9 * dt / rho
    assert np.abs(oc["d"] - g["vol"] / oc["ap"]).max() == 0.0
    # interior cell: b = ap0*u0 + sumf*u - vol*gp_x + sumss_x  (sumss_x = mu*h^2 * sum_faces sgn*(gradients . n) = 0)
    inner = np.array([np.all(lf[idx[e]:idx[e + 1]] > 0) for e in range(oc.ne)])
    sumf = np.add.reduceat(np.where(interior, f_in, 0.0), idx[:-1])
    want = rho * h ** 3 / dt * oc["u0"][:oc.ne] + sumf * oc["u"][:oc.ne] - h ** 3 * oc["gp"][0:3 * oc.ne:3]
    assert np.abs(oc["bu"] - want)[inner].max() < 1e-12 * np.abs(want).max()


def test_pc_matrix_properties(mesh):
    """mod_uvwp.f90:305-366: anb symmetric, ap = sum(anb) (singular Neumann), sum(b) = 0."""
    _, raw, oc, g = mesh
    rng = np.random.default_rng(9)
    for f in ("u", "v", "w", "p"):
        oc[f][:] = 0.1 * rng.standard_normal(oc[f].size)
    oc.update_boundaries()
    oc.calc_coef_uvw()
    oc.calc_mip(False)
    oc.update_boundaries()  # boundary mip = 0
    oc.calc_coef_p()
    idx = g["ef2nb_idx"].astype(np.int64)
    nbid, lf = unpack(g["ef2nb_nb"])
    anb, ap, b = oc["anb"], oc["ap"], oc["b"]
    assert np.abs(np.add.reduceat(anb, idx[:-1] - 1) - ap).max() <= 1e-13 * ap.max()
    for e in range(1, g["ne"] + 1):
        for k in range(idx[e - 1], idx[e]):
            if lf[k - 1] > 0:
                r = idx[nbid[k - 1] - 1] + lf[k - 1] - 1
                assert abs(anb[k - 1] - anb[r - 1]) <= 1e-12 * abs(anb[k - 1])
            else:
                assert anb[k - 1] == 0.0
    assert abs(b.sum()) < 1e-12 * np.abs(b).max() * len(b) ** 0.5 + 1e-18


def test_first_iteration_v_w_have_zero_rhs(cfdl, oracle):
    """mod_solver.f90:283-287: from rest only the lid drives u; v and w start with res_i = 0 => it = 0."""
    raw, oc, g = make_case(cfdl, oracle, kind=0, n=6, n_subdomains=1)
    oc.update_boundaries()
    hist = oc.solve_uvwp()
    assert hist[0, 0] >= 1 and hist[0, 1] > 0
    assert hist[1, 0] == 0 and hist[1, 1] == 0.0
    assert hist[2, 0] == 0 and hist[2, 1] == 0.0


def chain_system(n):
    """1-D Laplacian as a 'mesh' in the reference's CSR format: two slots per cell, halo ends."""
    idx = np.arange(0, 2 * n + 1, 2, dtype=np.int32) + 1
    nb = np.zeros(2 * n, np.int32)
    for e in range(1, n + 1):
        left = ((e - 1) << 5) | 2 if e > 1 else ((n + 1) << 5)
        right = ((e + 1) << 5) | 1 if e < n else ((n + 2) << 5)
        nb[2 * (e - 1)] = left
        nb[2 * (e - 1) + 1] = right
    return idx, nb


def test_sgs_hand_computed_iterates(oracle):
    """mod_solver.f90:289-307: forward then backward sweep in natural order, SOR 1.02 for 'pc'."""
    n = 7
    idx, nb = chain_system(n)
    ap = np.full(n, 2.0)
    anb = np.ones(2 * n)
    rng = np.random.default_rng(2)
    b = rng.standard_normal(n)
    phi0 = np.concatenate([rng.standard_normal(n), [0.25, -0.5]])  # two Dirichlet halos
    for is_pc, sor in ((False, 1.0), (True, 1.02)):
        phi = phi0.copy()
        val = lambda i: phi[n] if i == 0 else (phi[n + 1] if i == n + 1 else phi[i - 1])
        for _ in range(2):
            for order in (range(1, n + 1), range(n, 0, -1)):
                for e in order:
                    s = b[e - 1] + anb[2 * e - 2] * val(e - 1)
                    s = s + anb[2 * e - 1] * val(e + 1)
                    phi[e - 1] = (s + (sor - 1.0) * ap[e - 1] * phi[e - 1]) / ap[e - 1] / sor
        got = oracle.flat_smoother_gs(is_pc, phi0, ap, anb, b, idx, nb, nit=2)
        assert np.array_equal(got, phi)
        # solve_gs == smoother_gs + stopping rule on the RMS residual
        got2, st = oracle.flat_solve_gs(is_pc, phi0, ap, anb, b, idx, nb, nit=2)
        if st[0] == 2:
            assert np.array_equal(got2, phi)
        res, _ = oracle.flat_calc_residual(got2, ap, anb, b, idx, nb)
        assert abs(res - st[2]) <= 1e-15 + 1e-13 * abs(res)


def test_mass_closure_after_exact_pc_solve(cfdl, oracle):
    """mod_uvwp.f90:329-340,409-413: with pc solved exactly the corrected fluxes close per cell."""
    raw, oc, g = make_case(cfdl, oracle, kind=0, n=4, n_subdomains=1)
    rng = np.random.default_rng(4)
    for f in ("u", "v", "w"):
        oc[f][:] = 0.1 * rng.standard_normal(oc[f].size)
    oc.update_boundaries()
    oc.calc_coef_uvw()
    oc.calc_mip(False)
    oc.update_boundaries()
    oc.calc_coef_p()
    phi = np.zeros(oc.ne + oc.nbf)
    for _ in range(200):
        phi, st = oc.solve(True, oc["ap"], oc["anb"], oc["b"], phi, nit=100)
    assert st[1] < 1e-15 or st[2] < 1e-15
    oc["phic"][:] = phi
    oc.adjust_pc()
    oc.update_uvwp()
    idx = g["ef2nb_idx"] - 1
    fg = g["ef2nb_fg"].astype(np.int64)
    imbalance = np.add.reduceat(-np.sign(fg) * oc["mip"][np.abs(fg) - 1], idx[:-1])
    assert np.abs(imbalance).max() < 1e-12


@pytest.mark.parametrize("P", [2, 4, 8])
def test_rcb_blocks_on_cube(cfdl, oracle, P):
    """mod_agglomeration.f90:380-561: midpoint bisection gives P equal boxes on the cube."""
    raw, oc, g = make_case(cfdl, oracle, kind=0, n=8, n_subdomains=P)
    blk = oc["gf2g"]
    assert np.array_equal(np.bincount(blk)[1:], np.full(P, oc.ne // P))
    p, pidx = oc["g2gf_p"], oc["g2gf_idx"]
    assert sorted(p.tolist()) == list(range(1, oc.ne + 1))
    for b in range(P):
        cells = p[pidx[b] - 1:pidx[b + 1] - 1]
        assert np.all(blk[cells - 1] == b + 1)
        ext = [np.ptp(g[k][cells - 1]) for k in ("xc", "yc", "zc")]
        vol_box = np.prod([e + 1.0 / 8 for e in ext])
        assert abs(vol_box - 1.0 / P) < 1e-12  # a box, not a scattered set


def test_block_solver_converges_to_the_same_solution(cfdl, oracle):
    """multi_subdomain_solver (mod_solver.f90:124-189) and solve_gs solve the same system."""
    raw1, oc1, g = make_case(cfdl, oracle, kind=0, n=6, n_subdomains=1)
    raw4, oc4, _ = make_case(cfdl, oracle, kind=0, n=6, n_subdomains=4)
    oc1.update_boundaries(); oc1.calc_coef_uvw()
    ap, anb, b = oc1["ap"].copy(), oc1["anb"].copy(), oc1["bu"].copy()
    phi1 = oc1["u"].copy()
    phi4 = phi1.copy()
    for _ in range(40):
        phi1, _ = oc1.solve(False, ap, anb, b, phi1, nit=50)
        phi4, st = oc4.solve(False, ap, anb, b, phi4, nit=50)
    assert st[0] % 2 == 0
    assert np.abs(phi1 - phi4).max() < 1e-12


def qsort_key_nrec_py(key, b):
    """Second, independent transliteration of qsort_key_nRec (mod_util.f90:1683-1730)."""
    key, b = list(key), list(b)
    n = len(key)
    beg, end = {0: 0}, {0: n}
    i = 0
    while i >= 0:
        L, R = beg[i], end[i] - 1
        if L < R:
            piv, b0 = key[L], b[L]
            while L < R:
                while key[R] >= piv and L < R:
                    R -= 1
                if L < R:
                    key[L], b[L] = key[R], b[R]
                    L += 1
                while key[L] <= piv and L < R:
                    L += 1
                if L < R:
                    key[R], b[R] = key[L], b[L]
                    R -= 1
            key[L], b[L] = piv, b0
            beg[i + 1], end[i + 1], end[i] = L + 1, end[i], L
            i += 1
            if end[i] - beg[i] > end[i - 1] - beg[i - 1]:
                beg[i], beg[i - 1] = beg[i - 1], beg[i]
                end[i], end[i - 1] = end[i - 1], end[i]
        else:
            i -= 1
    return key, b


@pytest.mark.parametrize("kw", [dict(kind=0, n=6, n_subdomains=4), dict(kind=1, n=3, jitter=0.2, shuffle=True, n_subdomains=3)])
def test_unstable_block_sort_is_reproduced(cfdl, oracle, kw):
    """The block-local sweep order of the reference is whatever its unstable quicksort yields
    (mod_mg_lvl_uns.f90:883-902); the oracle's g2gf must equal an independent transliteration."""
    raw, oc, g = make_case(cfdl, oracle, **kw)
    key, perm = qsort_key_nrec_py(oc["gf2g"].tolist(), list(range(1, oc.ne + 1)))
    assert key == sorted(key)
    assert perm == oc["g2gf_p"].tolist()
