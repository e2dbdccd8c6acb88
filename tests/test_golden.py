"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them (CPU), and the CUDA path must reproduce them through the
C ABI (GPU) — residual history and final fields to 1e-10 relative (north_star tolerance)."""
import os

import numpy as np
import pytest

from conftest import make_solver, rel_err

import importlib.util

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

TOL = 1e-10


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_reproduces_golden(cfdl, oracle, name):
    kw = mg.CASES[name]
    gold = load(name)
    raw = cfdl.meshgen(kw["kind"], kw["n"], jitter=kw["jitter"], shuffle=kw["shuffle"], seed=12345)
    oc = oracle.OracleCase(raw, n_subdomains=kw["n_subdomains"])
    chk = np.array([oc["vol"].sum(), oc["aip"].sum(), oc["rip"].sum(), float(oc["ef2nb_nb"].astype(np.int64).sum()),
                    float(oc["ef2nb_fg"].astype(np.int64).sum())])
    assert rel_err(chk, gold["geom_checksum"]) < 1e-13, "synthetic mesh generator changed"
    hist, _ = oc.run(kw["ntstep"], kw["ncoef"])
    assert np.array_equal(hist[:, :, 0], gold["hist"][:, :, 0])
    assert rel_err(hist[:, :, 1:3], gold["hist"][:, :, 1:3]) < 1e-12
    for f in ("u", "v", "w", "p", "mip", "gp"):
        assert rel_err(oc[f], gold[f]) < 1e-12, f
    if kw["n_subdomains"] > 1:
        assert np.array_equal(oc["g2gf_p"], gold["g2gf_p"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.CASES))
def test_cuda_reproduces_golden(cfdl, name):
    """No oracle involved: mesh from the product's own builder, block order from the fixture."""
    kw = mg.CASES[name]
    gold = load(name)
    raw = cfdl.meshgen(kw["kind"], kw["n"], jitter=kw["jitter"], shuffle=kw["shuffle"], seed=12345)
    geom = cfdl.mesh_build(raw)
    P = kw["n_subdomains"]
    if P > 1:
        _, p, idx = cfdl.partition_rcb(geom, P)  # the product's own RCB + block order
        assert np.array_equal(p, gold["g2gf_p"])
        s = cfdl.Solver(geom, cfdl.default_bcs(raw), n_subdomains=P, g2gf_p=p, g2gf_idx=idx)
    else:
        s = cfdl.Solver(geom, cfdl.default_bcs(raw))
    try:
        hist = s.run(dt=0.01, nit=100, ntstep=kw["ntstep"], ncoef=kw["ncoef"])
        assert np.array_equal(hist[:, :, 0], gold["hist"][:, :, 0])
        assert rel_err(hist[:, :, 1:3], gold["hist"][:, :, 1:3]) < TOL
        for f in ("u", "v", "w", "p", "mip", "gp"):
            assert rel_err(s.download(f)[:len(gold[f])], gold[f]) < TOL, f
    finally:
        s.close()
