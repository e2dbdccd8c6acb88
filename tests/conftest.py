import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--emul", action="store_true", default=False,
                     help="run the gpu-marked tests against tests/emul/_build/libcfdl_emul.so (host emulation of the "
                          "CUDA sources, test infrastructure only) instead of the real library")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    if config.getoption("--emul"):
        use_emulated_library()


EMUL_LIB = os.path.join(ROOT, "tests", "emul", "_build", "libcfdl_emul.so")
EMULATED = False  # True when this test process runs against the cuemu build (a few tests then use smaller meshes)


def build_emulated_library():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul"), "-j", "8"])
    return EMUL_LIB


def use_emulated_library():
    """Point the ctypes mirror at the cuemu build for THIS test process (never done by the product)."""
    import ctypes
    import cfdl as m
    build_emulated_library()
    m._lib = ctypes.CDLL(EMUL_LIB)
    m._lib.cfdl_last_error.restype = ctypes.c_char_p
    assert m._lib.cfdl_emulated() == 1
    m.LIB_PATH = EMUL_LIB
    global EMULATED
    EMULATED = True


@pytest.fixture(scope="session")
def cfdl():
    import cfdl as m
    return m


@pytest.fixture(scope="session")
def oracle():
    import oracle as m
    m.build()
    return m


def make_case(cfdl, oracle, kind, n, jitter=0.0, shuffle=False, n_subdomains=1, seed=12345, symmetry=()):
    """Synthetic mesh -> (raw, oracle case, geometry dict in the reference's array format).
    symmetry: indices (2-D section order) of the BCs that become `symmetry` (mod_uvwp.f90:517-546) instead of walls."""
    raw = cfdl.meshgen(kind, n, jitter=jitter, shuffle=shuffle, seed=seed)
    oc = oracle.OracleCase(raw, n_subdomains=n_subdomains)
    for i in symmetry:
        oc.set_bc(i, 2)  # BC_SYMMETRY: mirrored halo velocity, bc_type 'zero_flux'
    geom = dict(ne=oc.ne, nf=oc.nf, nbf=oc.nbf)
    for k in ("ef2nb_idx", "ef2nb_nb", "ef2nb_fg", "s2g", "bs", "xc", "yc", "zc", "aip", "rip", "vol"):
        geom[k] = oc[k].copy()
    return raw, oc, geom


def make_solver(cfdl, raw, oc, geom, **kw):
    bcs = oc.bc_table()
    if oc.n_subdomains > 1:
        return cfdl.Solver(geom, bcs, n_subdomains=oc.n_subdomains, g2gf_p=oc["g2gf_p"].copy(), g2gf_idx=oc["g2gf_idx"].copy(), **kw)
    return cfdl.Solver(geom, bcs, **kw)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def same_history(a, b, rtol=1e-12):
    """Residual histories (it, res_i, res_f, res_max) of two runs that must be the same computation: iteration
    counts, opening residuals and maxima are exact; the RMS norms are sums whose order depends on the launch
    geometry of the solver form a handle has chosen by timing, so they agree to rounding only."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(a[..., 0], b[..., 0]), (a[..., 0], b[..., 0])
    assert np.allclose(a[..., 1:], b[..., 1:], rtol=rtol, atol=0.0), np.abs(a[..., 1:] - b[..., 1:]).max()
    return True
