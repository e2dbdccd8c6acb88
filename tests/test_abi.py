"""The C-ABI boundary: libcfdl.so loads on a CPU-only box, exports every symbol declared in
include/cfdl.h, runs its host-side mesh tooling, and refuses compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, make_case


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cfdl.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cfdl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(cfdl):
    L = cfdl.lib()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_error_string(cfdl):
    L = cfdl.lib()
    assert L.cfdl_version() >= 100
    assert isinstance(L.cfdl_last_error(), bytes)


def test_create_fails_loudly_without_gpu(cfdl, oracle):
    if cfdl.device_count() > 0:
        pytest.skip("a GPU is present")
    raw, oc, geom = make_case(cfdl, oracle, kind=0, n=3)
    with pytest.raises(cfdl.CfdlError) as e:
        cfdl.Solver(geom, oc.bc_table())
    assert "no CUDA device" in str(e.value)


def test_meshgen_sizes_and_errors(cfdl):
    raw = cfdl.meshgen(cfdl.MESH_HEX, 4)
    assert (raw["ne"], raw["nbf"], raw["nvx"]) == (64, 96, 125)
    raw = cfdl.meshgen(cfdl.MESH_TET, 3, jitter=0.2, shuffle=True)
    assert (raw["ne"], raw["nbf"]) == (162, 108)
    with pytest.raises(cfdl.CfdlError):
        cfdl.meshgen(cfdl.MESH_HEX, 0)
    with pytest.raises(cfdl.CfdlError):
        cfdl.meshgen(7, 4)


@pytest.mark.parametrize("kw", [dict(kind=0, n=5), dict(kind=0, n=4, jitter=0.3, shuffle=True),
                                dict(kind=1, n=3, jitter=0.2, shuffle=True), dict(kind=1, n=5)])
def test_mesh_build_equals_reference_setup(cfdl, oracle, kw):
    """cfdl_mesh_build (sort-based) == the oracle's line-by-line restatement of find_element_nb,
    calc_aip_xyzip_uns, calc_vol_cv_centers_uns — bit for bit, including face numbering."""
    raw, oc, geom = make_case(cfdl, oracle, **kw)
    g = cfdl.mesh_build(raw)
    for k in geom:
        if k in ("ne", "nf", "nbf"):
            assert g[k] == geom[k]
        else:
            assert np.array_equal(g[k], geom[k]), k


def test_mesh_build_rejects_broken_mesh(cfdl):
    raw = cfdl.meshgen(cfdl.MESH_HEX, 3)
    bad = dict(raw)
    bad["e2vx"] = raw["e2vx"].copy()
    bad["e2vx"][-8:-4] = bad["e2vx"][-16:-12]  # duplicate a boundary quad (rows are 8 wide): a face is left uncovered
    with pytest.raises(cfdl.CfdlError):
        cfdl.mesh_build(bad)


def test_default_bcs_follow_section_names(cfdl):
    raw = cfdl.meshgen(cfdl.MESH_HEX, 3)
    esec, kind, uvw = cfdl.default_bcs(raw)
    names = [raw["names"][32 * s:32 * s + 32].decode().strip() for s in range(1, 7)]
    assert names == ["bottom", "top", "west", "east", "south", "north"]
    assert kind.tolist() == [0, 1, 0, 0, 0, 0] and uvw[3:6].tolist() == [1.0, 0.0, 0.0]
    assert esec[0] == raw["ne"] + 1 and esec[-1] == raw["ne"] + raw["nbf"]
