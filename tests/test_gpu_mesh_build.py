"""cfdl_mesh_build_gpu (connectivity + geometry on the GPU) against the host builder cfdl_mesh_build and against
the set-up arrays of the reference's own source (tests/golden/ref_*.npz: find_element_nb, calc_aip_xyzip_uns,
calc_vol_cv_centers_uns executed by the source interpreter): every array bit for bit."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ARRAYS = ("ef2nb_idx", "ef2nb_nb", "ef2nb_fg", "s2g", "bs", "xc", "yc", "zc", "aip", "rip", "vol")


@pytest.mark.parametrize("kind,n,jitter,shuffle", [(0, 3, 0.0, False), (0, 9, 0.2, False), (0, 20, 0.0, False), (1, 2, 0.2, True), (1, 5, 0.2, True), (1, 7, 0.0, False)])
def test_gpu_mesh_build_equals_host_builder(cfdl, kind, n, jitter, shuffle):
    raw = cfdl.meshgen(kind, n, jitter=jitter, shuffle=shuffle, seed=4711)
    want = cfdl.mesh_build(raw)
    got = cfdl.mesh_build(raw, gpu=True)
    for k in ARRAYS:
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "ref_*.npz"))))
def test_gpu_mesh_build_equals_reference_source(cfdl, path):
    g = np.load(path)
    kind, n = int(g["case"][0]), int(g["case"][1])
    raw = cfdl.meshgen(kind, n, jitter=float(g["jitter"]), shuffle=bool(g["shuffle"]), seed=12345)
    got = cfdl.mesh_build(raw, gpu=True)
    for k in ARRAYS:
        assert np.array_equal(np.asarray(got[k]), g["setup_" + k]), k


def test_gpu_mesh_build_refuses_a_broken_mesh(cfdl):
    raw = cfdl.meshgen(0, 4)
    raw = dict(raw)
    e2vx = np.array(raw["e2vx"]).copy()
    e2vx[3] = e2vx[2]  # a degenerate cell: one of its faces no longer matches its neighbour's
    raw["e2vx"] = e2vx
    with pytest.raises(cfdl.CfdlError):
        cfdl.mesh_build(raw, gpu=True)
