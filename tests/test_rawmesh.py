"""Raw mesh file (cfdl_rawmesh_write / _sizes / _read): the CGNS-free input path for the reference's
driver (cfd-lite_b200/fortran/mod_rawmesh.f90 reads the same layout).  Round trip must return the
very arrays that were written, and the connectivity/geometry built from the file must equal the one
built from the generator's arrays."""
import os
import struct

import numpy as np
import pytest


@pytest.mark.parametrize("kind,n", [(0, 5), (1, 3)])
def test_rawmesh_round_trip(cfdl, tmp_path, kind, n):
    raw = cfdl.meshgen(kind, n, jitter=0.2 if kind else 0.0, shuffle=bool(kind))
    path = str(tmp_path / "mesh.raw")
    cfdl.rawmesh_write(raw, path)
    back = cfdl.rawmesh_read(path)
    for k in ("nvx", "ne", "nbf", "nsec", "ne2vx_max"):
        assert back[k] == raw[k], k
    for k in ("x", "y", "z", "e2vx", "etype", "esec"):
        assert np.array_equal(back[k], raw[k]), k
    for s in range(raw["nsec"]):
        assert bytes(back["names"][32 * s:32 * s + 32]).rstrip(b" \0") == bytes(raw["names"][32 * s:32 * s + 32]).rstrip(b" \0")
    g0, g1 = cfdl.mesh_build(raw), cfdl.mesh_build(back)
    for k, v in g0.items():
        assert np.array_equal(v, g1[k]), k
    # header layout the Fortran reader relies on
    with open(path, "rb") as f:
        head = f.read(32)
    assert head[:8] == b"CFDLRAW1"
    nvx, nelem, nsec, w = struct.unpack("<qqii", head[8:32])
    assert (nvx, nelem, nsec, w) == (raw["nvx"], raw["ne"] + raw["nbf"], raw["nsec"], raw["ne2vx_max"])
    assert os.path.getsize(path) == 32 + nsec * 44 + 24 * nvx + 4 * w * nelem


def test_rawmesh_rejects_other_files(cfdl, tmp_path):
    p = tmp_path / "junk.raw"
    p.write_bytes(b"not a mesh at all, just bytes" * 4)
    with pytest.raises(cfdl.CfdlError):
        cfdl.rawmesh_read(str(p))
    with pytest.raises(cfdl.CfdlError):
        cfdl.rawmesh_read(str(tmp_path / "missing.raw"))
