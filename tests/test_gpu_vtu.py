"""cfdl_write_vtu against a Python restatement of the reference's writer (src/modules/mod_vtu_output.f90:6-326), byte for byte:
XML header lines with the reference's number formatting (I12 offsets, only right-trimmed), UInt64 block sizes, cell data of
the 3-D sections in section order, points, 0-based connectivity, offsets, VTK types — and the writer's side effect on dc."""
import struct

import numpy as np
import pytest

from conftest import make_case, make_solver

pytestmark = pytest.mark.gpu

DIM = {0: 0, 1: 3, 2: 0, 3: 1, 4: 1, 5: 2, 6: 2, 7: 2, 8: 2, 9: 2, 10: 3, 12: 3, 14: 3, 17: 3}
NVX = {10: 4, 12: 5, 14: 6, 17: 8}
VTK = {10: 10, 17: 12, 12: 14, 14: 13}


def reference_vtu_bytes(raw, variables):
    """variables: list of (name, ndim, array in the reference's cell numbering)"""
    lf = "\n"
    nsec, ne, nvx = int(raw["nsec"]), int(raw["ne"]), int(raw["nvx"])
    esec, etype = np.asarray(raw["esec"]).reshape(nsec, 2), np.asarray(raw["etype"])
    m = int(raw["ne2vx_max"])
    e2vx = np.asarray(raw["e2vx"]).reshape(-1, m)
    secs3d = [n for n in range(nsec) if DIM[int(etype[n])] != 2]
    out, noff = [], 0
    out.append('<?xml version="1.0"?>' + lf)
    out.append('<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">' + lf)
    out.append(" <UnstructuredGrid>" + lf)
    out.append('  <Piece NumberOfPoints="%d" NumberOfCells="%d">' % (nvx, ne) + lf)
    out.append('   <CellData Scalars="scalars">' + lf)
    nbytes = []
    for name, k, _ in variables:
        out.append('    <DataArray type="Float64" Name="%s" NumberOfComponents="%d" format="appended" offset="%12d" />' % (name, k, noff) + lf)
        nbytes.append(ne * k * 8)
        noff += 8 + nbytes[-1]
    out.append("   </CellData>" + lf + "   <Points>" + lf)
    out.append('    <DataArray type="Float64" NumberOfComponents="3" format="appended" offset="%12d" />' % noff + lf)
    nb_pts = 3 * nvx * 8
    noff += 8 + nb_pts
    out.append("   </Points>" + lf + "   <Cells>" + lf)
    out.append('    <DataArray type="Int32" Name="connectivity" format="appended" offset="%12d" />' % noff + lf)
    nb_conn = 4 * sum((esec[n, 1] - esec[n, 0] + 1) * NVX[int(etype[n])] for n in secs3d)
    noff += 8 + nb_conn
    out.append('    <DataArray type="Int32" Name="offsets" format="appended" offset="%12d" />' % noff + lf)
    nb_off = 4 * sum(esec[n, 1] - esec[n, 0] + 1 for n in secs3d)
    noff += 8 + nb_off
    out.append('    <DataArray type="Int32" Name="types" format="appended" offset="%12d" />' % noff + lf)
    out.append("   </Cells>" + lf + "  </Piece>" + lf + " </UnstructuredGrid>" + lf + ' <AppendedData encoding="raw">' + lf + "_")
    b = "".join(out).encode()
    for (name, k, arr), nb in zip(variables, nbytes):
        b += struct.pack("<q", nb)
        a = np.asarray(arr, dtype=np.float64).reshape(-1, k)
        for n in secs3d:
            b += a[esec[n, 0] - 1:esec[n, 1]].tobytes()
    b += struct.pack("<q", nb_pts) + np.stack([raw["x"], raw["y"], raw["z"]], axis=1).astype(np.float64).tobytes()
    b += struct.pack("<q", nb_conn)
    for n in secs3d:
        b += (e2vx[esec[n, 0] - 1:esec[n, 1], :NVX[int(etype[n])]] - 1).astype(np.int32).tobytes()
    b += struct.pack("<q", nb_off)
    l, offs = 0, []
    for n in secs3d:
        for _ in range(esec[n, 0], esec[n, 1] + 1):
            l += NVX[int(etype[n])]
            offs.append(l)
    b += np.array(offs, dtype=np.int32).tobytes()
    b += struct.pack("<q", 4 * ne)
    for n in secs3d:
        b += np.full(esec[n, 1] - esec[n, 0] + 1, VTK[int(etype[n])], dtype=np.int32).tobytes()
    return b + (lf + " </AppendedData>" + lf + "</VTKFile>" + lf).encode()


@pytest.mark.parametrize("kind,n", [(0, 5), (1, 3)])
def test_write_vtu_has_the_reference_bytes(cfdl, oracle, tmp_path, kind, n):
    raw, oc, geom = make_case(cfdl, oracle, kind=kind, n=n, jitter=0.2 if kind else 0.1, shuffle=bool(kind))
    s = make_solver(cfdl, raw, oc, geom)
    try:
        s.run(dt=0.01, nit=20, ntstep=1, ncoef=2, want_hist=False)
        fields = {f: s.download(f) for f in ("u", "v", "w", "p", "gpc", "mip")}
        # the writer's side effect: dc(e) = sum of mip(fg)*sign(fg) over the cell's faces in face order (:114-123)
        idx, fg = geom["ef2nb_idx"].astype(np.int64) - 1, geom["ef2nb_fg"].astype(np.int64)
        dc = np.zeros(oc.ne)
        for e in range(oc.ne):
            acc = 0.0
            for j in range(idx[e], idx[e + 1]):
                acc = acc + fields["mip"][abs(fg[j]) - 1] * (1.0 if fg[j] > 0 else -1.0)
            dc[e] = acc
        path = str(tmp_path / "uvwp.vtu")
        s.write_vtu(path, raw, equation=0)
        assert np.array_equal(s.download("dc"), dc)
        want = reference_vtu_bytes(raw, [("u", 1, fields["u"]), ("v", 1, fields["v"]), ("w", 1, fields["w"]), ("p", 1, fields["p"]),
                                         ("gpc", 3, fields["gpc"]), ("mip", 1, dc)])
        assert open(path, "rb").read() == want
        # energy and scalar output
        s.energy_init(); s.scalar_init(bc_value=np.arange(oc.nbc) % 2)
        s.update_boundaries(); s.solve_energy(0.01, 10); s.solve_scalar(0.01, 10)
        for eq, names in ((1, (("enthalpy", 1, "h"), ("grad_enthalpy", 3, "gh"), ("temperature", 1, "t"), ("grad_temperature", 3, "gt"))),
                          (2, (("phi", 1, "s"), ("grad", 3, "gs")))):
            path = str(tmp_path / ("eq%d.vtu" % eq))
            s.write_vtu(path, raw, equation=eq)
            want = reference_vtu_bytes(raw, [(nm, k, s.download(f)) for nm, k, f in names])
            assert open(path, "rb").read() == want
    finally:
        s.close()
