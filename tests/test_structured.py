"""cfdl_create_structured_hex: the per-rank analytic generator for the n^3 cavity (the way past
the reference's 2^26 packed-id limit to the 512^3 target) must describe exactly the mesh that
cfdl_meshgen_fill + cfdl_mesh_build — the reference-format route, itself pinned against the
oracle's set-up in test_golden.py — produce: same neighbours, same face orientation, the same
geometry bits."""
import numpy as np
import pytest


@pytest.mark.parametrize("n,nranks", [(4, 1), (6, 4), (7, 2), (8, 8)])
def test_structured_arrays_equal_general_builder(cfdl, n, nranks):
    geom = cfdl.mesh_build(cfdl.meshgen(cfdl.MESH_HEX, n))
    o = cfdl.structured_hex_arrays(n, nranks)
    nb_ref = (geom["ef2nb_nb"].astype(np.uint32) >> 5).astype(np.int64) - 1
    assert np.array_equal(nb_ref, o["nb"])
    fg_ref = geom["ef2nb_fg"]
    assert np.array_equal(np.sign(fg_ref), np.sign(o["fg"]))
    for k in ("xc", "yc", "zc", "vol"):
        assert np.array_equal(geom[k], o[k]), k
    for k in ("aip", "rip"):  # per slot, through each numbering's own face id
        want = geom[k].reshape(-1, 3)[np.abs(fg_ref) - 1]
        got = o[k].reshape(-1, 3)[np.abs(o["fg"]) - 1]
        assert np.array_equal(want, got), k
    # the two face numberings are a bijection of each other
    pairs = set(zip(np.abs(fg_ref).tolist(), np.abs(o["fg"]).tolist()))
    assert len(pairs) == geom["nf"] == 3 * n * n * (n + 1)
    if nranks > 1 and n % 2 == 0:  # even n: the bisection is the reference's RCB
        c2r = cfdl.partition_rcb(geom, nranks, want_order=False)[0]
        assert np.array_equal(c2r, o["cell2rank"])
    if nranks > 1:
        assert sorted(set(o["cell2rank"].tolist())) == list(range(1, nranks + 1))


def test_structured_rejects_bad_arguments(cfdl):
    with pytest.raises(cfdl.CfdlError):
        cfdl.structured_hex_arrays(1)
    with pytest.raises(cfdl.CfdlError):
        cfdl.structured_hex_arrays(6, 3)
    with pytest.raises(cfdl.CfdlError):
        cfdl.structured_hex_arrays(701)


@pytest.mark.gpu
def test_structured_handle_equals_general_handle(cfdl):
    n = 14
    raw = cfdl.meshgen(cfdl.MESH_HEX, n)
    geom = cfdl.mesh_build(raw)
    a = cfdl.Solver(geom, cfdl.default_bcs(raw))
    b = cfdl.Solver.structured_hex(n)
    for s in (a, b):
        s.set_option("solver", cfdl.SOLVER_MCSGS)
    ha = a.run(dt=0.01, nit=100, ntstep=2, ncoef=3)
    hb = b.run(dt=0.01, nit=100, ntstep=2, ncoef=3)
    from conftest import same_history
    assert same_history(ha, hb)
    for f in ("u", "v", "w", "p", "pc", "gp", "gu"):
        assert np.array_equal(a.download(f), b.download(f)), f
    # face fields: same values, structured numbering (x-, y-, z-normal faces)
    o = cfdl.structured_hex_arrays(n)
    ref_id, own_id = np.abs(geom["ef2nb_fg"]) - 1, np.abs(o["fg"]) - 1
    assert np.array_equal(a.download("mip")[ref_id], b.download("mip")[own_id])
    a.close()
    b.close()
