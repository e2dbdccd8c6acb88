#!/usr/bin/env python3
"""Medium-size runs on the cuemu build (grids of tens of CTAs, several rows per thread — the unit
tests' meshes fit in one or two CTAs): (1) exact mode against the oracle on an n^3 hex cavity,
(2) the throughput mode as shipped (face statics, side-by-side momentum passes, persistent pc solve, 16-bit
offsets, dependent launches) against the same mode with all of that off (the reference's forms): identical bits.
TEST INFRASTRUCTURE ONLY.   usage: medium_check.py <n>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cfdl  # noqa: E402
import conftest  # noqa: E402
import oracle  # noqa: E402


def main():
    n = int(sys.argv[1])
    conftest.use_emulated_library()
    oracle.build()
    raw = cfdl.meshgen(cfdl.MESH_HEX, n, jitter=0.15)
    geom = cfdl.mesh_build(raw)
    bcs = cfdl.default_bcs(raw)
    # (1) exact natural-order mode, 4 subdomains like the reference's default, against the oracle
    oc = oracle.OracleCase(raw, n_subdomains=4)
    s = cfdl.Solver(geom, bcs, n_subdomains=4, g2gf_p=oc["g2gf_p"].copy(), g2gf_idx=oc["g2gf_idx"].copy(), device=0)
    s.set_option("solver", cfdl.SOLVER_PARITY)
    want, _ = oc.run(1, 2)
    got = s.run(dt=0.01, nit=100, ntstep=1, ncoef=2)
    assert np.array_equal(got[:, :, 0], want[:, :, 0]), (got[:, :, 0], want[:, :, 0])
    worst = 0.0
    for f in ("u", "v", "w", "p"):
        a, b = s.download(f), oc[f]
        worst = max(worst, np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    assert worst <= 1e-10, worst
    s.close()
    # (2) throughput mode: shipped paths vs all of them off
    res = {}
    for new in (0, 1):
        s = cfdl.Solver(geom, bcs, device=0)
        s.set_option("solver", cfdl.SOLVER_MCSGS)
        if not new:
            for k, v in (("uvw_fused", 0), ("pc_sumap", 0), ("grad_variant", 0), ("statics", 0), ("rbq", 0), ("rb_idx16", 0), ("pdl", 0)):
                s.set_option(k, v)
        h = s.run(dt=0.5, nit=100, ntstep=2, ncoef=3)
        res[new] = (h, {f: s.download(f) for f in ("u", "v", "w", "p", "gu", "gp", "mip")},
                    [int(s.get_info("rbq_active"))])
        s.close()
    assert np.array_equal(res[0][0][:, :, 0], res[1][0][:, :, 0])
    assert np.allclose(res[0][0], res[1][0], rtol=1e-12, atol=0.0)
    for f, v in res[0][1].items():
        assert np.array_equal(v, res[1][1][f]), f
    assert res[1][2] == [1] and res[0][2] == [0]
    print("medium emulation ok: n=%d exact-mode err %.1e; persistent pc solve active = %s; momentum its %s"
          % (n, worst, res[1][2], res[1][0][-1, :3, 0].astype(int).tolist()))


if __name__ == "__main__":
    main()
