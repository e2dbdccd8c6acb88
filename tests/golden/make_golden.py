"""Generates tests/golden/*.npz from the CPU oracle (oracle/): residual history and final
fields of short runs of the main.f90:50-63 schedule on small synthetic meshes.

The reference itself has no golden vectors and cannot be run here (no Fortran compiler, no
CGNS), so these fixtures pin the ORACLE's behaviour (regression pins), not the reference's;
the analytic KATs in tests/test_oracle_kat.py are what ties the oracle to the reference maths.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = {
    "hex6_sub4": dict(kind=0, n=6, jitter=0.0, shuffle=False, n_subdomains=4, ntstep=2, ncoef=3),
    "hex5_jitter_sub1": dict(kind=0, n=5, jitter=0.25, shuffle=False, n_subdomains=1, ntstep=2, ncoef=3),
    "tet3_shuffled_sub1": dict(kind=1, n=3, jitter=0.2, shuffle=True, n_subdomains=1, ntstep=2, ncoef=3),
    "tet3_shuffled_sub2": dict(kind=1, n=3, jitter=0.2, shuffle=True, n_subdomains=2, ntstep=1, ncoef=3),
}


def main():
    import cfdl
    import oracle
    for name, kw in CASES.items():
        raw = cfdl.meshgen(kw["kind"], kw["n"], jitter=kw["jitter"], shuffle=kw["shuffle"], seed=12345)
        oc = oracle.OracleCase(raw, n_subdomains=kw["n_subdomains"])
        hist, _ = oc.run(kw["ntstep"], kw["ncoef"])
        out = {k: oc[k].copy() for k in ("u", "v", "w", "p", "mip", "gp")}
        out["hist"] = hist
        out["geom_checksum"] = np.array([oc["vol"].sum(), oc["aip"].sum(), oc["rip"].sum(), float(oc["ef2nb_nb"].astype(np.int64).sum()),
                                         float(oc["ef2nb_fg"].astype(np.int64).sum())])
        if kw["n_subdomains"] > 1:
            out["g2gf_p"] = oc["g2gf_p"].copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "cells", oc.ne, "pc its", hist[:, 3, 0].tolist())


if __name__ == "__main__":
    main()
