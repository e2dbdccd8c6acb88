"""Energy and scalar equations on partitioned handles against the same run on one rank, on the cuemu build (ranks = threads of
one process; see multirank_check.py).  usage: multirank_transport_check.py WORLD N [nccl|slabs]

Every rank: cfdl_energy_init with per-cell tc / cp given in the reference numbering, cfdl_scalar_init, then time steps of
update_boundaries, solve_uvwp, solve_scalar, solve_energy, update_time.  t, phi (h), s and their gradients must equal the
single-rank fields (1e-12; the residual norms are summed per rank, then in rank order), the solver iteration counts exactly."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cfdl  # noqa: E402
import conftest  # noqa: E402


def drive(s, tc, cp, bcv, steps=3):
    s.energy_init(tc=tc, cp=cp)
    s.scalar_init(dcoef=0.7, vel=(3.0, -2.0, 5.0), bc_value=bcv)
    its = []
    for _ in range(steps):
        s.update_boundaries()
        s.solve_uvwp(0.01, 30)
        hs = s.solve_scalar(0.01, 100)
        he = s.solve_energy(50.0, 100)  # (a long time step: several iterations)
        s.update_time()
        its.append((hs, he))
    return its


def main():
    world, n = int(sys.argv[1]), int(sys.argv[2])
    how = sys.argv[3] if len(sys.argv) > 3 else "p2p"
    nccl, slabs = how == "nccl", how == "slabs"
    if nccl:
        os.environ["CFDL_NCCL_PATH"] = os.path.join(ROOT, "tests", "emul", "_build", "libnccl_emul.so")
    conftest.use_emulated_library()
    raw = cfdl.meshgen(0, n, jitter=0.15)
    geom = cfdl.mesh_build(raw)
    bcs = cfdl.default_bcs(raw)
    ne = int(raw["ne"])
    rng = np.random.default_rng(5)
    tc = 5.0 * (1.0 + 0.3 * rng.random(ne))
    cp = 1000.0 * (1.0 + 0.2 * rng.random(ne))
    bcv = rng.integers(0, 2, len(bcs[1])).astype(np.float64)
    c2r, _, _ = cfdl.partition_rcb(geom, world)
    if slabs:
        k = np.arange(n ** 3) // (n * n)
        c2r = np.zeros(n ** 3, np.int32)
        for r in range(world):
            c2r[(k >= n * r // world) & (k < n * (r + 1) // world)] = r + 1
    bar = threading.Barrier(world)
    handles = [None] * world
    out = [None] * world
    errs = []
    wanted = ("t", "h", "h0", "s", "s0", "gt", "gh", "gs", "u", "p")

    def rank_main(rank):
        try:
            s = cfdl.Solver(geom, bcs, device=0, cell2rank=c2r, rank=rank, nranks=world)
            s.set_option("solver", cfdl.SOLVER_MCSGS)
            if nccl:
                if rank == 0:
                    handles[0] = cfdl.comm_unique_id()
                bar.wait()
                s.comm_init(handles[0])
            else:
                handles[rank] = s.ipc_handle()
                bar.wait()
                s.ipc_connect(handles)
            bar.wait()
            its = drive(s, tc, cp, bcv)
            fields = {}
            for f in wanted:
                a = np.full(s.field_size(f), np.nan)
                s.download_into(f, a)
                fields[f] = a
            out[rank] = (its, fields)
            bar.wait()
            s.close()
        except Exception as ex:  # a failing rank must not leave the others at the barrier
            errs.append((rank, repr(ex)))
            bar.abort()

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    one = cfdl.Solver(geom, bcs, device=0)
    one.set_option("solver", cfdl.SOLVER_MCSGS)
    want_its = drive(one, tc, cp, bcv)
    for r in range(world):
        for (hs, he), (ws, we) in zip(out[r][0], want_its):
            assert hs[0] == ws[0] and he[0] == we[0], (r, hs, ws, he, we)
            assert abs(hs[2] - ws[2]) <= 1e-10 * abs(ws[2]) + 1e-300 and abs(he[2] - we[2]) <= 1e-10 * abs(we[2]) + 1e-300, (r, hs, ws, he, we)
    worst = 0.0
    for f in wanted:
        m = np.full_like(out[0][1][f], np.nan)
        for r in range(world):
            ok = ~np.isnan(out[r][1][f])
            m[ok] = out[r][1][f][ok]
        assert not np.isnan(m).any(), f + ": some entries were reported by no rank"
        w = one.download(f)
        err = np.abs(m - w).max() / max(np.abs(w).max(), 1e-300)
        assert err < 1e-12, (f, err)
        worst = max(worst, err)
    one.close()
    print("multirank transport ok: world=%d n=%d %s worst field err %.2e, scalar / energy iterations %s"
          % (world, n, how, worst, [(int(a[0]), int(b[0])) for a, b in want_its]))


if __name__ == "__main__":
    main()
