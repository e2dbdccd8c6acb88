#!/usr/bin/env python3
"""Partitioned run on the cuemu build: the ranks of a multi-GPU job are THREADS of this process,
every rank owns a handle of the emulated library, and the "IPC handle" of a slab is its address —
so the peer-to-peer path of comm.cu / kernels_rb.inc (interface CTAs storing into the neighbours'
ghost cells, flag words, mailbox all-reduce) runs for real, concurrently, on the host.  The merged
result must equal the single-rank run like in tests/test_gpu_multi.py.  TEST INFRASTRUCTURE ONLY.
usage: multirank_check.py <world> <n> [structured|slabs|structured-slabs|pcg|tet|nccl|nccl-tet|nccl-pcg]
(slabs: z-slab partition — the persistent pc solve with chunk-to-chunk synchronisation across ranks must be in use)
(nccl*: the library's NCCL exchange mode against tests/emul/fake_nccl.cpp instead of the peer-to-peer slabs)
"""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cfdl  # noqa: E402
import conftest  # noqa: E402


def main():
    world, n = int(sys.argv[1]), int(sys.argv[2])
    structured = len(sys.argv) > 3 and sys.argv[3] in ("structured", "structured-slabs", "tall-slabs")
    slabs = len(sys.argv) > 3 and sys.argv[3] in ("slabs", "structured-slabs", "tall-slabs")
    nz = (n * world) // 2 + 1 if (len(sys.argv) > 3 and sys.argv[3] == "tall-slabs") else None  # n x n x nz cells (cfdl_create_structured_hex_slabs)
    pcg = len(sys.argv) > 3 and sys.argv[3] in ("pcg", "nccl-pcg", "pcg-ssor", "nccl-pcg-ssor")
    precond = 1 if (pcg and sys.argv[3].endswith("ssor")) else 0  # multicolour-SSOR (colour-wise ghost exchanges inside the preconditioner) / Jacobi
    nccl = len(sys.argv) > 3 and sys.argv[3].startswith("nccl")
    tet = len(sys.argv) > 3 and sys.argv[3] in ("nccl-tet", "tet")  # "tet": peer-to-peer, one staged exchange per colour
    if nccl:
        os.environ["CFDL_NCCL_PATH"] = os.path.join(ROOT, "tests", "emul", "_build", "libnccl_emul.so")  # conjugate gradients for pc: ghost exchange of p + all-reduced dot products
    mode = cfdl.SOLVER_PCG if pcg else cfdl.SOLVER_MCSGS
    tol_f, tol_h = (1e-8, 1e-8) if pcg else (1e-12, 1e-10)  # the dot products are summed per rank, then in rank order
    conftest.use_emulated_library()
    # optional: pin the momentum solves side by side (1) / one by one (0) on the ranks, and a time step large enough for
    # the momentum equations to need several, different iteration counts
    fused = os.environ.get("CFDL_TEST_UVW_FUSED")
    dt = float(os.environ.get("CFDL_TEST_DT", "0.01"))
    raw = cfdl.meshgen(1, n, jitter=0.2, shuffle=True) if tet else cfdl.meshgen(0, n)
    geom = cfdl.mesh_build(raw)
    bcs = cfdl.default_bcs(raw)
    c2r, _, _ = cfdl.partition_rcb(geom, world)
    if slabs:  # z-slabs like cfdl_create_structured_hex_slabs: cell id = i + n (j + n k)
        k = np.arange(n ** 3) // (n * n)
        c2r = np.zeros(n ** 3, np.int32)
        for r in range(world):
            c2r[(k >= n * r // world) & (k < n * (r + 1) // world)] = r + 1
    bar = threading.Barrier(world)
    handles = [None] * world
    out = [None] * world
    errs = []
    fields_wanted = ("u", "v", "w", "p", "gp") + (() if structured else ("mip",))

    def rank_main(rank):
        try:
            if structured:
                s = cfdl.Solver.structured_hex(n, device=0, rank=rank, nranks=world, slabs=slabs, nz=nz)
            else:
                s = cfdl.Solver(geom, bcs, device=0, cell2rank=c2r, rank=rank, nranks=world)
            s.set_option("solver", mode)
            if pcg:
                s.set_option("pcg_precond", precond)
            if fused is not None:
                s.set_option("uvw_fused", int(fused))
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
            hist = s.run(dt=dt, nit=100, ntstep=2, ncoef=2)
            fields = {}
            for f in fields_wanted:
                a = np.full(s.field_size(f), np.nan)
                s.download_into(f, a)
                fields[f] = a
            out[rank] = (hist, fields)
            if slabs:
                assert int(s.get_info("rbq_dist")) == 1, ("persistent partitioned pc solve not in use", rank, s.get_info("rbq_dist"))
                if os.environ.get("CFDL_TEST_RBQ_ROUNDS"):  # more chunks than CTAs, handed out from a counter
                    assert int(s.get_info("rbq_chunks")) > int(s.get_info("rbq_grid")), (rank, s.get_info("rbq_chunks"), s.get_info("rbq_grid"))
            bar.wait()
            # cfdl_step_host with partition-local arrays (bench.py's e2e path on several GPUs) against the
            # separate local upload / solve / download calls, from the same state: same bits on every rank
            ins = "u v w p u0 v0 w0 gu gv gw gp mip mip0".split()
            outs = "u v w p gu gv gw gp gpc mip".split()
            state = {k: s.download_local(k, np.zeros(s.local_size(k))) for k in ins}
            s.update_boundaries()
            h_sep = s.solve_uvwp(0.01, 30)
            want = {k: s.download_local(k, np.zeros(s.local_size(k))) for k in outs}
            bar.wait()
            got = {k: np.zeros(s.local_size(k)) for k in outs}
            h_one = s.step_host({k: state[k] for k in ins}, got, dt=0.01, nit=30, apply_bcs=True, local=True)
            assert np.array_equal(h_sep[:, 0], h_one[:, 0]), (rank, h_sep, h_one)
            for k in outs:
                assert np.array_equal(got[k], want[k]), (rank, k)
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
    one = cfdl.Solver(geom, bcs, device=0) if nz is None else cfdl.Solver.structured_hex(n, device=0, slabs=True, nz=nz)
    one.set_option("solver", mode)
    if pcg:
        one.set_option("pcg_precond", precond)
    want_hist = one.run(dt=dt, nit=100, ntstep=2, ncoef=2)
    for r in range(world):
        hist = out[r][0]
        assert np.array_equal(hist[:, :, 0], want_hist[:, :, 0]), (r, hist[:, :, 0], want_hist[:, :, 0])
        err_h = np.abs(hist[:, :, 1:3] - want_hist[:, :, 1:3]).max() / np.abs(want_hist[:, :, 1:3]).max()
        assert err_h < tol_h, err_h
    worst = 0.0
    for f in fields_wanted:
        m = np.full_like(out[0][1][f], np.nan)
        for r in range(world):
            ok = ~np.isnan(out[r][1][f])
            m[ok] = out[r][1][f][ok]
        assert not np.isnan(m).any(), f + ": some entries were reported by no rank"
        w = one.download(f)
        err = np.abs(m - w).max() / max(np.abs(w).max(), 1e-300)
        assert err < tol_f, (f, err)
        worst = max(worst, err)
    one.close()
    print("multirank emulation ok: world=%d n=%d structured=%s pcg=%s worst field err %.2e, pc iterations %s, momentum iterations %s"
          % (world, n, structured, pcg, worst, want_hist[:, 3, 0].astype(int).tolist(), want_hist[:, :3, 0].astype(int).tolist()))


if __name__ == "__main__":
    main()
