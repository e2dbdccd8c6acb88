"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU, one
RCB block per process, ghost exchange over NCCL.  Because the colouring is global and ghosts
are refreshed after every colour sweep, the partitioned run must reproduce the single-GPU
multicolour run: identical iteration counts, fields equal to 1e-12 (residual norms differ only
by the summation order of the all-reduce)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


STRUCTURED = 2  # "kind" of the cases built with cfdl_create_structured_hex
SLABS = 3       # hex mesh cut into z-slabs (general builder): the persistent pc solve with chunk-to-chunk synchronisation over NVLink
STRUCTURED_SLABS = 4  # cfdl_create_structured_hex_slabs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _transport_steps(s, tc, cp, bcv):
    """energy + scalar equations next to uvwp: three time steps; returns the (scalar, energy) solve records"""
    s.energy_init(tc=tc, cp=cp)
    s.scalar_init(dcoef=0.7, vel=(3.0, -2.0, 5.0), bc_value=bcv)
    recs = []
    for _ in range(3):
        s.update_boundaries()
        s.solve_uvwp(0.01, 30)
        hs = s.solve_scalar(0.01, 100)
        he = s.solve_energy(50.0, 100)  # (a long time step: several iterations)
        s.update_time()
        recs.append((hs, he))
    return recs


def _worker(rank, world, port, kind, n, p2p, out_q, transport=False):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
    import cfdl
    try:
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        structured = kind in (STRUCTURED, STRUCTURED_SLABS)
        slabs = kind in (SLABS, STRUCTURED_SLABS)
        if structured or slabs:  # per-rank analytic generator; the reference run below uses the general builder
            kind = 0
        raw = cfdl.meshgen(kind, n, jitter=0.2 if kind else 0.0, shuffle=bool(kind))
        geom = cfdl.mesh_build(raw)
        bcs = cfdl.default_bcs(raw)
        if structured:
            s = cfdl.Solver.structured_hex(n, device=rank, rank=rank, nranks=world, slabs=slabs)
        else:
            c2r, _, _ = cfdl.partition_rcb(geom, world)
            if slabs:
                k = np.arange(n ** 3) // (n * n)
                c2r = np.zeros(n ** 3, np.int32)
                for r in range(world):
                    c2r[(k >= n * r // world) & (k < n * (r + 1) // world)] = r + 1
            s = cfdl.Solver(geom, bcs, device=rank, cell2rank=c2r, rank=rank, nranks=world)
        ids = [cfdl.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        s.comm_init(ids[0])
        s.set_option("solver", cfdl.SOLVER_MCSGS)
        if p2p:  # peer-to-peer ghost exchange through CUDA IPC instead of NCCL send/recv
            handles = [None] * world
            dist.all_gather_object(handles, s.ipc_handle())
            s.ipc_connect(handles)
        hist = s.run(dt=0.01, nit=100, ntstep=2, ncoef=3)
        if slabs and p2p:
            assert int(s.get_info("rbq_dist")) == 1, "the persistent partitioned pc solve is not in use"
        recs = None
        if transport:  # per-cell properties in the reference numbering, the same on every rank
            rng = np.random.default_rng(5)
            ne = int(raw["ne"])
            tc = 5.0 * (1.0 + 0.3 * rng.random(ne))
            cp = 1000.0 * (1.0 + 0.2 * rng.random(ne))
            bcv = rng.integers(0, 2, len(bcs[1])).astype(np.float64)
            recs = _transport_steps(s, tc, cp, bcv)
        fields = {}
        for f in ("u", "v", "w", "p", "gp") + (() if structured else ("mip",)) + (("t", "h", "s", "gt", "gh", "gs") if transport else ()):  # structured: own face numbering
            a = np.full(s.field_size(f), np.nan)
            s.download_into(f, a)  # writes only the entries this rank owns
            fields[f] = a
        gathered = [None] * world
        dist.all_gather_object(gathered, fields)
        if rank == 0:
            merged = {}
            for f in fields:
                m = np.full_like(fields[f], np.nan)
                for g in gathered:
                    ok = ~np.isnan(g[f])
                    m[ok] = g[f][ok]
                assert not np.isnan(m).any(), f + ": some entries were reported by no rank"
                merged[f] = m
            one = cfdl.Solver(geom, bcs, device=0)
            one.set_option("solver", cfdl.SOLVER_MCSGS)
            want_hist = one.run(dt=0.01, nit=100, ntstep=2, ncoef=3)
            assert np.array_equal(hist[:, :, 0], want_hist[:, :, 0]), (hist[:, :, 0], want_hist[:, :, 0])
            err_h = np.abs(hist[:, :, 1:3] - want_hist[:, :, 1:3]).max() / np.abs(want_hist[:, :, 1:3]).max()
            assert err_h < 1e-10, err_h
            if transport:
                want = _transport_steps(one, tc, cp, bcv)
                for (hs, he), (ws, we) in zip(recs, want):
                    assert hs[0] == ws[0] and he[0] == we[0], (hs, ws, he, we)
            for f in merged:
                w = one.download(f)
                err = np.abs(merged[f] - w).max() / max(np.abs(w).max(), 1e-300)
                assert err < 1e-12, (f, err)
            one.close()
        s.close()
        dist.barrier()
        dist.destroy_process_group()
        out_q.put((rank, "ok"))
    except Exception as ex:
        import traceback
        out_q.put((rank, "FAIL: " + "".join(traceback.format_exception(type(ex), ex, ex.__traceback__))))


@pytest.mark.parametrize("kind,n,p2p", [(0, 10, False), (1, 5, False), (0, 10, True), (0, 24, True), (STRUCTURED, 12, False), (STRUCTURED, 16, True), (1, 5, True),
                                          (SLABS, 24, True), (SLABS, 48, True), (STRUCTURED_SLABS, 32, True), (SLABS, 12, False)])
def test_partitioned_run_equals_single_gpu(cfdl, kind, n, p2p):
    _run_ranks(cfdl, kind, n, p2p, False)


@pytest.mark.parametrize("kind,n,p2p", [(0, 10, True), (0, 10, False), (SLABS, 16, True)])
def test_partitioned_energy_and_scalar_equations_equal_single_gpu(cfdl, kind, n, p2p):
    """cfdl_solve_energy / cfdl_solve_scalar on partitioned handles (kernels_transport.cu): t, phi, s and their gradients equal the
    single-GPU fields, same iteration counts — peer-to-peer (the solve runs on a slab-resident work array) and NCCL exchange."""
    _run_ranks(cfdl, kind, n, p2p, True)


def _run_ranks(cfdl, kind, n, p2p, transport):
    world = min(cfdl.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    if kind == STRUCTURED and world == 3:
        world = 2  # the structured generator bisects: power-of-two rank counts
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n, p2p, q, transport)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    bad = [r for r in results if r[1] != "ok"]
    assert not bad, bad
