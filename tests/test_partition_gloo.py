"""Host-side logic of the multi-GPU path, exercised on CPU with world_size-2 (and 3) `gloo`
process groups: every rank derives its partition plan (owned cells, ghost cells, send/receive
lists) from the global mesh without communication; the test then checks across ranks that the
plans tile the mesh, that each sender's list equals the receiver's ghost slice element for
element, and that a halo exchange driven by those lists reproduces the global field — the
property update_halos (src/modules/mod_subdomains.f90:191-212) has in the reference.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, n, out_q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
    import cfdl
    try:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        raw = cfdl.meshgen(kind, n, jitter=0.2 if kind else 0.0, shuffle=bool(kind))
        geom = cfdl.mesh_build(raw)
        c2r, _, _ = cfdl.partition_rcb(geom, world)
        plan = cfdl.partition_plan(geom, c2r, world, rank)
        ne = geom["ne"]
        plans = [None] * world
        dist.all_gather_object(plans, {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in plan.items()})
        # 1. owned sets tile the mesh and agree with cell2rank
        allowned = np.concatenate([np.array(p["owned"]) for p in plans])
        assert sorted(allowned.tolist()) == list(range(1, ne + 1))
        assert np.all(c2r[plan["owned"] - 1] == rank + 1)
        # owned cells are colour-major and the colouring is consistent across ranks
        assert plan["color_ptr"][-1] == len(plan["owned"]) and all(p["ncolors"] == plan["ncolors"] for p in plans)
        # 2. sender list == receiver ghost slice, element for element
        for i, r in enumerate(plan["nbr_rank"]):
            mine = plan["send_cells"][plan["send_ptr"][i]:plan["send_ptr"][i + 1]]
            other = plans[r]
            j = other["nbr_rank"].index(rank)
            theirs = np.array(other["ghost"][other["recv_ptr"][j]:other["recv_ptr"][j + 1]])
            assert np.array_equal(mine, theirs), (rank, r)
            assert np.all(c2r[mine - 1] == rank + 1)
        # 3. every face-neighbour of an owned cell is owned or ghost
        known = set(plan["owned"].tolist()) | set(plan["ghost"].tolist())
        idx = geom["ef2nb_idx"].astype(np.int64)
        nb = geom["ef2nb_nb"].astype(np.int64)
        for e in plan["owned"]:
            for k in range(idx[e - 1] - 1, idx[e] - 1):
                if nb[k] & 31:
                    assert (nb[k] >> 5) in known
        # 4. halo exchange with gloo send/recv driven by the lists reproduces the global field
        field = np.sin(np.arange(1, ne + 1) * 0.37) + 2.0
        local = np.full(len(plan["owned"]) + len(plan["ghost"]), np.nan)
        local[:len(plan["owned"])] = field[plan["owned"] - 1]
        o2l = {int(g): i for i, g in enumerate(plan["owned"])}
        reqs, recvs = [], []
        for i, r in enumerate(plan["nbr_rank"]):
            cells = plan["send_cells"][plan["send_ptr"][i]:plan["send_ptr"][i + 1]]
            buf = torch.tensor(local[[o2l[int(c)] for c in cells]])
            reqs.append(dist.isend(buf, dst=int(r)))
            rb = torch.empty(int(plan["recv_ptr"][i + 1] - plan["recv_ptr"][i]), dtype=torch.float64)
            recvs.append((i, rb, dist.irecv(rb, src=int(r))))
        for q in reqs:
            q.wait()
        for i, rb, q in recvs:
            q.wait()
            local[len(plan["owned"]) + plan["recv_ptr"][i]:len(plan["owned"]) + plan["recv_ptr"][i + 1]] = rb.numpy()
        assert np.array_equal(local[len(plan["owned"]):], field[plan["ghost"] - 1])
        # 5. a global reduction (what the residual norm needs)
        t = torch.tensor([float(np.sum(local[:len(plan["owned"])] ** 2))], dtype=torch.float64)
        dist.all_reduce(t)
        assert abs(t.item() - float(np.sum(field ** 2))) < 1e-9 * float(np.sum(field ** 2))
        dist.barrier()
        dist.destroy_process_group()
        out_q.put((rank, "ok"))
    except Exception as ex:  # report instead of hanging the peers
        import traceback
        out_q.put((rank, "FAIL: " + "".join(traceback.format_exception(type(ex), ex, ex.__traceback__))))


@pytest.mark.parametrize("world,kind,n", [(2, 0, 6), (2, 1, 3), (3, 0, 5)])
def test_partition_plans_and_halo_exchange(world, kind, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    for _ in range(world):
        results.append(q.get(timeout=180))
    for p in procs:
        p.join(timeout=60)
    bad = [r for r in results if r[1] != "ok"]
    assert not bad, bad


@pytest.mark.parametrize("world,kind,n", [(4, 0, 8), (8, 0, 8), (8, 1, 4), (4, 0, 9)])
def test_plans_of_4_and_8_ranks_match_pairwise(cfdl, world, kind, n):
    """The peer-to-peer exchange addresses a neighbour's ghost slice by (neighbour index,
    offset), so what one rank sends must equal, element for element and per colour, what the
    other expects — checked here for every pair of the 4- and 8-rank partitions (one process,
    no communication: every plan is derived from the global mesh alone)."""
    raw = cfdl.meshgen(kind, n, jitter=0.2 if kind else 0.0, shuffle=bool(kind))
    geom = cfdl.mesh_build(raw)
    c2r, _, _ = cfdl.partition_rcb(geom, world, want_order=False)
    plans = [cfdl.partition_plan(geom, c2r, world, r) for r in range(world)]
    ne = geom["ne"]
    assert sorted(np.concatenate([p["owned"] for p in plans]).tolist()) == list(range(1, ne + 1))
    assert len({p["ncolors"] for p in plans}) == 1
    for r, p in enumerate(plans):
        nbrs = [int(x) for x in p["nbr_rank"]]
        assert nbrs == sorted(nbrs) and r not in nbrs and len(nbrs) <= 8
        for i, q in enumerate(nbrs):
            other = plans[q]
            j = [int(x) for x in other["nbr_rank"]].index(r)  # symmetric neighbour sets
            mine = p["send_cells"][p["send_ptr"][i]:p["send_ptr"][i + 1]]
            theirs = other["ghost"][other["recv_ptr"][j]:other["recv_ptr"][j + 1]]
            assert np.array_equal(mine, theirs), (r, q)
