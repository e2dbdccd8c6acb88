"""Multi-GPU micro-benchmark of the pc solve (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/tune_multi.py [size]
Times cfdl_solve_eq(pc, nit=100) from phi=0 on a fixed system with the exchange modes p2p / nccl.
CFDL_P2P_DEBUG (1 no waits, 2 no remote stores, 4 rank-local residual) isolates the cost of each part
of the peer-to-peer protocol (results are wrong with any bit set; timing only)."""
import os
import sys

sys.path.insert(0, "cfd-lite_b200/python")
os.environ["NCCL_DEBUG"] = "NONE"
import numpy as np
import torch
import torch.distributed as dist
import cfdl

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(round(size * world ** (1.0 / 3.0)))
raw = cfdl.meshgen(0, n)
geom = cfdl.mesh_build(raw)
c2r, _, _ = cfdl.partition_rcb(geom, world, want_order=False)
s = cfdl.Solver(geom, cfdl.default_bcs(raw), device=lr, cell2rank=c2r, rank=rank, nranks=world)
ids = [cfdl.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
s.comm_init(ids[0])
handles = [None] * world
dist.all_gather_object(handles, s.ipc_handle())
s.ipc_connect(handles)
s.set_option("solver", 1)
for i in range(2):
    s.update_boundaries()
    s.solve_uvwp()
s.update_boundaries()
s.calc_coef_uvw()
for eq in (0, 1, 2):
    s.solve_eq(eq)
for a, b in (("u", "gu"), ("v", "gv"), ("w", "gw")):
    s.calc_grad(a, b)
s.calc_mip(True)
s.calc_coef_p()
zeros = np.zeros(s.H)
for mode in ("p2p", "nccl"):
    s.set_option("p2p", 1 if mode == "p2p" else 0)
    for rep in range(2):
        s.upload("pc", zeros)
        dist.barrier()
        s.timer_record(0)
        out = s.solve_eq(3, 100)
        s.timer_record(1)
        ms = s.timer_elapsed_ms(0, 1)
    # same solve with per-launch events: kernel time alone (gaps between launches excluded)
    s.upload("pc", zeros)
    s.set_option("profile", 1)
    s.set_option("reset_counters", 1)
    dist.barrier()
    import time
    t0 = time.perf_counter()
    out2 = s.solve_eq(3, 100)
    wall = (time.perf_counter() - t0) * 1e3
    kms, kn = s.get_info("prof_ms_sgs"), s.get_info("prof_n_sgs")
    s.set_option("profile", 0)
    if rank == 0:
        print("%s world=%d n=%d dbg=%s: it=%d solve %.3f ms -> %.1f us per iteration (owned %d, ghosts %d); "
              "profiled: %d passes, kernel time %.3f ms = %.1f us per pass, wall %.3f ms" % (
                  mode, world, n, os.environ.get("CFDL_P2P_DEBUG", "0"), out[0], ms, 1e3 * ms / max(out[0], 1),
                  int(s.get_info("owned_cells")), int(s.get_info("ghost_cells")), kn, kms, 1e3 * kms / max(kn, 1), wall), flush=True)
dist.barrier()
s.close()
dist.destroy_process_group()
