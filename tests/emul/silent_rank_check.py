#!/usr/bin/env python3
"""A rank that never launches (a dead process / a GPU that hung elsewhere) must surface as an error code on the ranks that
wait for it, not as a hang: every wait on a word another rank writes is time-limited (state.h: spin_until_ge, the persistent
pc solve's tagged slots), the waiter raises the handle's error word, later waits return at once, the kernels run to their end
and the API call reports CFDL_ERR_COMM.  Ranks are threads of this process on the cuemu build (tests/emul); rank 1 connects
and then stays silent.  TEST INFRASTRUCTURE ONLY.  usage: silent_rank_check.py [slabs]"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cfdl  # noqa: E402
import conftest  # noqa: E402


def main():
    slabs = len(sys.argv) > 1 and sys.argv[1] == "slabs"
    conftest.use_emulated_library()
    n, world = 10, 2
    raw = cfdl.meshgen(0, n)
    geom = cfdl.mesh_build(raw)
    bcs = cfdl.default_bcs(raw)
    c2r, _, _ = cfdl.partition_rcb(geom, world)
    if slabs:
        k = np.arange(n ** 3) // (n * n)
        c2r = np.where(k < n // 2, 1, 2).astype(np.int32)
    bar = threading.Barrier(world)
    handles = [None] * world
    result = {}

    def rank_main(rank):
        s = cfdl.Solver(geom, bcs, device=0, cell2rank=c2r, rank=rank, nranks=world)
        s.set_option("solver", cfdl.SOLVER_MCSGS)
        handles[rank] = s.ipc_handle()
        bar.wait()
        s.ipc_connect(handles)
        bar.wait()
        if rank == 1:
            bar.wait()  # silent: never computes; keeps its slab alive until rank 0 has given up
            s.close()
            return
        t0 = time.time()
        try:
            s.run(dt=0.01, nit=100, ntstep=1, ncoef=1)
            result["outcome"] = "returned normally"
        except cfdl.CfdlError as ex:
            result["outcome"] = "error"
            result["code"] = ex.code if hasattr(ex, "code") else None
            result["msg"] = str(ex)
        result["seconds"] = time.time() - t0
        bar.wait()
        s.close()

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert result.get("outcome") == "error", result
    assert "time limit" in result["msg"] or "timed out" in result["msg"], result
    assert result["seconds"] < 120, result
    print("silent rank ok: rank 0 got '%s' after %.1f s" % (result["msg"][:90], result["seconds"]))


if __name__ == "__main__":
    main()
