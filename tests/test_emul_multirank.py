"""Partitioned peer-to-peer runs on the cuemu build (ranks = threads of one process, see
tests/emul/multirank_check.py): 2, 4 and 8 ranks must reproduce the single-rank fields, like
tests/test_gpu_multi.py asserts on real GPUs.  Covers the interface-CTA stores, flag words,
staged exchanges and the mailbox all-reduce on the host; NVLink ordering itself is only
exercised by the gpu test."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,n,structured", [(4, 10, False), (8, 12, False), (4, 12, True), (4, 10, "pcg"), (4, 10, "pcg-ssor"), (3, 4, "tet"),
                                                (4, 12, "slabs"), (3, 10, "slabs"), (4, 12, "structured-slabs")])
def test_partitioned_p2p_run_equals_single_rank_under_emulation(world, n, structured):
    extra = [structured] if isinstance(structured, str) else (["structured"] if structured else [])
    cmd = [sys.executable, os.path.join(ROOT, "tests", "emul", "multirank_check.py"), str(world), str(n)] + extra
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "multirank emulation ok" in r.stdout


@pytest.mark.parametrize("world,n,mode", [(4, 10, "nccl"), (3, 4, "nccl-tet"), (4, 10, "nccl-pcg"), (3, 8, "nccl-pcg-ssor")])
def test_partitioned_nccl_mode_equals_single_rank_under_emulation(world, n, mode):
    """The library's NCCL exchange mode (grouped send/recv of ghost values per colour, all-reduced residual norms and
    dot products, broadcast of pc(1)) against an in-process stand-in for NCCL (tests/emul/fake_nccl.cpp): hex mesh,
    tet mesh with more than two colours, conjugate gradients."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "multirank_check.py"), str(world), str(n), mode], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "multirank emulation ok" in r.stdout


@pytest.mark.parametrize("fused", ["0", "1"])
def test_partitioned_momentum_solves_side_by_side_and_one_by_one(fused):
    """4 ranks, a time step large enough for different iteration counts of u, v, w: the side-by-side passes with
    peer-to-peer stores of all three equations (and the one-by-one solves) reproduce the single-rank run."""
    env = dict(os.environ, CFDL_TEST_UVW_FUSED=fused, CFDL_TEST_DT="5.0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "multirank_check.py"), "4", "12"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "multirank emulation ok" in r.stdout


@pytest.mark.parametrize("world,n", [(2, 16), (4, 16)])
def test_partitioned_persistent_pc_solve_with_chunks_from_a_counter(world, n):
    """The partitioned persistent pc solve with more chunks than co-resident CTAs (chunks of 64 rows handed out in order from a
    counter, interface chunks first, tagged 128-bit stores across the ranks), forced on small slabs: same fields as one rank."""
    env = dict(os.environ, CFDL_RBQ_LMAX="64", CFDL_RBQ_LBIG="64", CFDL_RBQ_LS_MIN="64", CFDL_TEST_RBQ_ROUNDS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "multirank_check.py"), str(world), str(n), "slabs"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "multirank emulation ok" in r.stdout


@pytest.mark.parametrize("world,n,how", [(4, 10, "p2p"), (3, 8, "nccl"), (4, 12, "slabs")])
def test_partitioned_energy_and_scalar_equations_equal_single_rank(world, n, how):
    """cfdl_solve_energy / cfdl_solve_scalar on partitioned handles (tc / cp with ghost cells, ghost exchange of t, phi and the
    gradients, the solve on a slab-resident work array in peer-to-peer mode): fields equal to the single-rank run, same
    iteration counts (tests/emul/multirank_transport_check.py)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "multirank_transport_check.py"), str(world), str(n), how], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "multirank transport ok" in r.stdout


@pytest.mark.parametrize("mode", ["rcb", "slabs"])
def test_silent_rank_is_an_error_code_not_a_hang(mode):
    """Every wait on a word another rank writes is time-limited: a rank that connects and then never launches surfaces on
    its neighbour as CFDL_ERR_COMM after the limit (tests/emul/silent_rank_check.py), with the pass-by-pass and the persistent
    form of the pc solve alike."""
    cmd = [sys.executable, os.path.join(ROOT, "tests", "emul", "silent_rank_check.py")] + (["slabs"] if mode == "slabs" else [])
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "silent rank ok" in r.stdout
