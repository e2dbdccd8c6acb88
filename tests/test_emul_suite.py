"""The gpu-marked parity tests, executed on the host against the cuemu build of the CUDA sources.

tests/emul/ compiles cfd-lite_b200/csrc/*.cu for the host (kernel launches rewritten, CUDA
runtime and device intrinsics emulated with fibers, guard pages behind every device allocation)
into tests/emul/_build/libcfdl_emul.so.  Running the GPU parity suite against it checks, in a
container without a GPU, that the kernels' arithmetic and indexing and the orchestration in
api.cu / kernels_solver.cu reproduce the oracle — it is a checker for the sources, not a way
to run the product: the product library has no CPU path and nothing outside tests/ can load the
emulated one.  Timing, occupancy and memory-ordering across CTAs are of course NOT covered; the
`-m gpu` run on the B200 remains the parity gate.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_suite_passes_under_emulation():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "--emul", "-x", "-q",
                        "-p", "no:cacheprovider"], cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, "gpu suite under emulation failed:\n%s\n%s" % (tail, r.stderr[-2000:])
    assert " passed" in tail


def test_bench_script_runs_under_emulation():
    """Plumbing check of bench.py (argument handling, step loop, e2e through cfdl_step_host, JSON line) on a
    tiny mesh against the cuemu build; the numbers are host-emulation artefacts and are not looked at."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "run_bench_emul.py"), "--size", "8", "--steps", "3", "--warmup", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["e2e"]["path"].startswith("cfdl_step_host") and line["e2e"]["path_error"] is None
    assert line["gpu_launches"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_medium_mesh_runs_under_emulation():
    """Grids of tens of CTAs with several rows per thread (tests/emul/medium_check.py): exact mode equals the
    oracle; the throughput mode as shipped equals the same mode with every optimised path switched off, bit for bit."""
    import os as _os
    for seed in ("3",):
        env = dict(_os.environ, CUEMU_RANDOM_TIMES=seed)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "medium_check.py"), "16"], cwd=ROOT, env=env,
                           capture_output=True, text=True, timeout=1200)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        assert "medium emulation ok" in r.stdout
