# Round 2, GPU call 23 (1 GPU): quotients from reciprocals in the persistent pc solve (quot<true>), bit comparison with the pass-by-pass kernels on hardware, A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -x -q -k "persistent or fast_mode or hex128" > gpurun_out/r02c23_pytest.log 2>&1; tail -3 gpurun_out/r02c23_pytest.log
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 3 "$@" > gpurun_out/r02c23_bench_$name.json 2> gpurun_out/r02c23_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c23_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    c=d["config"]
    print("$name", "value %.1fM ms/step %.3f" % (d["value"]/1e6, d["ms_per_step"]), "pass_us %.2f frac %.3f" % (r.get("avg_launch_ms",0)*1e3, r.get("frac",0)), c["solver_iterations_last_step(u,v,w,pc)"], "sgs %.3f" % d["phase_ms_per_step"].get("sgs", 0))
except Exception as e: print("$name ERR", e)
PY
}
run fast_a
run fast_pf0 --opt rbq_prefetch=0
run fast_b
run fast_ctas4 --opt rbq_ctas=4
run 160_fast --size 160 --steps 6
run 144_fast --size 144 --steps 6
