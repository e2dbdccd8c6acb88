# Round 2, GPU call 9 (1 GPU): hoisted connectivity loads in the assembly kernels; persistent pc solve on a 16.8 M cell mesh
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_oracle_vs_reference_source.py -m gpu -x -q > gpurun_out/r02c9_pytest.log 2>&1; tail -3 gpurun_out/r02c9_pytest.log
run() { name=$1; shift; timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02c9_bench_$name.json 2> gpurun_out/r02c9_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c9_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1fM ms/step %.3f pass_us %.2f frac %.3f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"]*1e3, d["roofline"]["frac"]), d["config"].get("pc_solve","")[:30], {k: round(v, 3) for k, v in d["phase_ms_per_step"].items()}, d["config"]["solver_iterations_last_step(u,v,w,pc)"])
except Exception as e: print("$name ERR", e)
PY
}
run n128
run n128_miphoist0 --opt mip_hoist=0
run n256 --size 256 --structured
run n256_rbq0 --size 256 --structured --opt rbq=0
run n200 --size 200
run n200_rbq0 --size 200 --opt rbq=0
