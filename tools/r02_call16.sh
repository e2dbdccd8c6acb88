# Round 2, GPU call 16 (1 GPU): persistent pc solve with chunks handed out in order from a counter (large meshes): on/off, chunk sizes
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -x -q -k "persistent or restart" > gpurun_out/r02c16_pytest.log 2>&1; tail -3 gpurun_out/r02c16_pytest.log
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 6 --warmup 3 "$@" > gpurun_out/r02c16_bench_$name.json 2> gpurun_out/r02c16_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c16_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    c=d["config"]
    print("$name", "value %.1fM ms/step %.3f" % (d["value"]/1e6, d["ms_per_step"]), "pass_us %.2f frac %.3f" % (r.get("avg_launch_ms",0)*1e3, r.get("frac",0)), c.get("pc_solve_chunks"), c["solver_iterations_last_step(u,v,w,pc)"], "sgs %.3f" % d["phase_ms_per_step"].get("sgs", 0))
except Exception as e: print("$name ERR", e)
PY
}
run 256_off --size 256 --opt rbq_rounds=0
run 256_lbig2048 --size 256
CFDL_RBQ_LBIG=1024 run 256_lbig1024 --size 256
CFDL_RBQ_LBIG=4096 run 256_lbig4096 --size 256
run 200_off --size 200 --opt rbq_rounds=0
run 200_lbig2048 --size 200
run 160_off --size 160 --opt rbq_rounds=0
run 160_on --size 160
run 128_default --steps 12
CFDL_RBQ_LBIG=4096 run 200_lbig4096 --size 200
CFDL_RBQ_LBIG=1024 run 200_lbig1024 --size 200
CFDL_RBQ_LBIG=1024 run 160_lbig1024 --size 160
CFDL_RBQ_LMAX=1024 CFDL_RBQ_LBIG=1024 run 128_tickets1024 --steps 12
CFDL_RBQ_LMAX=1024 CFDL_RBQ_LBIG=2048 run 128_tickets2048 --steps 12
