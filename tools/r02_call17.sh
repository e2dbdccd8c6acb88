# Round 2, GPU call 17 (1 GPU): why does the persistent pc solve lose on meshes whose value arrays exceed the L2?  ncu of rbq_kernel (chunks from a counter) and of the pass-by-pass kernels at 200^3; band-width experiments
set -u
mkdir -p gpurun_out
NB="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --size 200"
timeout 600 ncu --set full --clock-control none -k regex:"rbq_kernel" --launch-skip 3 -c 1 -o gpurun_out/r02c17_rbq_200 $NB > gpurun_out/r02c17_ncu_rbq.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"rb_red_kernel|rb_black_kernel" --launch-skip 300 -c 2 -o gpurun_out/r02c17_rb_200 $NB --opt rbq_rounds=0 > gpurun_out/r02c17_ncu_rb.log 2>&1
ls -la gpurun_out/r02c17_*.ncu-rep | awk '{print $5, $9}'
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 6 --warmup 3 "$@" > gpurun_out/r02c17_bench_$name.json 2> gpurun_out/r02c17_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c17_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    c=d["config"]
    print("$name", "value %.1fM ms/step %.3f" % (d["value"]/1e6, d["ms_per_step"]), "pass_us %.2f frac %.3f" % (r.get("avg_launch_ms",0)*1e3, r.get("frac",0)), c.get("pc_solve_chunks"), c["solver_iterations_last_step(u,v,w,pc)"], "sgs %.3f" % d["phase_ms_per_step"].get("sgs", 0))
except Exception as e: print("$name ERR", e)
PY
}
run 200_cap148 --size 200 --opt rbq_cap=148
run 200_cap148_l512 --size 200 --opt rbq_cap=148 --opt rbq_lbig=512
run 200_cap296 --size 200 --opt rbq_cap=296
run 176_on --size 176
run 176_off --size 176 --opt rbq_rounds=0
run 144_on --size 144
run 144_off --size 144 --opt rbq_rounds=0
CFDL_RBQ_LMAX=1024 CFDL_RBQ_LBIG=1536 run 128_tickets1536 --steps 12
CFDL_RBQ_LMAX=1024 CFDL_RBQ_LBIG=3072 run 128_tickets3072 --steps 12
