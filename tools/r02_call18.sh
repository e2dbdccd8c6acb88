# Round 2, GPU call 18 (1 GPU): 128^3 A/B — one chunk per CTA against chunks handed out from a counter; the automatic choice at 144^3 .. 256^3
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 3 "$@" > gpurun_out/r02c18_bench_$name.json 2> gpurun_out/r02c18_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c18_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    c=d["config"]
    print("$name", "value %.1fM ms/step %.3f" % (d["value"]/1e6, d["ms_per_step"]), "pass_us %.2f frac %.3f" % (r.get("avg_launch_ms",0)*1e3, r.get("frac",0)), c.get("pc_solve_chunks"), c["solver_iterations_last_step(u,v,w,pc)"], "sgs %.3f" % d["phase_ms_per_step"].get("sgs", 0))
except Exception as e: print("$name ERR", e)
PY
}
run 128_static_a
run 128_counter2048_a --opt rbq_counter=1
run 128_static_b
run 128_counter2048_b --opt rbq_counter=1
run 128_counter1792 --opt rbq_counter=1 --opt rbq_lbig=1792
run 128_counter1536 --opt rbq_counter=1 --opt rbq_lbig=1536
run 128_counter2304 --opt rbq_counter=1 --opt rbq_lbig=2304
run 128_counter1280 --opt rbq_counter=1 --opt rbq_lbig=1280
run 128_static_c
run 128_counter2048_c --opt rbq_counter=1
run 144_auto --size 144 --steps 6
run 160_auto --size 160 --steps 6
run 160_lbig3072 --size 160 --steps 6 --opt rbq_lbig=3072
run 168_auto --size 168 --steps 6
run 168_f70 --size 168 --steps 6 --opt rbq_l2_fraction=0.7
run 176_auto --size 176 --steps 6
run 256_auto --size 256 --steps 6
