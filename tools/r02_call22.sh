# Round 2, GPU call 22 (1 GPU): L2 prefetch of the next trips' constants in the persistent pc solve (option rbq_prefetch = trips ahead), A/B on one box
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 3 "$@" > gpurun_out/r02c22_bench_$name.json 2> gpurun_out/r02c22_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c22_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    c=d["config"]
    print("$name", "value %.1fM ms/step %.3f" % (d["value"]/1e6, d["ms_per_step"]), "pass_us %.2f frac %.3f" % (r.get("avg_launch_ms",0)*1e3, r.get("frac",0)), c["solver_iterations_last_step(u,v,w,pc)"], "sgs %.3f" % d["phase_ms_per_step"].get("sgs", 0))
except Exception as e: print("$name ERR", e)
PY
}
run pf0_a
run pf1_a --opt rbq_prefetch=1
run pf2_a --opt rbq_prefetch=2
run pf3_a --opt rbq_prefetch=3
run pf0_b
run pf1_b --opt rbq_prefetch=1
run pf2_b --opt rbq_prefetch=2
run pf1_static --opt rbq_prefetch=1 --opt rbq_counter=0
run 160_pf0 --size 160 --steps 6
run 160_pf1 --size 160 --steps 6 --opt rbq_prefetch=1
run 160_pf2 --size 160 --steps 6 --opt rbq_prefetch=2
run smp_nvml --steps 30
run smp_smi --steps 30 --clock-sampler smi
run smp_off --steps 30 --clock-sampler off
python - <<'PY'
import json
for n in ("smp_nvml","smp_smi"):
    d=json.loads(open("gpurun_out/r02c22_bench_%s.json"%n).read().strip().splitlines()[-1]); print(n, d["clocks"])
PY
