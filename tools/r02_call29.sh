# Round 2, GPU call 29 (1 GPU, the last seconds): the persistent pc solve built for four CTAs per SM (64 registers, 48 bytes of spills) against the shipped three
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 20 python bench.py --no-cpu-baseline --no-e2e --steps 12 --warmup 3 --clock-sampler off "$@" > gpurun_out/r02c29_$name.json 2>/dev/null; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c29_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]/1e6,1), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"]*1e3,2), d["config"]["pc_solve_chunks"])
except Exception as e: print("$name ERR", e)
PY
}
run occ3
cp tools/exp/libcfdl_rbq_occ4.so cfd-lite_b200/lib/libcfdl.so
run occ4_lbig1536 --opt rbq_lbig=1536
run occ4_static --opt rbq_counter=0
