# Round 2, GPU call 27 (1 GPU): default bench line of the final tree (clock samples of the timed region only)
set -u
mkdir -p gpurun_out
timeout 80 python bench.py > gpurun_out/r02c27_bench_default.json 2> gpurun_out/r02c27_bench_default.err; echo rc=$?; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c27_bench_default.json").read().strip().splitlines()[-1]); print(d["value"]/1e6, d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("dram_frac"), d["clocks"], d["e2e"]["value"]/1e6, d["cpu_baseline"]["value"])
PY
