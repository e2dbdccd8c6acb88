# Round 2, GPU call 26 (1 GPU, the last seconds of the budget): transport tests and a short default bench on the final tree
set -u
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_transport.py tests/test_gpu_vtu.py -m gpu -x -q > gpurun_out/r02c26_pytest.log 2>&1; tail -2 gpurun_out/r02c26_pytest.log
timeout 60 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02c26_bench.json 2> gpurun_out/r02c26_bench.err; echo rc=$?; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c26_bench.json").read().strip().splitlines()[-1]); print(d["value"]/1e6, d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("dram_frac"), d["clocks"])
PY
