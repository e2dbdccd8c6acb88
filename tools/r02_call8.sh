# Round 2, GPU call 8 (2 GPUs): flag-in-data exchange (no fences, no flag words) in the partitioned persistent pc solve
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q  > gpurun_out/r02c8_pytest_multi.log 2>&1; tail -3 gpurun_out/r02c8_pytest_multi.log
run() { name=$1; shift; timeout 600 $TR bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02c8_bench_$name.json 2> gpurun_out/r02c8_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c8_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1fM ms/step %.3f pass_us %.2f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"]*1e3), d["config"].get("pc_solve","")[:40], d["e2e"] and d["e2e"]["value"]/1e6, (d.get("parity_check") or {}).get("result"), d["phase_ms_per_step"])
except Exception as e: print("$name ERR", e)
PY
}
run stack
CFDL_RBQ_DEBUG=2 run nopush --no-e2e --no-parity-check
CFDL_RBQ_DEBUG=6 run nopush_noland --no-e2e --no-parity-check
run g256 --global-size 256 --no-e2e
run strong128 --global-size 128 --no-e2e --no-parity-check
