# Round 2, GPU call 6 (2 GPUs): where does the partitioned persistent pc solve lose its time?
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
run() { name=$1; shift; timeout 600 $TR bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check "$@" > gpurun_out/r02c6_bench_$name.json 2> gpurun_out/r02c6_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c6_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1fM ms/step %.3f pass_us %.2f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"]*1e3), d["config"].get("pc_solve","")[:40])
except Exception as e: print("$name ERR", e)
PY
}
run fencefix
CFDL_RBQ_DEBUG=2 run nopush
CFDL_RBQ_DEBUG=4 run noremote
CFDL_RBQ_DEBUG=6 run nopush_noremote
run perpass --opt rbq=0
