# Round 2, GPU call 10 (4 GPUs): weak series (stacked 128^3 blocks) and the 256^3 cube on 4 GPUs
set -u
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516"
run() { name=$1; shift; timeout 900 $TR bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02c10_n${N}_$name.json 2> gpurun_out/r02c10_n${N}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c10_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1fM ms/step %.3f pass_us %.2f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"]*1e3), d["config"].get("pc_solve","")[:40], d["e2e"] and d["e2e"]["value"]/1e6, (d.get("parity_check") or {}).get("result"), d["config"]["solver_iterations_last_step(u,v,w,pc)"], d["phase_ms_per_step"])
except Exception as e: print("$name ERR", e)
PY
}
run stack
run g256 --global-size 256 --no-e2e
run g256_rcb --global-size 256 --no-e2e --partition rcb --structured --no-parity-check
