# Round 2, GPU call 19 (N GPUs, default 2): multi-GPU tests, stacked 128^3 blocks with chunks from a counter (default) and one chunk per CTA, strong 256^3 / 160-per-GPU cases
set -u
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02c19_n${N}_pytest_multi.log 2>&1; tail -3 gpurun_out/r02c19_n${N}_pytest_multi.log
run() { name=$1; shift; timeout 1200 $TR bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02c19_n${N}_$name.json 2> gpurun_out/r02c19_n${N}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c19_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1fM ms/step %.3f pass_us %.2f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"]*1e3), d["config"].get("pc_solve","")[:30], d["config"].get("pc_solve_chunks"), d["e2e"] and round(d["e2e"]["value"]/1e6,1), (d.get("parity_check") or {}).get("result"), d["config"]["solver_iterations_mean_over_timed_steps(u,v,w,pc)"], {k: round(v,3) for k,v in d["phase_ms_per_step"].items()}, "setup", d["config"]["setup_seconds"])
except Exception as e: print("$name ERR", e); import subprocess; print(open("gpurun_out/r02c19_n${N}_$name.err").read()[-800:])
PY
}
run stack
CFDL_RBQ_COUNTER=0 run stack_static --no-e2e
run stack_b --no-e2e
run stack160 --size 160 --no-e2e
CFDL_RBQ_COUNTER=0 run stack160_passes --size 160 --no-e2e
