# Round 2, GPU call 12 (8 GPUs): weak series (8 stacked 128^3 blocks), BASELINE config 5 (512^3 on 8 GPUs), multi-GPU tests
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { name=$1; shift; timeout 1200 $TR bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02c12_n${N}_$name.json 2> gpurun_out/r02c12_n${N}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c12_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1fM ms/step %.3f pass_us %.2f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"]*1e3), d["config"].get("pc_solve","")[:40], d["e2e"] and round(d["e2e"]["value"]/1e6,1), (d.get("parity_check") or {}).get("result"), d["config"]["solver_iterations_mean_over_timed_steps(u,v,w,pc)"], {k: round(v,3) for k,v in d["phase_ms_per_step"].items()}, "setup", d["config"]["setup_seconds"])
except Exception as e: print("$name ERR", e); import subprocess; print(open("gpurun_out/r02c12_n${N}_$name.err").read()[-800:])
PY
}
run stack
run g512 --global-size 512 --no-e2e
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02c12_pytest_multi.log 2>&1; tail -3 gpurun_out/r02c12_pytest_multi.log
