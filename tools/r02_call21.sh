# Round 2, GPU call 21 (8 GPUs): weak series end point on the final tree (8 stacked 128^3 blocks; clock sampler on rank 0 only; with / without pinned host cores), BASELINE config 5 (512^3), multi-GPU tests
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
nproc; 
run() { name=$1; shift; timeout 900 $TR bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02c21_n${N}_$name.json 2> gpurun_out/r02c21_n${N}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c21_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1fM ms/step %.3f pass_us %.2f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"]*1e3), d["config"].get("pc_solve","")[:30], d["config"].get("pc_solve_chunks"), d["e2e"] and round(d["e2e"]["value"]/1e6,1), (d.get("parity_check") or {}).get("result"), d["config"]["solver_iterations_mean_over_timed_steps(u,v,w,pc)"], "sum_of_phases", round(d["phase_ms_per_step"]["sum_of_phases"],3), "setup", d["config"]["setup_seconds"], d["config"]["host_cores"], d["config"]["host_cpus"], d["clocks"])
except Exception as e: print("$name ERR", e); import subprocess; print(open("gpurun_out/r02c21_n${N}_$name.err").read()[-800:])
PY
}
run stack
run stack_bound --bind-cores --no-e2e
run g512 --global-size 512 --no-e2e
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02c21_pytest_multi.log 2>&1; tail -3 gpurun_out/r02c21_pytest_multi.log
