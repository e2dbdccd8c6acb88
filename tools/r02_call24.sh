# Round 2, GPU call 24 (8 GPUs): 8 stacked 128^3 blocks, 30 timed steps, clock sampler in-process (NVML) / off / nvidia-smi child
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524"
run() { name=$1; shift; timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02c24_n${N}_$name.json 2> gpurun_out/r02c24_n${N}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c24_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.1fM ms/step %.3f pass_us %.2f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"]*1e3), d["e2e"] and round(d["e2e"]["value"]/1e6,1), (d.get("parity_check") or {}).get("result"), "sum_of_phases", round(d["phase_ms_per_step"]["sum_of_phases"],3), d["config"]["ms_per_step_by_rank"], d["clocks"])
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02c24_n${N}_$name.err").read()[-800:])
PY
}
run stack_off --clock-sampler off --no-e2e --no-parity-check
run stack_nvml
run stack_smi --clock-sampler smi --no-e2e --no-parity-check
