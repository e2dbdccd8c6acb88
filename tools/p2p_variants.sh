#!/bin/bash
# 2-GPU timing experiments for the peer-to-peer solver passes (results of the DEBUG variants are wrong by design)
mkdir -p gpurun_out
run() { # name, env, extra args
  env $2 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e $3 > gpurun_out/h_$1.json 2> gpurun_out/h_$1.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/h_$1.json").read().strip().splitlines()[-1])
    print("$1", round(l["value"]/1e6,1), round(l["ms_per_step"],3), l["gpu_launches"])
except Exception as e:
    print("$1 failed", e)
PY
}
run default "A=1" ""
run localres "CFDL_P2P_DEBUG=4" ""
run nowait "CFDL_P2P_DEBUG=5" ""
run ctas7 "A=1" "--opt ctas_per_sm=7"
run ctas6 "A=1" "--opt ctas_per_sm=6"
