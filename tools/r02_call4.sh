# Round 2, GPU call 4 (2 GPUs): state of the partitioned path before the persistent multi-rank pc solve
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02c4_pytest_multi.log 2>&1; tail -3 gpurun_out/r02c4_pytest_multi.log
run() { name=$1; shift; timeout 600 $TR bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02c4_bench_$name.json 2> gpurun_out/r02c4_bench_$name.err; tail -c 300 gpurun_out/r02c4_bench_$name.json; echo; }
run weak161
run strong128 --global-size 128 --no-e2e
run g256 --global-size 256 --structured --no-e2e
timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02c4_bench_n1.json 2> gpurun_out/r02c4_bench_n1.err; tail -c 300 gpurun_out/r02c4_bench_n1.json
