# Round 2, GPU call 5 (2 GPUs): the persistent pc solve on partitioned meshes (chunk-to-chunk synchronisation over NVLink)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02c5_pytest_multi.log 2>&1; tail -3 gpurun_out/r02c5_pytest_multi.log
run() { name=$1; shift; timeout 600 $TR bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02c5_bench_$name.json 2> gpurun_out/r02c5_bench_$name.err; tail -c 400 gpurun_out/r02c5_bench_$name.json; echo; tail -3 gpurun_out/r02c5_bench_$name.err; }
run stack
run stack_rcb --partition rcb --no-e2e --no-parity-check
run g256 --global-size 256 --no-e2e
run strong128 --global-size 128 --no-e2e --no-parity-check
timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02c5_bench_n1.json 2> gpurun_out/r02c5_bench_n1.err; tail -c 300 gpurun_out/r02c5_bench_n1.json
