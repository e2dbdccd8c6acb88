# Round 2, GPU call 25 (1 GPU): the final tree — full -m gpu suite and the default bench line
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/r02c25_pytest.log 2>&1; tail -3 gpurun_out/r02c25_pytest.log
timeout 100 python bench.py > gpurun_out/r02c25_bench_default.json 2> gpurun_out/r02c25_bench_default.err; echo bench rc=$?; cut -c1-300 gpurun_out/r02c25_bench_default.json
