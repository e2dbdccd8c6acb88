# Round 2, GPU call 14 (1 GPU): full -m gpu suite on the committed tree (incl. the GPU mesh builder), smoke(), default bench line, launch list of the default bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c14_pytest.log 2>&1; tail -3 gpurun_out/r02c14_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c14_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r02c14_smoke.log
timeout 900 python bench.py > gpurun_out/r02c14_bench_default.json 2> gpurun_out/r02c14_bench_default.err; echo bench rc=$?; cat gpurun_out/r02c14_bench_default.json | cut -c1-1800
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c14_bench_reference.json 2> gpurun_out/r02c14_bench_reference.err; echo ref rc=$?; cut -c1-600 gpurun_out/r02c14_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02c14_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02c14_ncu_bench.log 2>&1; wc -l gpurun_out/r02c14_launches.csv
