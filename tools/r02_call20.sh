# Round 2, GPU call 20 (1 GPU): the final tree — full -m gpu suite, smoke(), default bench line, launch list, steady-state ncu capture of the pc solve (chunks from a counter)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c20_pytest.log 2>&1; tail -3 gpurun_out/r02c20_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c20_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r02c20_smoke.log
timeout 900 python bench.py > gpurun_out/r02c20_bench_default.json 2> gpurun_out/r02c20_bench_default.err; echo bench rc=$?; cut -c1-1500 gpurun_out/r02c20_bench_default.json
NB="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02c20_launches.csv $NB > gpurun_out/r02c20_ncu_bench.log 2>&1; wc -l gpurun_out/r02c20_launches.csv
timeout 600 ncu --set full --clock-control none -k regex:"rbq_kernel" --launch-skip 4 -c 1 -o gpurun_out/r02c20_rbq_steady $NB > gpurun_out/r02c20_ncu_rbq.log 2>&1
ls -la gpurun_out/r02c20_*.ncu-rep | awk '{print $5, $9}'
timeout 600 python bench.py --size 160 --steps 6 --no-cpu-baseline > gpurun_out/r02c20_bench_160.json 2> gpurun_out/r02c20_bench_160.err; cut -c1-300 gpurun_out/r02c20_bench_160.json
