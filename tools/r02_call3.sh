# Round 2, GPU call 3: occupancy of the persistent pc solve, slot-parallel calc_coef_uvw, ncu of rbq_kernel, compute-sanitizer
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coef_uvw or persistent" > gpurun_out/r02c3_pytest.log 2>&1; tail -3 gpurun_out/r02c3_pytest.log
run() { name=$1; shift; timeout 400 python bench.py --steps 30 --warmup 3 "$@" > gpurun_out/r02c3_bench_$name.json 2> gpurun_out/r02c3_bench_$name.err; tail -c 200 gpurun_out/r02c3_bench_$name.json; echo; }
run default
for m in 4 5 6; do run minb$m --opt rbq_minb=$m --no-e2e --no-cpu-baseline; done
for v in 1 2 3 4; do run slots$v --opt uvw_slots=$v --no-e2e --no-cpu-baseline --steps 12; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c3_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02c3_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rbq_kernel" -c 1 \
  -o gpurun_out/r02c3_rbq python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02c3_ncu_rbq.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"coef_uvw_slots" -c 1 \
  -o gpurun_out/r02c3_slots python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --opt uvw_slots=1 > gpurun_out/r02c3_ncu_slots.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hex8_sub4 and (coef_uvw or persistent or run_parity or calc_mip or coef_p or solve_uvwp)" > gpurun_out/r02c3_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02c3_sanitizer_memcheck.log
tail -4 gpurun_out/r02c3_sanitizer_memcheck.log
ls -la gpurun_out | tail -15
