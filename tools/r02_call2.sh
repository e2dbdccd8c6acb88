# Round 2, GPU call 2: the persistent pc solve (kernels_rbq.inc) — parity, timing against the pass-by-pass kernels, ncu of the assembly kernels
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_host_routines.py -m gpu -x -q -rxXs --durations=8 > gpurun_out/r02c2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c2_pytest_gpu.log
tail -5 gpurun_out/r02c2_pytest_gpu.log
run() { name=$1; shift; timeout 400 python bench.py --steps 30 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02c2_bench_$name.json 2> gpurun_out/r02c2_bench_$name.err; tail -c 300 gpurun_out/r02c2_bench_$name.json; echo; }
run rbq
run rbq_off --opt rbq=0 --no-e2e
run rbq_c2 --opt rbq_ctas=2 --no-e2e
run rbq_c1 --opt rbq_ctas=1 --no-e2e
run rbq_i32 --opt rb_idx16=0 --no-e2e
PIN="--opt autotune=0 --opt uvw_variant=13 --opt grad_variant=1 --opt coef_p_variant=3 --opt mip_fast=1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c2_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $PIN > gpurun_out/r02c2_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"coef_uvw|coef_p_statics|grad_lsq|mip_cells|correct_faces|residual_kernel" -c 14 \
  -o gpurun_out/r02c2_assembly python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $PIN > gpurun_out/r02c2_ncu_assembly.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rbq_kernel|rb3_" -c 5 \
  -o gpurun_out/r02c2_solver python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $PIN > gpurun_out/r02c2_ncu_solver.log 2>&1
ls -la gpurun_out | tail -20
