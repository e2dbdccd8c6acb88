# Round 2, GPU call 1: probe for a Fortran compiler, parity gate, bench with the library's own choices, ncu.
set -u
mkdir -p gpurun_out
{
  echo "== Fortran compiler probe on the gpurun box ($(date -u +%FT%TZ)) =="
  for c in gfortran gfortran-13 gfortran-12 flang flang-new nvfortran pgfortran lfortran ifx ifort f95 f77 g77; do
    printf "%-14s " $c; command -v $c || echo "absent"; done
  echo "f951 via gcc: $(gcc -print-prog-name=f951)"; ls -la "$(gcc -print-prog-name=f951)" 2>&1
  ls /usr/lib/gcc/x86_64-linux-gnu/*/ 2>/dev/null | tr '\n' ' '; echo
  find / -xdev \( -name 'f951' -o -name 'gfortran*' -o -name 'nvfortran*' -o -name 'flang*' \) -not -path '/proc/*' 2>/dev/null | head -20
  echo "== host =="; lscpu | head -25; free -g
} > gpurun_out/r02_fortran_probe.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r02_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q -rxXs > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
run() { name=$1; shift; timeout 400 python bench.py --steps 30 --warmup 3 "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; tail -c 400 gpurun_out/r02_bench_$name.json; echo; }
run default
run r01_kernels --opt autotune=0 --opt uvw_fused=0 --no-e2e --no-cpu-baseline
run pcg --solver pcg --no-e2e --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 700 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"coef_uvw|coef_p_statics|grad_|mip_cells|rb3_|rb_red|rb_black|residual|correct_faces" -s 120 -c 24 \
  -o gpurun_out/r02_assembly python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_assembly.log 2>&1
kill $SMI
ls -la gpurun_out
