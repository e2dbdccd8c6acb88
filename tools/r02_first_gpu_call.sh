#!/bin/bash
# First gpurun call of the next round: everything written after round 1's GPU minutes were spent gets
# its GPU verdict in ONE call (1 GPU, ~12 min).  Results land in gpurun_out/ (copy what is to be judged
# into profiles/).   gpurun --timeout 1500 -- 'bash tools/r02_first_gpu_call.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r02_clocks.csv &
SMI=$!
# 1. parity gate (XPASS of test_gpu_host_routines.py = the confirmation those tests wait for)
timeout 900 python -m pytest tests -m gpu -x -q -rxX > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
# 2. bench: autotuned default, round-1 kernels (autotune off), separate e2e calls, pcg
run() { name=$1; shift; timeout 400 python bench.py --steps 30 --warmup 3 "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; tail -c 600 gpurun_out/r02_bench_$name.json; echo; }
run default
run r01_kernels --opt autotune=0 --opt uvw_fused=0 --e2e-separate
run uvw_sep --opt uvw_fused=0 --no-e2e --no-cpu-baseline
run uvw_fused --opt uvw_fused=1 --no-e2e --no-cpu-baseline
run pcg --solver pcg --no-e2e --no-cpu-baseline
for pk in "0 0" "1 0" "0 35" "1 35" "0 60" "1 60"; do set -- $pk; run rb_p$1_k$2 --opt rb_persistent=$1 --opt rb_keep_mb=$2 --opt rb_idx16=0 --steps 12 --no-e2e --no-cpu-baseline; run rb_p$1_k$2_i16 --opt rb_persistent=$1 --opt rb_keep_mb=$2 --opt rb_idx16=1 --steps 12 --no-e2e --no-cpu-baseline; done
for t in 2 4 8 16; do for r in 1 2 4; do run wave_t${t}_r$r --opt rb_wave=$t --opt rb_wave_rows=$r --steps 12 --no-e2e --no-cpu-baseline; done; done
run wave_t8_r1_i16 --opt rb_wave=8 --opt rb_idx16=1 --steps 12 --no-e2e --no-cpu-baseline
run wave_t8_b8 --opt rb_wave=8 --opt rb_wave_block=8 --steps 12 --no-e2e --no-cpu-baseline
for v in 2 3 5 6 7 8 4 9 10 11 12 13 14; do run uvw_v$v --opt uvw_variant=$v --steps 10 --no-e2e --no-cpu-baseline; done
for v in 0 1; do run grad_v$v --opt grad_variant=$v --steps 10 --no-e2e --no-cpu-baseline; run coefp_v$v --opt coef_p_variant=$v --steps 10 --no-e2e --no-cpu-baseline; done
# 3. launch list of one default step + full capture of the assembly kernels and the side-by-side passes
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 700 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"coef_uvw|coef_p_statics|grad_|mip_cells|rb3_|correct_faces" -s 60 -c 14 \
  -o gpurun_out/r02_assembly python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_assembly.log 2>&1
kill $SMI
ls -la gpurun_out
