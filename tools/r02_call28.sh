# Round 2, GPU call 28 (1 GPU, the last seconds of the budget): ncu --set full of one steady launch of the final persistent pc solve
set -u
mkdir -p gpurun_out
timeout 55 ncu --set full --clock-control none -k regex:"rbq_kernel" --launch-skip 4 -c 1 -o gpurun_out/r02c28_rbq_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --clock-sampler off > gpurun_out/r02c28_ncu.log 2>&1; echo rc=$?
ls -la gpurun_out/r02c28_*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
