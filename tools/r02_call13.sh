# Round 2, GPU call 13 (1 GPU): cp.async form of calc_coef_uvw, steady-state ncu capture of the persistent pc solve, PCG iteration statistics
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_transport.py tests/test_gpu_vtu.py -m gpu -x -q > gpurun_out/r02c13_pytest.log 2>&1; tail -3 gpurun_out/r02c13_pytest.log
run() { name=$1; shift; timeout 900 python bench.py --no-cpu-baseline --no-e2e --steps 12 "$@" > gpurun_out/r02c13_bench_$name.json 2> gpurun_out/r02c13_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c13_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    c=d["config"]
    print("$name", "value %.1fM ms/step %.3f" % (d["value"]/1e6, d["ms_per_step"]), "pass_us %.2f" % (r.get("avg_launch_ms",0)*1e3), {k: round(v, 3) for k, v in d["phase_ms_per_step"].items()}, c["solver_iterations_mean_over_timed_steps(u,v,w,pc)"], c["pc_residual_reduction_mean(res_f/res_i)"])
except Exception as e: print("$name ERR", e)
PY
}
run async1
run async0 --opt uvw_async=0
run pcg_ssor --solver pcg --steps 30
run pcg_jacobi --solver pcg --opt pcg_precond=0 --steps 30
run mcsgs30 --steps 30
NB="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"coef_uvw_async" -c 1 -o gpurun_out/r02c13_coef_uvw_async $NB > gpurun_out/r02c13_ncu_uvw.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"rbq_kernel" --launch-skip 4 -c 1 -o gpurun_out/r02c13_rbq_steady $NB > gpurun_out/r02c13_ncu_rbq.log 2>&1
ls -la gpurun_out/r02c13_*.ncu-rep | awk '{print $5, $9}'
