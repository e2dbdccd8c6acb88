# Round 2, GPU call 11 (1 GPU): full -m gpu suite, final bench lines (MCSGS, PCG-SSOR, PCG-Jacobi; 128^3 and 256^3), ncu launch list
# and full captures of the kernels the bench times
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c11_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02c11_pytest_gpu.log
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r02c11_bench_$name.json 2> gpurun_out/r02c11_bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c11_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"] or {}
    print("$name", "value %.1fM ms/step %.3f" % (d["value"]/1e6, d["ms_per_step"]), "pass_us %.2f frac %.3f" % (r.get("avg_launch_ms",0)*1e3, r.get("frac",0)), d["e2e"] and round(d["e2e"]["value"]/1e6,1), {k: round(v, 3) for k, v in d["phase_ms_per_step"].items()}, d["config"]["last_step_history(it,res_i,res_f,res_max)"][3])
except Exception as e: print("$name ERR", e)
PY
}
run default
run pcg_ssor --solver pcg --no-cpu-baseline --no-e2e --steps 12
run pcg_jacobi --solver pcg --opt pcg_precond=0 --no-cpu-baseline --no-e2e --steps 12
run n256 --size 256 --structured --no-cpu-baseline --steps 12
run n256_pcg --size 256 --structured --solver pcg --no-cpu-baseline --no-e2e --steps 12
run tet --mesh tet --size 60 --no-cpu-baseline --no-e2e --steps 12
NB="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c11_launches.csv $NB > gpurun_out/r02c11_ncu_launches.log 2>&1
cap() { name=$1; regex=$2; shift 2; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$regex" -c 1 -o gpurun_out/r02c11_$name $NB "$@" > gpurun_out/r02c11_ncu_$name.log 2>&1; ls -la gpurun_out/r02c11_$name.ncu-rep 2>/dev/null | awk '{print $5, $9}'; }
cap rbq "rbq_kernel"
cap coef_uvw "coef_uvw_statics"
cap coef_p "coef_p_statics"
cap mip "mip_cells"
cap grad3 "grad_lsq_kernel|grad_kernel"
cap residual "residual_kernel"
cap rb3 "rb3_red"
cap perpass "rb_red_kernel" --opt rbq=0
