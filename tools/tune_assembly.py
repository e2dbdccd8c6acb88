"""Micro-benchmark of the assembly kernels on the 128^3 cavity after a few SIMPLE iterations:
per-launch CUDA-event times of calc_coef_uvw (both variants), calc_coef_p, calc_mip, calc_grad.
Usage: python tools/tune_assembly.py [n]"""
import sys

sys.path.insert(0, "cfd-lite_b200/python")
import cfdl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
raw = cfdl.meshgen(0, n)
geom = cfdl.mesh_build(raw)
s = cfdl.Solver(geom, cfdl.default_bcs(raw))
s.set_option("solver", 1)
for i in range(3):
    s.update_boundaries()
    s.solve_uvwp()
N, F, H = s.ne, s.nf, s.H
B = H - N
Z = 6 * N
bytes_ = {"coef_uvw": 196 * N + 16 * Z + 56 * F + 48 * H, "coef_p": 36 * N + 16 * Z + 56 * F + 24 * H, "mip": 124 * N + 72 * F,
          "grad": 28 * N + 4 * Z + 32 * H}


def timeit(label, fn, key, reps=20):
    fn()
    s.set_option("profile", 1)
    s.set_option("reset_counters", 1)
    for _ in range(reps):
        fn()
    ms, cnt = s.get_info("prof_ms_" + key), s.get_info("prof_n_" + key)
    s.set_option("profile", 0)
    avg = ms / max(cnt, 1)
    print("%-34s %.4f ms  -> %.0f GB/s algorithmic (%d launches)" % (label, avg, bytes_[key] / (avg * 1e-3) / 1e9, cnt), flush=True)


for v in (0, 1, 2):
    s.set_option("uvw_variant", v)
    timeit("calc_coef_uvw variant %d" % v, lambda: s.calc_coef_uvw(0.01), "coef_uvw")
for st in (0, 1):
    s.set_option("statics", st)
    timeit("calc_coef_p statics=%d" % st, s.calc_coef_p, "coef_p")
    timeit("calc_mip statics=%d" % st, lambda: s.calc_mip(True, 0.01), "mip")
s.timer_record(0)
for i in range(6):
    s.update_boundaries()
    s.solve_uvwp()
s.timer_record(1)
print("step %.3f ms" % (s.timer_elapsed_ms(0, 1) / 6))
