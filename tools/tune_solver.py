"""Micro-benchmark of the pc solve on a FIXED system (B200): assemble the pressure-correction
system of the 4th SIMPLE iteration of the 128^3 cavity once, then time cfdl_solve_eq(pc, nit=100)
from phi=0 for several launch configurations.  Usage: python tools/tune_solver.py [n]"""
import sys

sys.path.insert(0, "cfd-lite_b200/python")
import numpy as np
import cfdl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
raw = cfdl.meshgen(0, n)
geom = cfdl.mesh_build(raw)
s = cfdl.Solver(geom, cfdl.default_bcs(raw))
s.set_option("solver", 1)
for i in range(3):
    s.update_boundaries()
    s.solve_uvwp()
s.update_boundaries()
s.calc_coef_uvw()
for eq in (0, 1, 2):
    s.solve_eq(eq)
for a, b in (("u", "gu"), ("v", "gv"), ("w", "gw")):
    s.calc_grad(a, b)
s.calc_mip(True)
s.calc_coef_p()
zeros = np.zeros(s.H)
N = s.ne


def run(label):
    s.upload("pc", zeros)
    s.solve_eq(3, 100)  # warm
    s.upload("pc", zeros)
    s.set_option("profile", 1)
    s.set_option("reset_counters", 1)
    s.timer_record(0)
    out = s.solve_eq(3, 100)
    s.timer_record(1)
    ms_tot = s.timer_elapsed_ms(0, 1)
    ms, cnt = s.get_info("prof_ms_sgs"), s.get_info("prof_n_sgs")
    s.set_option("profile", 0)
    s.upload("pc", zeros)
    s.timer_record(0)
    out = s.solve_eq(3, 100)
    s.timer_record(1)
    ms_np = s.timer_elapsed_ms(0, 1)
    print("%-28s it=%3d passes=%4d avg pass %.2f us  solve %.3f ms (profiled %.3f)  -> %.0f B/cell/it at %.0f GB/s"
          % (label, out[0], cnt, 1e3 * ms / max(cnt, 1), ms_np, ms_tot, 0, 120.0 * N * out[0] / (ms_np * 1e-3) / 1e9), flush=True)


for fused in (1, 0):
    s.set_option("fused", fused)
    for ctas in (4, 8, 12, 16, 24, 32, 100000):
        s.set_option("ctas_per_sm", ctas)
        run("fused=%d ctas_per_sm=%d" % (fused, ctas))
