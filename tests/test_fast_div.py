"""quot<true> in device_math.cuh replaces a/b by three FP64 instructions on a precomputed
reciprocal (uvw_variant 3/4 of calc_coef_uvw).  It is only admissible because it returns the
correctly rounded quotient; tools/check_fast_div.cpp compares it with IEEE division on random and
adversarial operands (host fma == device fma.rn)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reciprocal_fma_quotient_equals_division(tmp_path):
    exe = str(tmp_path / "check_fast_div")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tools", "check_fast_div.cpp")])
    out = subprocess.run([exe, "20000000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "one-correction mismatches=0" in out.stdout
