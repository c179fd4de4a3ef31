"""exp_tab.h (the table-driven exp of the per-individual kernels) against libm, compiled for the host."""
import math
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_table_exp_is_within_two_ulp_of_libm_and_handles_the_ends(tmp_path):
    exe = str(tmp_path / "exp_tab_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "csrc", "exp_tab_check.cpp")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    worst = float(out[0].split()[0])
    assert worst <= 2.0, out[0]
    vals = [float(v) for v in out[1:10]]
    assert vals[0] == 1.0 and vals[1] == 1.0
    assert vals[2] == 0.0 and vals[3] == 0.0 and vals[4] == 0.0          # below -708.39: flushed to zero
    assert math.isinf(vals[5]) and math.isinf(vals[6]) and math.isinf(vals[7])
    assert math.isnan(vals[8])
