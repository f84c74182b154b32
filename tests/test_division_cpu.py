"""The GAE kernels divide by the cached d_t table with q0 = x*r, q = fma(fma(-d,q0,x), r, q0), r = RN(1/d)
(di_hpc_b200/csrc/gae.cu div_by_table) instead of an IEEE division; DESIGN.md 4.2 claims this IS the correctly rounded
quotient.  Check the claim with the host's hardware FMA over every denominator the recurrence produces."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.timeout(300)
def test_table_division_is_correctly_rounded(tmp_path):
    exe = str(tmp_path / "divcheck")
    subprocess.run(["/usr/bin/gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe,
                    os.path.join(HERE, "csrc", "div_by_table_check.c"), "-lm"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    cases, bad = int(out[0]), int(out[1])
    assert cases >= 60_000_000 and bad == 0, (cases, bad)
