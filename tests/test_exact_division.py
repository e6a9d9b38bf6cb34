"""div_rcp (csrc/swb_common.h) is the IEEE quotient, not an approximation: the geometry lookups and
the per-link constants (full depth, conduit length) divide through a precomputed reciprocal in three
operations, and every qualifying divisor must reproduce `x / d` bit for bit."""
import os
import subprocess

import parity_common as pc


def test_reciprocal_division_is_exact(tmp_path):
    exe = str(tmp_path / "divcheck")
    subprocess.run([pc.HOST_CXX, "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", f"-I{pc.CSRC}",
                    os.path.join(pc.EMUL_DIR, "divcheck.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "4000"], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    words = r.stdout.split()
    assert int(words[3]) > 2500 and int(words[5]) == 0, r.stdout


def test_slot_exponent_polynomial_matches_reference_expression(tmp_path):
    """Device-only replacement of pow(yNorm, 2.4) inside the Sjoberg slot factor: exp(-p(yNorm)) stays
    within 12 ulp of the reference's exp(-pow(yNorm, 2.4)) over the whole reachable range (measured 7;
    the reference expression itself is 3.8 ulp from the exact value)."""
    exe = str(tmp_path / "sjoberg_check")
    subprocess.run([pc.HOST_CXX, "-O2", "-std=c++20", "-mfma", "-ffp-contract=off", f"-I{pc.CSRC}", f"-I{pc.ROOT}/include",
                    os.path.join(pc.EMUL_DIR, "sjoberg_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
