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
