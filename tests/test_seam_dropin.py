"""Drop-in boundary: the unmodified reference CLI (oracle/_ref/runswmm) with the seam shim preloaded
must write the same binary .out and the same .rpt as the reference alone.

CPU: seam shim linked against the host emulation of the device engine -> byte-identical files.
GPU: the shipped libswmm5_b200_seam.so (CUDA) -> .out compared value by value at float32
resolution (the file stores float32, output.c:48-51) within 1e-6 relative.
"""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

import parity_common as pc

REF = os.path.join(pc.ROOT, "oracle", "_ref")
RUNSWMM = os.path.join(REF, "runswmm")
SEAM_EMUL = os.path.join(pc.EMUL_DIR, "libswmm5_b200_seam_emul.so")
SEAM_CUDA = os.path.join(pc.ROOT, "stormwater-management-model_b200", "seam", "libswmm5_b200_seam.so")
SEAM_SRC = os.path.join(pc.ROOT, "stormwater-management-model_b200", "seam")
REFSRC = "/root/reference"


def build_seam_emul(emul_lib):
    if not os.path.isdir(REFSRC):
        return os.path.exists(SEAM_EMUL)
    srcs = [os.path.join(SEAM_SRC, f) for f in ("seam.c", "flatten.c", "flatten.h")]
    if os.path.exists(SEAM_EMUL) and all(os.path.getmtime(s) <= os.path.getmtime(SEAM_EMUL) for s in srcs):
        return True
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-O2", "-fPIC", "-shared", "-w", f"-I{REFSRC}/src/solver",
                    f"-I{REFSRC}/src/solver/include", f"-I{pc.ROOT}/include", f"-I{SEAM_SRC}",
                    srcs[0], srcs[1], f"-L{pc.EMUL_DIR}", "-lswb_emul", "-Wl,-rpath,$ORIGIN",
                    "-o", SEAM_EMUL], check=True)
    return True


def run_cli(inp_text, preload=None):
    d = tempfile.mkdtemp(prefix="swb_cli_")
    inp = os.path.join(d, "m.inp")
    open(inp, "w").write(inp_text)
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = preload
    r = subprocess.run([RUNSWMM, inp, os.path.join(d, "m.rpt"), os.path.join(d, "m.out")],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(d, "m.rpt"), os.path.join(d, "m.out")


def rpt_body(path):
    skip = ("Analysis begun", "Analysis ended", "Total elapsed")
    return [ln for ln in open(path, errors="replace") if not any(s in ln for s in skip)]


def out_results(path):
    """(header ints, float32 result block) of a SWMM binary output file (output.c:121-537)."""
    raw = open(path, "rb").read()
    tail = struct.unpack("<6i", raw[-24:])
    out_start, n_periods, err = tail[2], tail[3], tail[4]
    assert err == 0 and tail[5] == 516114522
    body = raw[out_start:-24]
    per = len(body) // n_periods
    vals = np.frombuffer(body, dtype=np.uint8).reshape(n_periods, per)[:, 8:]
    return n_periods, np.ascontiguousarray(vals).view(np.float32)


@pytest.mark.parametrize("case", ["c1_tree", "c2_grid12_slot", "c2_grid12_extran", "c3_mixed", "c3b_shapes",
                                  "c3c_culverts_hw", "c3c_culverts_dw", "c3_large_6h"])
def test_cli_with_emulated_seam_is_byte_identical(case, emul_lib, have_reference):
    if not have_reference or not build_seam_emul(emul_lib):
        pytest.skip("oracle/_ref or the seam shim is not built")
    text = pc.case_inp(case)
    rpt0, out0 = run_cli(text)
    rpt1, out1 = run_cli(text, preload=SEAM_EMUL)
    assert open(out0, "rb").read() == open(out1, "rb").read()
    assert rpt_body(rpt0) == rpt_body(rpt1)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c1_tree", "c2_grid12_slot", "c3_mixed", "c3b_shapes", "c2_grid30_slot",
                                  "c3c_culverts_hw", "c3c_culverts_dw", "c3_large"])
def test_cli_with_cuda_seam_matches_reference(case, cuda_lib, have_reference):
    assert have_reference and os.path.exists(SEAM_CUDA), "oracle/_ref and the CUDA seam must travel"
    text = pc.case_inp(case)
    rpt0, out0 = run_cli(text)
    rpt1, out1 = run_cli(text, preload=SEAM_CUDA)
    n0, a = out_results(out0)
    n1, b = out_results(out1)
    assert n0 == n1 and a.shape == b.shape
    scale = np.maximum(np.abs(a), 1e-3)
    rel = float(np.max(np.abs(a - b) / scale))
    frac_equal = float(np.mean(a == b))
    print(case, "periods", n0, "max rel", rel, "identical float32 values", frac_equal)
    assert rel <= 1e-6, (case, rel)
    # continuity errors within 0.01 percentage points (north_star)
    def cont(path):
        vals = []
        for ln in open(path, errors="replace"):
            if "Continuity Error (%)" in ln:
                vals.append(float(ln.split()[-1]))
        return vals
    c0, c1 = cont(rpt0), cont(rpt1)
    assert len(c0) == len(c1) and all(abs(x - y) <= 0.01 for x, y in zip(c0, c1)), (c0, c1)
