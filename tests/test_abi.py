"""The C-ABI: header, ctypes mirror, X-macro field lists and exported symbols agree."""
import ctypes as C
import os
import re

import pytest

import parity_common as pc
from swmm_b200 import abi, solver

HEADER = os.path.join(pc.ROOT, "include", "swmm_b200.h")
STATE_H = os.path.join(pc.CSRC, "swb_state.h")


def declared_functions():
    text = abi._strip_comments(open(HEADER).read())
    return sorted(set(re.findall(r"\b(swb_[a-z_0-9]+)\s*\(", text)))


def test_header_parses_and_field_ids_unique():
    ids = list(abi.FIELD.values())
    assert len(ids) == len(set(ids))
    assert abi.FIELD["SWB_LINK_NEW_FLOW"] == 32 and abi.FIELD["SWB_FIELD_COUNT"] == 96
    assert abi.FIELD["SWB_NODE_OLD_INFLOW"] == 22


def test_device_descriptor_macro_matches_header():
    txt = open(STATE_H).read()
    block = txt[txt.index("#define SWB_DESC_ARRAYS"):txt.index("// link_flags bits")]
    macro = re.findall(r"X\((int|double), (\w+), (\w+)\)", block)
    header = [(base, name) for name, base in abi.DESC_ARRAYS]
    assert [(b, n) for b, n, _ in macro] == header


def test_state_field_macro_covers_public_fields():
    txt = open(STATE_H).read()
    block = txt[txt.index("#define SWB_STATE_FIELDS"):txt.index("#define SWB_MAX_TRIALS_CAP")]
    ids = re.findall(r"X\([a-z ]+, \w+, (SWB_\w+), \w+\)", block)
    assert len(ids) == len(set(ids))
    public = [k for k in abi.FIELD if k != "SWB_FIELD_COUNT"]
    # SWB_COND_Q2 is served from the q1 array (q2 == q1 under dynamic wave, dwflow.c:285-286)
    assert sorted(ids + ["SWB_COND_Q2"]) == sorted(public)


@pytest.mark.parametrize("which", ["emul", "cuda"])
def test_library_exports_every_declared_symbol(which, emul_lib):
    path = emul_lib if which == "emul" else solver.CUDA_LIB
    if which == "cuda" and not os.path.exists(path):
        pytest.skip("CUDA library not built in this checkout")
    lib = C.CDLL(path)
    for fn in declared_functions():
        assert hasattr(lib, fn), f"{fn} declared in swmm_b200.h but not exported by {path}"
    assert lib.swb_version() == 100


def test_cuda_library_has_no_cpu_path():
    """On a box without a GPU every compute entry point must fail loudly (no fallback)."""
    if not os.path.exists(solver.CUDA_LIB):
        pytest.skip("CUDA library not built in this checkout")
    if pc.cuda_available():
        pytest.skip("a GPU is present")
    net, _ = pc.load_golden("c1_tree")
    with pytest.raises(solver.SwbError):
        solver.Solver(net, 1)


def test_unsupported_elements_are_rejected(emul_lib):
    net, _ = pc.load_golden("c1_tree")
    net.arrays["xs_culvert"][3] = 99                 # beyond the 57 FHWA codes (culvert.c:33)
    with pytest.raises(solver.SwbError):
        solver.Solver(net, 1, lib_path=emul_lib)
    net, _ = pc.load_golden("c1_tree")
    net.arrays["xs_culvert"][3] = 5                  # a valid culvert code is accepted
    solver.Solver(net, 1, lib_path=emul_lib).close()


def test_member_count_rules(emul_lib):
    net, _ = pc.load_golden("c1_tree")
    with pytest.raises(solver.SwbError):
        solver.Solver(net, 33, lib_path=emul_lib)
    s = solver.Solver(net, 32, lib_path=emul_lib)
    assert s.M == 32
    s.close()
