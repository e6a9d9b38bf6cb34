"""Per-object statistics on the device: after a whole run every plane of
swb_get_node_stats / swb_get_link_stats must equal the reference engine's own NodeStats / StorageStats /
OutfallStats / LinkStats (stats.c:449-754) for the same model -- maximum depths / flows and their times,
time flooded / surcharged, lateral inflow volumes, flow-class times, flow turns, Courant-critical and
non-converged counts, outfall loads -- and MaxOutfallFlow.

CPU: host build of the device engine, bit-exact.  GPU: 1e-6 on the values (counts exact)."""
import numpy as np
import pytest

import parity_common as pc
from swmm_b200 import abi, solver

NS, LS, SS = abi.NODE_STAT, abi.LINK_STAT, abi.SYSTEM_STAT
COUNT_PLANES_N = [NS["SWB_NS_NONCONV_COUNT"], NS["SWB_NS_TIME_COURANT"]]
COUNT_PLANES_L = [LS["SWB_LS_TURNS"], LS["SWB_LS_TURN_SIGN"], LS["SWB_LS_TIME_COURANT"]]
TIME_PLANES_N = [NS["SWB_NS_MAX_DEPTH_TIME"], NS["SWB_NS_MAX_INFLOW_TIME"], NS["SWB_NS_MAX_OVERFLOW_TIME"]]
TIME_PLANES_L = [LS["SWB_LS_MAX_FLOW_TIME"]]


def run_both(case, lib_path, n_members=1):
    """Reference to the end of the simulation + the ensemble driver on the same model."""
    e, _ = pc.open_reference(pc.case_inp(case))
    try:
        net = e.network()
        s = pc.make_solver_from_engine(e, lib_path, n_members)
        s.enable_statistics(0.0)
        t_end = e.total_duration_s()
        ref_steps = 1
        while e.step() != 0:
            ref_steps += 1
        ref = e.statistics(net.n_nodes, net.n_links, net.n_pollut)
    finally:
        e.end()
        e.close()
    while s.stats()[0].sim_time < t_end:
        s.run_steps(100000, t_end)
    got = s.statistics()
    steps = s.stats()[0].steps
    s.close()
    return net, ref, got, ref_steps, steps


def compare(net, ref, got, tol, m=0):
    nd, ld, sd = got
    rn, rl, rmax = ref
    worst = 0.0
    storage = net.arrays["node_type"] == 2
    for p in range(nd.shape[1]):
        a, b = nd[m, p], rn[p]
        if p in COUNT_PLANES_N:
            assert np.array_equal(a, b), ("node count plane", p, np.nonzero(a != b)[0][:5])
        elif p in TIME_PLANES_N or p == NS["SWB_NS_X_MAX_TIME"]:
            # dates pass through DateTime (days): 1 ms; (outfalls keep a period COUNT in this plane)
            assert np.all(np.abs(a - b) <= 2e-3), ("node time plane", p, float(np.max(np.abs(a - b))))
        else:
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-6)
            worst = max(worst, float(err.max()))
    pump = net.arrays["link_type"] == 1
    for p in range(ld.shape[1]):
        a, b = ld[m, p], rl[p]
        if p in COUNT_PLANES_L or p in (LS["SWB_LS_PUMP_STARTUPS"], LS["SWB_LS_PUMP_PERIODS"]):
            ok = np.array_equal(a, b) if p in COUNT_PLANES_L else np.array_equal(a[pump], b[pump])
            assert ok, ("link count plane", p, np.nonzero(a != b)[0][:5])
        elif p in TIME_PLANES_L:
            assert np.all(np.abs(a - b) <= 2e-3), ("link time plane", p)
        else:
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-6)
            worst = max(worst, float(err.max()))
    assert abs(sd[m, SS["SWB_SS_MAX_OUTFALL_FLOW"]] - rmax) <= tol * max(abs(rmax), 1e-6), (sd[m], rmax)
    assert worst <= tol, worst
    return worst


@pytest.mark.parametrize("case", ["c1_tree", "c2_grid12_slot", "c2_grid12_extran"])
def test_emulated_statistics_equal_reference(case, emul_lib, have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref is not built")
    net, ref, got, ref_steps, steps = run_both(case, emul_lib)
    assert steps == ref_steps
    assert got[2][0, SS["SWB_SS_REPORT_STEPS"]] == steps
    worst = compare(net, ref, got, 1e-12)
    print(case, "steps", steps, "worst relative difference", worst)


def test_statistics_are_per_member(emul_lib, have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref is not built")
    net, ref, got, _, _ = run_both("c2_grid12_slot", emul_lib, n_members=32)
    for m in (0, 17, 31):
        compare(net, ref, got, 1e-12, m)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c1_tree", "c2_grid30_slot"])
def test_cuda_statistics_match_reference(case, cuda_lib, have_reference):
    assert have_reference, "oracle/_ref must travel to the GPU box"
    net, ref, got, ref_steps, steps = run_both(case, None, n_members=32)
    assert steps == ref_steps
    for m in (0, 31):
        worst = compare(net, ref, got, 1e-6, m)
    print(case, "steps", steps, "worst relative difference", worst)
