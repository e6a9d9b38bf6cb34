"""Routing totals on the device: flow-routing and quality continuity errors of
an ensemble member against the reference's own massbal report for the same model.
Bar (BASELINE.json north_star): within 0.01 percentage points."""
import json
import os

import numpy as np
import pytest

import parity_common as pc
from swmm_b200 import solver

TOL_PP = 0.01
GOLDEN = os.path.join(pc.GOLDEN, "continuity.json")


def run_case(case, lib_path, n_members=1, chunk=100000):
    net, g = pc.load_golden(case)
    s = solver.Solver(net, n_members, lib_path=lib_path)
    s.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
    s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"], ts_q=g["inf_ts_q"],
                  sfactor=g["inf_sfactor"], baseline=g["inf_baseline"],
                  concen=g["inf_concen"] if net.n_pollut else None,
                  start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
    init = s.storage()
    t_end = float(g["t_end"])
    while s.stats()[0].sim_time < t_end:
        s.run_steps(chunk, t_end)
    flow, qual = s.continuity(init)
    tot = s.routing_totals()
    s.close()
    return flow, qual, tot


@pytest.mark.parametrize("case", ["c1_tree", "c2_grid12_slot", "c2_grid12_extran"])
def test_emulated_continuity_matches_reference_report(case, emul_lib):
    ref = json.load(open(GOLDEN))[case]
    # several launches (the pending half step is carried across them) and one launch give the same
    flow, qual, tot = run_case(case, emul_lib, chunk=173)
    assert abs(flow[0] - ref["flow_error_pct"]) <= TOL_PP, (flow, ref)
    if qual.size:
        worst = qual[0][np.argmax(np.abs(qual[0]))]
        assert abs(worst - ref["qual_error_pct"]) <= TOL_PP, (qual, ref)
    # the totals themselves, against the reference's report table (volumes in ft3)
    # (parallel partial sums instead of the reference's sequential loops: 1e-9 relative)
    for k, v in ref["flow_totals_ft3"].items():
        assert abs(tot[0][k][0] - v) <= 1e-9 * max(abs(v), 1.0), (k, tot[0][k][0], v)
    for p, q in enumerate(ref.get("qual_totals", [])):
        for k in ("ex_inflow", "flooding", "outflow", "reacted", "seepage", "final_storage"):
            assert abs(tot[1][k][0][p] - q[k]) <= 1e-9 * max(abs(q[k]), 1.0), (p, k, tot[1][k][0][p], q[k])


def test_continuity_is_per_member(emul_lib):
    flow, qual, tot = run_case("c2_grid12_slot", emul_lib, n_members=32)
    assert np.all(np.abs(flow - flow[0]) < 1e-9) and np.all(np.abs(qual - qual[0]) < 1e-9)
    assert tot[0]["ex_inflow"][0] > 0.0 and tot[0]["outflow"][0] > 0.0


@pytest.mark.gpu
def test_cuda_continuity_matches_reference_report():
    assert pc.cuda_available()
    golden = json.load(open(GOLDEN))
    for case in ("c1_tree", "c2_grid12_slot"):
        ref = golden[case]
        flow, qual, _ = run_case(case, None, n_members=32)
        assert np.all(np.abs(flow - ref["flow_error_pct"]) <= TOL_PP), (case, flow[:4], ref)
        if qual.size:
            worst = qual[0][np.argmax(np.abs(qual[0]))]
            assert abs(worst - ref["qual_error_pct"]) <= TOL_PP, (case, qual[0], ref)
