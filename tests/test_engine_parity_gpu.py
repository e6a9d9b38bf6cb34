"""GPU parity (run with -m gpu on the B200): the CUDA path through the C-ABI against the reference.

Truth sources: committed golden trajectories (always) and the live reference engine oracle/_ref
(prebuilt, travels to the GPU box).  Tolerance 1e-6 relative on depths / flows / concentrations
(north_star); time steps must agree to the millisecond floor.  CUDA's pow/exp/sin differ from
glibc's by <= 2 ulp, so results are not required to be bit-identical, only within tolerance.
"""
import numpy as np
import pytest

import parity_common as pc

TOL = 1e-6
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case,steps", [("c1_tree", None), ("c1_tree_slot", None),
                                        ("c2_grid12_slot", None), ("c2_grid12_extran", 400)])
def test_cuda_engine_replays_golden(case, steps, cuda_lib):
    r = pc.run_golden_case(case, None, max_steps=steps)
    print(case, r)
    assert r["snapshots"] >= 10
    assert r["time_err_s"] < 1e-9, r
    assert r["iters_match"], r
    assert r["max_rel"] <= TOL, r


def test_cuda_full_grid_with_quality_vs_live_reference(cuda_lib, have_reference):
    assert have_reference, "oracle/_ref must travel to the GPU box"
    r = pc.lockstep_vs_reference(pc.case_inp("c2_grid30_slot"), None, every=20)
    print(r)
    assert r["time_err_s"] < 1e-9 and r["max_rel"] <= TOL, r
    assert r["non_converged"] == r["ref_non_converged"], r
    assert r["crit_mismatch"] == 0, r


def test_cuda_lockstep_ensemble_members(cuda_lib):
    """64 identical members in one cooperative launch reproduce the single-member trajectory."""
    r = pc.run_golden_case("c2_grid12_slot", None, max_steps=300, n_members=64)
    print(r)
    assert r["time_err_s"] < 1e-9 and r["iters_match"] and r["max_rel"] <= TOL, r


def test_cuda_multi_step_launch_equals_single_steps(cuda_lib):
    """n steps in ONE persistent launch == n launches of one step (no host round trip needed)."""
    from swmm_b200 import solver
    net, g = pc.load_golden("c2_grid12_slot")
    nP = net.n_pollut

    def fresh():
        s = solver.Solver(net, 32)
        s.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
        s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"],
                      ts_q=g["inf_ts_q"], sfactor=g["inf_sfactor"], baseline=g["inf_baseline"],
                      concen=g["inf_concen"] if nP else None,
                      start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
        return s
    a, b = fresh(), fresh()
    t_end = float(g["t_end"])
    a.run_steps(250, t_end)
    for _ in range(250):
        b.run_steps(1, t_end)
    for f in ("SWB_NODE_NEW_DEPTH", "SWB_LINK_NEW_FLOW", "SWB_NODE_NEW_QUAL"):
        assert np.array_equal(a.get_field(f), b.get_field(f)), f
    assert a.stats()[0].iterations == b.stats()[0].iterations
    assert a.launch_count() == 1 and b.launch_count() == 250


def test_cuda_ragged_ensemble_members_equal_single_runs(cuda_lib):
    """Different members need different Picard trip counts; with the alive-member compaction each
    must still equal its own single-member run bit for bit (same device arithmetic)."""
    from test_engine_parity_cpu import _ensemble_vs_single
    scales = np.linspace(0.2, 3.0, 64)
    iters = _ensemble_vs_single(None, "c2_grid12_extran", 64, 900, scales, [0, 21, 40, 63])
    print("iterations per member: min", min(iters), "max", max(iters))


def test_cuda_step_host_batch_equals_sequential(cuda_lib):
    """Pipelined member blocks on their own streams give the same bits as one whole-ensemble step."""
    pc.batch_step_equals_sequential(cuda_lib, M=128, blocks=4)
