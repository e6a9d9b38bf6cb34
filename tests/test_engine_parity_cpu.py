"""CPU-only parity: the host emulation of the device engine (same headers, same barriers) against
(a) the committed golden trajectories of the reference and (b) the live reference (oracle/_ref).

Bar (BASELINE.json north_star): node depths, link flows, concentrations within 1e-6 relative on the
time series.  In practice the restatement is bit-exact here because host libm == the reference's.
"""
import numpy as np
import pytest

import parity_common as pc

TOL = 1e-6


@pytest.mark.parametrize("case,steps", [("c1_tree", 400), ("c1_tree_slot", 400),
                                        ("c2_grid12_slot", 350), ("c2_grid12_extran", 350)])
def test_emulated_engine_replays_golden(case, steps, emul_lib):
    r = pc.run_golden_case(case, emul_lib, max_steps=steps)
    assert r["snapshots"] >= 10
    assert r["time_err_s"] == 0.0, r            # every variable step identical (ms-floored)
    assert r["iters_match"], r                 # same Picard trip count every step
    assert r["max_rel"] <= TOL, r
    assert r["max_rel"] == 0.0, r              # host build: bit-exact


def test_emulated_engine_full_tree_run_vs_live_reference(emul_lib, have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref not built")
    r = pc.lockstep_vs_reference(pc.case_inp("c1_tree"), emul_lib, every=50)
    assert r["steps"] == 4321
    assert r["time_err_s"] == 0.0 and r["max_rel"] == 0.0, r


def test_emulated_engine_grid_with_quality_vs_live_reference(emul_lib, have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref not built")
    for sur in ("slot", "extran"):
        r = pc.lockstep_vs_reference(pc.case_inp(f"c2_grid16_{sur}"), emul_lib, max_steps=700, every=25)
        assert r["time_err_s"] == 0.0 and r["max_rel"] == 0.0, (sur, r)
        assert r["non_converged"] == r["ref_non_converged"], (sur, r)
        # time-step critical element (dynwave.c:813-827): same arg-min as the reference every step
        assert r["crit_checked"] > 500 and r["crit_mismatch"] == 0, (sur, r)


def test_lockstep_members_are_independent(emul_lib):
    """32 identical members in one launch must each reproduce the single-member trajectory."""
    r = pc.run_golden_case("c2_grid12_slot", emul_lib, max_steps=60, n_members=32)
    assert r["time_err_s"] == 0.0 and r["iters_match"] and r["max_rel"] == 0.0, r


def test_seam_style_calls_match_ensemble_driver(emul_lib):
    """swb_get_routing_step + swb_old_state_swap + swb_dynwave_execute + swb_qualrout_execute
    (the reference's own call sequence) equals one swb_run_steps step."""
    from swmm_b200 import solver
    net, g = pc.load_golden("c2_grid12_slot")
    nP = net.n_pollut

    def fresh():
        s = solver.Solver(net, 1, lib_path=emul_lib)
        s.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
        s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"],
                      ts_q=g["inf_ts_q"], sfactor=g["inf_sfactor"], baseline=g["inf_baseline"],
                      concen=g["inf_concen"] if nP else None,
                      start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
        return s
    a, b = fresh(), fresh()
    t_end = float(g["t_end"])
    for step in range(40):
        a.run_steps(1, t_end)
        # host-driven sequence on b: the host does the bookkeeping the reference's routing.c does
        dt = b.get_routing_step(net.options["route_step"])[0]
        lat = a.get_field("SWB_NODE_NEW_LATFLOW")[0]
        b.old_state_swap(dt, with_quality=True)      # uses b's previous latflow for node_initFlows
        b.set_field("SWB_NODE_NEW_LATFLOW", lat)
        # quality preload = concentration x lateral inflow (routing.c:476-489)
        if nP:
            pre = np.zeros((net.n_nodes, nP))
            for k, node in enumerate(g["inf_node"]):
                if lat[node] >= 0:
                    pre[node] = g["inf_concen"].reshape(-1, nP)[k] * lat[node]
            b.set_field("SWB_NODE_NEW_QUAL", pre)
        it = b.dynwave_execute(dt)
        b.qualrout_execute(dt)
        assert it[0] == g["series_iters"][step]
        for f in ("SWB_NODE_NEW_DEPTH", "SWB_LINK_NEW_FLOW", "SWB_NODE_NEW_QUAL", "SWB_LINK_NEW_QUAL"):
            assert np.array_equal(a.get_field(f), b.get_field(f)), (step, f)


def test_step_host_equals_device_driver(emul_lib):
    """swb_step_host (host lateral inflows / quality loads in, depths / flows / next dt out, device
    transposes) reproduces swb_run_steps step for step, for several members at once."""
    from swmm_b200 import solver
    net, g = pc.load_golden("c2_grid12_slot")
    nP = net.n_pollut
    M = 32

    def fresh():
        s = solver.Solver(net, M, lib_path=emul_lib)
        s.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
        s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"],
                      ts_q=g["inf_ts_q"], sfactor=g["inf_sfactor"], baseline=g["inf_baseline"],
                      concen=g["inf_concen"] if nP else None,
                      member_scale=np.linspace(0.5, 1.5, M),
                      start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
        return s
    a, b = fresh(), fresh()
    lat = b.host_array((M, net.n_nodes))
    load = b.host_array((M, net.n_nodes, nP))
    depth = b.host_array((M, net.n_nodes))
    flow = b.host_array((M, net.n_links))
    next_dt = b.host_array((M,))
    iters = b.host_array((M,), dtype=np.int32)
    conc = np.zeros((net.n_nodes, nP))
    conc[g["inf_node"]] = g["inf_concen"].reshape(-1, nP)
    for step in range(60):
        a.run_steps(1, 1e9)
        lat[:] = a.get_field("SWB_NODE_NEW_LATFLOW")
        load[:] = np.maximum(lat, 0.0)[:, :, None] * conc[None]
        b.step_host(lat, qual_load=load, node_depth=depth, link_flow=flow, next_dt=next_dt, iters=iters)
        assert np.array_equal(depth, a.get_field("SWB_NODE_NEW_DEPTH")), step
        assert np.array_equal(flow, a.get_field("SWB_LINK_NEW_FLOW")), step
        assert np.array_equal(b.get_field("SWB_NODE_NEW_QUAL"), a.get_field("SWB_NODE_NEW_QUAL")), step
        sa = a.stats()
        assert all(next_dt[k] == sa[k].next_dt for k in range(M))
    assert len(set(np.round(depth[:, 5], 12))) > 1      # members really differ


def _ensemble_vs_single(lib_path, case, M, steps, scales, picks):
    """Members of a lockstep ensemble (different inflow scales -> different Picard trip counts,
    so the alive-member compaction is exercised) equal the same scenarios run one at a time."""
    from swmm_b200 import solver
    net, g = pc.load_golden(case)
    nP = net.n_pollut

    def make(m, scale):
        s = solver.Solver(net, m, lib_path=lib_path)
        s.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
        s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"],
                      ts_q=g["inf_ts_q"], sfactor=g["inf_sfactor"], baseline=g["inf_baseline"],
                      concen=g["inf_concen"] if nP else None, member_scale=scale,
                      start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
        return s
    ens = make(M, scales)
    ens.run_steps(steps, 1e9)
    st = ens.stats()
    iters = [x.iterations for x in st]
    assert len(set(iters)) > 1, "members should need different iteration counts"
    for k in picks:
        one = make(1, np.array([scales[k]]))
        one.run_steps(steps, 1e9)
        s1 = one.stats()[0]
        assert s1.iterations == st[k].iterations and s1.sim_time == st[k].sim_time, k
        for f in ("SWB_NODE_NEW_DEPTH", "SWB_LINK_NEW_FLOW", "SWB_NODE_NEW_QUAL", "SWB_LINK_NEW_QUAL",
                  "SWB_LINK_NEW_VOLUME"):
            assert np.array_equal(ens.get_field(f, k, 1), one.get_field(f)), (k, f)
        one.close()
    ens.close()
    return iters


def test_ragged_ensemble_members_equal_single_runs(emul_lib):
    scales = np.linspace(0.2, 3.0, 32)
    iters = _ensemble_vs_single(emul_lib, "c2_grid12_extran", 32, 700, scales, [0, 13, 31])
    assert max(iters) > min(iters)


def test_step_host_batch_equals_sequential(emul_lib):
    pc.batch_step_equals_sequential(emul_lib)
