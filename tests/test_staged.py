"""The staged kernel chain (csrc/swb_staged.cuh, wide ensembles) against the persistent cooperative
kernel on the same inputs: both run the phase functions of swb_engine.h, so every state field, time step
and Picard count must be IDENTICAL bit for bit; only the mass-balance sums (unordered atomic adds in
either form) are compared to rounding.  The persistent kernel itself is pinned against the reference by
tests/test_engine_parity_gpu.py, and the staged form is replayed against the golden trajectory here too."""
import numpy as np
import pytest

import parity_common as pc

pytestmark = pytest.mark.gpu

FIELDS = ["SWB_NODE_NEW_DEPTH", "SWB_NODE_NEW_VOLUME", "SWB_NODE_INFLOW", "SWB_NODE_OUTFLOW", "SWB_NODE_OVERFLOW",
          "SWB_NODE_OLD_DEPTH", "SWB_NODE_OLD_NET_INFLOW", "SWB_NODE_NEW_LATFLOW", "SWB_NODE_DYDT",
          "SWB_LINK_NEW_FLOW", "SWB_LINK_OLD_FLOW", "SWB_LINK_NEW_DEPTH", "SWB_LINK_NEW_VOLUME", "SWB_LINK_DQDH",
          "SWB_LINK_FROUDE", "SWB_LINK_SURF_AREA1", "SWB_LINK_SURF_AREA2", "SWB_COND_A1", "SWB_COND_A2",
          "SWB_COND_Q1", "SWB_LINK_FLOW_CLASS", "SWB_COND_CAPACITY_LIMITED", "SWB_LINK_BYPASSED",
          "SWB_NODE_CONVERGED"]


def _golden_solver(case, M, scale=None, shift=None):
    from swmm_b200 import solver
    net, g = pc.load_golden(case)
    nP = net.n_pollut
    s = solver.Solver(net, M)
    s.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
    s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"], ts_q=g["inf_ts_q"],
                  sfactor=g["inf_sfactor"], baseline=g["inf_baseline"], concen=g["inf_concen"] if nP else None,
                  member_scale=scale, member_shift=shift,
                  start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
    return s, net, g


def _compare(a, b, nP, tag):
    for f in FIELDS + (["SWB_NODE_NEW_QUAL", "SWB_LINK_NEW_QUAL", "SWB_NODE_OLD_QUAL"] if nP else []):
        x, y = a.get_field(f), b.get_field(f)
        assert np.array_equal(x, y), (tag, f, float(np.max(np.abs(x - y))))
    sa, sb = a.stats(), b.stats()
    for x, y in zip(sa, sb):
        assert (x.sim_time, x.next_dt, x.iterations, x.steps, x.non_converged, x.crit_node, x.crit_link) == \
               (y.sim_time, y.next_dt, y.iterations, y.steps, y.non_converged, y.crit_node, y.crit_link), tag
    fa, qa = a.routing_totals()
    fb, qb = b.routing_totals()
    for d1, d2 in ((fa, fb), (qa, qb)):
        for k in d1:
            np.testing.assert_allclose(d1[k], d2[k], rtol=1e-11, atol=1e-9, err_msg=f"{tag} {k}")


@pytest.mark.parametrize("case,M,steps", [("c2_grid12_slot", 64, 600), ("c2_grid12_extran", 96, 900),
                                          ("c1_tree_slot", 32, 300)])
def test_staged_equals_persistent(case, M, steps, cuda_lib):
    """Ragged ensembles (different Picard trip counts per member, members finishing at different steps)."""
    scale = np.linspace(0.3, 2.5, M)
    shift = np.linspace(-0.02, 0.02, M)
    a, net, g = _golden_solver(case, M, scale, shift)
    b, _, _ = _golden_solver(case, M, scale, shift)
    a.enable_statistics(0.0)
    b.enable_statistics(0.0)
    t_end = float(g["t_end"])
    done = 0
    try:
        for chunk in (1, 7, 50, steps):
            a.set_staged_min_members(0)          # persistent
            a.run_steps(chunk, t_end)
            la = a.launch_count()
            a.set_staged_min_members(16)         # staged
            b.run_steps(chunk, t_end)
            done += chunk
            _compare(a, b, net.n_pollut, f"{case} after {done} steps")
        assert b.launch_count() > 10 * la, "the staged form launches one kernel per phase"
        na, la_, sa = a.statistics()
        nb, lb, sb = b.statistics()
        assert np.array_equal(na, nb) and np.array_equal(la_, lb) and np.array_equal(sa, sb)
        iters = [x.iterations for x in b.stats()]
        print(case, "iterations per member: min", min(iters), "max", max(iters))
    finally:
        a.set_staged_min_members(256)
        a.close()
        b.close()


def test_staged_run_to_end_in_one_call(cuda_lib):
    """n_steps far beyond the end of the simulation: the chain stops when every member has reached t_end."""
    a, net, g = _golden_solver("c2_grid12_slot", 32, np.linspace(0.5, 1.5, 32), None)
    b, _, _ = _golden_solver("c2_grid12_slot", 32, np.linspace(0.5, 1.5, 32), None)
    t_end = 1500.0
    try:
        a.set_staged_min_members(0)
        a.run_steps(10_000_000, t_end)
        a.set_staged_min_members(16)
        b.run_steps(10_000_000, t_end)
        _compare(a, b, net.n_pollut, "run to end")
        assert all(abs(x.sim_time - t_end) < 1e-9 for x in b.stats())
    finally:
        a.set_staged_min_members(256)
        a.close()
        b.close()


def test_staged_replays_golden(cuda_lib):
    from swmm_b200 import solver
    s = solver.Solver(pc.load_golden("c2_grid12_slot")[0], 1)
    try:
        s.set_staged_min_members(16)
        r = pc.run_golden_case("c2_grid12_slot", None, max_steps=400, n_members=32)
        print(r)
        assert r["time_err_s"] < 1e-9 and r["iters_match"] and r["max_rel"] <= 1e-6, r
    finally:
        s.set_staged_min_members(256)
        s.close()


def test_staged_mixed_elements_with_regulators(cuda_lib, have_reference):
    """Pumps, orifices, weirs, outlets, storage: the ordered regulator pass as a kernel of its own."""
    assert have_reference
    e, _ = pc.open_reference(pc.case_inp("c3_mixed"))
    a = pc.make_solver_from_engine(e, None, n_members=32, member_scale=np.linspace(0.5, 1.5, 32))
    b = pc.make_solver_from_engine(e, None, n_members=32, member_scale=np.linspace(0.5, 1.5, 32))
    try:
        for chunk in (3, 40, 400):
            a.set_staged_min_members(0)
            a.run_steps(chunk, 1e9)
            a.set_staged_min_members(16)
            b.run_steps(chunk, 1e9)
            _compare(a, b, a.net.n_pollut, f"c3_mixed +{chunk}")
    finally:
        a.set_staged_min_members(256)
        a.close()
        b.close()
        e.end()
        e.close()


def test_staged_step_host_batch(cuda_lib):
    """The host-buffer step (transposes as kernels of the chain) on pipelined member blocks."""
    from swmm_b200 import solver
    s = solver.Solver(pc.load_golden("c2_grid12_slot")[0], 1)
    try:
        s.set_staged_min_members(16)
        pc.batch_step_equals_sequential(cuda_lib, M=128, blocks=4)
    finally:
        s.set_staged_min_members(256)
        s.close()

