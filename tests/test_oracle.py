"""Pins the C restatement oracle (oracle/swmm_oracle.c) against the reference engine: the golden
trajectories written by the unmodified engine, and the live engine where oracle/_ref exists.
Bit-exact: same compiler family, same libm, same operation order."""
import numpy as np
import pytest

import parity_common as pc
import oracle as orc


def replay(case, max_steps=None):
    net, g = pc.load_golden(case)
    o = orc.Oracle(net)
    o.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
    o.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"], ts_q=g["inf_ts_q"],
                  sfactor=g["inf_sfactor"], baseline=g["inf_baseline"],
                  concen=g["inf_concen"] if net.n_pollut else None,
                  start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
    t_end = float(g["t_end"])
    snap_steps = list(g["snap_steps"])
    n = len(g["series_time"]) if max_steps is None else min(max_steps, len(g["series_time"]))
    checked = 0
    for step in range(1, n + 1):
        it = o.step(t_end)
        assert it == g["series_iters"][step - 1], (case, step, it)
        assert o.time == g["series_time"][step - 1], (case, step)
        if step in snap_steps:
            idx = snap_steps.index(step)
            for f in pc.SNAP_FIELDS:
                if f"snap_{f}" in g and f != "SWB_NODE_OVERFLOW":
                    assert np.array_equal(o.get_field(f), g[f"snap_{f}"][idx]), (case, step, f)
            assert np.array_equal(o.get_field("SWB_NODE_OVERFLOW"), g["snap_SWB_NODE_OVERFLOW"][idx])
            checked += 1
    o.close()
    return checked


@pytest.mark.parametrize("case", ["c1_tree", "c1_tree_slot", "c2_grid12_slot", "c2_grid12_extran"])
def test_oracle_replays_reference_golden_bit_for_bit(case):
    assert replay(case) >= 20


def test_oracle_rejects_what_it_does_not_cover(have_reference):
    if not have_reference:
        pytest.skip("needs the engine to flatten the mixed-element model")
    e, _ = pc.open_reference(pc.case_inp("c3_mixed"))
    try:
        with pytest.raises(NotImplementedError):
            orc.Oracle(e.network())
    finally:
        e.end(); e.close()


def test_oracle_vs_live_reference_on_a_fresh_grid(have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref not built")
    e, _ = pc.open_reference(pc.case_inp("c2_grid9_slot"))
    try:
        o = orc.Oracle(e.network())
        o.load_state(pc.grab_state(e))
        inf = e.inflows()
        o.set_inflows(**inf)
        t_end = e.total_duration_s()
        for step in range(600):
            e.step()
            it = o.step(t_end)
            assert it == e.last_iterations()
            assert o.time == e.routing_time_ms() / 1000.0
            if step % 40 == 0:
                for f in ("SWB_NODE_NEW_DEPTH", "SWB_LINK_NEW_FLOW", "SWB_NODE_NEW_QUAL", "SWB_LINK_NEW_QUAL"):
                    assert np.array_equal(o.get_field(f), e.field(f)), (step, f)
        o.close()
    finally:
        e.end(); e.close()
