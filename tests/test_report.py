"""Report-time result extraction: swb_get_results against the reference's own
node_getResults / link_getResults (node.c:497-528, link.c:674-724) called on the live engine at the
same routing step with the same weighting factor.  float32 records, compared bit for bit."""
import numpy as np
import pytest

import parity_common as pc
from swmm_b200 import solver


def lockstep_results(inp_text, lib_path, steps, check_every, n_members=1):
    e, _ = pc.open_reference(inp_text)
    try:
        s = pc.make_solver_from_engine(e, lib_path, n_members)
        net = s.net
        t_end = e.total_duration_s()
        checked = 0
        worst = 0.0
        for step in range(1, steps + 1):
            if e.step() == 0:
                break
            s.run_steps(1, t_end)
            if step % check_every:
                continue
            for f in (0.0, 0.37, 1.0):
                rn, rl = e.results(f, net.n_nodes, net.n_links, net.n_pollut)
                gn, gl = s.results(f)
                for k in range(n_members):
                    if not (np.array_equal(gn[k], rn) and np.array_equal(gl[k], rl)):
                        dn = np.abs(gn[k].astype(np.float64) - rn) / np.maximum(np.abs(rn), 1e-6)
                        dl = np.abs(gl[k].astype(np.float64) - rl) / np.maximum(np.abs(rl), 1e-6)
                        worst = max(worst, float(dn.max()), float(dl.max()))
                checked += 1
        s.close()
        return checked, worst
    finally:
        e.end()
        e.close()


def test_results_mixed_elements_from_engine_state(emul_lib, have_reference):
    """Pumps, orifices, weirs, outlet, storage, all conduit shapes: the engine runs alone (controls,
    DWF patterns and all); at check points its state is copied into the solver and the records are
    compared, every column."""
    if not have_reference:
        pytest.skip("oracle/_ref not built")
    for case in ("c3_mixed", "c3b_shapes"):
        e, _ = pc.open_reference(pc.case_inp(case))
        try:
            net = e.network()
            s = solver.Solver(net, 1, lib_path=emul_lib)
            checked = 0
            for step in range(1, 1501):
                if e.step() == 0:
                    break
                if step % 150:
                    continue
                s.load_state(pc.grab_state(e))
                for f in (0.0, 0.6, 1.0):
                    rn, rl = e.results(f, net.n_nodes, net.n_links, net.n_pollut)
                    gn, gl = s.results(f)
                    assert np.array_equal(gn[0], rn), (case, step, f, np.argwhere(gn[0] != rn)[:5])
                    assert np.array_equal(gl[0], rl), (case, step, f, np.argwhere(gl[0] != rl)[:5])
                    checked += 1
            assert checked >= 9
            s.close()
        finally:
            e.end()
            e.close()


@pytest.mark.parametrize("case", ["c2_grid12_slot", "c1_tree"])
def test_results_equal_reference_records(case, emul_lib, have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref not built")
    checked, worst = lockstep_results(pc.case_inp(case), emul_lib, steps=400, check_every=40)
    assert checked >= 20
    assert worst == 0.0, worst          # host build: every float32 identical


def test_results_layout_and_member_ranges(emul_lib):
    net, g = pc.load_golden("c2_grid12_slot")
    s = solver.Solver(net, 32, lib_path=emul_lib)
    s.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
    s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"], ts_q=g["inf_ts_q"],
                  sfactor=g["inf_sfactor"], baseline=g["inf_baseline"], concen=g["inf_concen"],
                  member_scale=np.linspace(0.5, 1.5, 32),
                  start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
    s.run_steps(120, 1e9)
    f = np.linspace(0.0, 1.0, 32)
    nd, ld = s.results(f)
    assert nd.shape == (32, net.n_nodes, 8) and ld.shape == (32, net.n_links, 7)
    sub_n, sub_l = s.results(f, member0=8, n_members=4)
    assert np.array_equal(sub_n, nd[8:12]) and np.array_equal(sub_l, ld[8:12])
    # depth record = interpolated depth in float32; head = depth + invert (node.c:510-513)
    d_old, d_new = s.get_field("SWB_NODE_OLD_DEPTH"), s.get_field("SWB_NODE_NEW_DEPTH")
    want = ((1.0 - f)[:, None] * d_old + f[:, None] * d_new).astype(np.float32)
    assert np.array_equal(nd[:, :, 0], want)
    assert np.array_equal(nd[:, :, 1], want + net.arrays["node_invert"].astype(np.float32)[None])
    assert float(np.abs(ld[:, :, 0]).max()) > 0.0 and np.all(ld[:, :, 4] >= 0.0) and np.all(ld[:, :, 4] <= 1.0)
    only_n, none_l = s.results(f, links=False)
    assert none_l is None and np.array_equal(only_n, nd)
    s.close()


@pytest.mark.gpu
def test_cuda_results_equal_host_emulation():
    """The CUDA report kernel against the host build of the same functions on the same state
    (copied field by field): A(y) is table arithmetic, so every float32 must be identical."""
    assert pc.cuda_available()
    emul = pc.build_emul()
    net, g = pc.load_golden("c2_grid12_slot")
    a = solver.Solver(net, 32)
    a.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
    a.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"], ts_q=g["inf_ts_q"],
                  sfactor=g["inf_sfactor"], baseline=g["inf_baseline"], concen=g["inf_concen"],
                  member_scale=np.linspace(0.5, 1.5, 32),
                  start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
    a.run_steps(300, 1e9)
    b = solver.Solver(net, 32, lib_path=emul)
    for fld in solver.Solver.STATE_FIELDS:
        if fld != "SWB_COND_Q2":
            b.set_field(fld, a.get_field(fld))
    f = np.linspace(0.0, 1.0, 32)
    an, al = a.results(f)
    bn, bl = b.results(f)
    assert float(np.abs(al[:, :, 0]).max()) > 0.0
    assert np.array_equal(an, bn) and np.array_equal(al, bl)
    a.close()
    b.close()


def test_results_argument_checks(emul_lib):
    net, g = pc.load_golden("c1_tree")
    s = solver.Solver(net, 1, lib_path=emul_lib)
    with pytest.raises(solver.SwbError):
        s.results(0.5, member0=1, n_members=1)          # member range out of bounds
    nd, ld = s.results(0.5)
    assert nd.shape == (1, net.n_nodes, 6) and ld.shape == (1, net.n_links, 5)   # no pollutants
    assert not np.any(nd[:, :, 0]) and not np.any(ld[:, :, 0])                      # fresh solver: all zero
    s.close()


def replay_report_golden(case, lib_path):
    """Committed fixtures of the reference's own records (tests/golden/make_report_golden.py):
    returns the worst relative difference and the fraction of float32 values that are identical."""
    net, g = pc.load_golden("report_" + case)
    s = solver.Solver(net, 1, lib_path=lib_path)
    worst, same, total = 0.0, 0, 0
    for k in range(len(g["steps"])):
        for key in g:
            if key.startswith(f"s{k}_"):
                s.broadcast_field(key[len(f"s{k}_"):], g[key])
        for i, f in enumerate(g["f"]):
            nd, ld = s.results(float(f))
            for got, ref in ((nd[0], g[f"node_{k}_{i}"]), (ld[0], g[f"link_{k}_{i}"])):
                same += int(np.sum(got == ref))
                total += ref.size
                d = np.abs(got.astype(np.float64) - ref) / np.maximum(np.abs(ref.astype(np.float64)), 1e-4)
                worst = max(worst, float(d.max()))
    s.close()
    return worst, same / total


@pytest.mark.parametrize("case", ["c2_grid12_slot", "c3_mixed", "c3b_shapes"])
def test_emulated_results_replay_golden_records(case, emul_lib):
    worst, same = replay_report_golden(case, emul_lib)
    assert worst == 0.0 and same == 1.0, (worst, same)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c2_grid12_slot", "c3_mixed", "c3b_shapes"])
def test_cuda_results_replay_golden_records(case, cuda_lib):
    """The reference's records on the device: 1e-6 relative (north_star); CUDA's libm differs from
    glibc by <= 2 ulp in the shapes that use acos / pow, which can move a float32 by one ulp."""
    worst, same = replay_report_golden(case, None)
    assert worst <= 1e-6, (worst, same)
    assert same >= 0.999, (worst, same)
