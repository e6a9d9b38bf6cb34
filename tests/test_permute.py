"""swb_permute_members: re-enumerating the members of an ensemble changes nothing but their index."""
import numpy as np
import pytest

import parity_common as pc
from swmm_b200 import solver


def _golden_solver(case, M, scale, shift, lib_path):
    net, g = pc.load_golden(case)
    nP = net.n_pollut
    s = solver.Solver(net, M, lib_path=lib_path)
    s.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
    s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"], ts_q=g["inf_ts_q"],
                  sfactor=g["inf_sfactor"], baseline=g["inf_baseline"], concen=g["inf_concen"] if nP else None,
                  member_scale=scale, member_shift=shift,
                  start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
    return s, net, g


def _permute_check(lib_path, M, steps):
    """swb_permute_members re-enumerates the members and nothing else: a permuted ensemble, stepped on, equals
    the unpermuted one member by member (state, clocks, counters, statistics, routing totals) bit for bit."""
    scale = np.linspace(0.3, 2.5, M)
    shift = np.linspace(-0.02, 0.02, M)
    a, net, g = _golden_solver("c2_grid12_slot", M, scale, shift, lib_path)
    b, _, _ = _golden_solver("c2_grid12_slot", M, scale, shift, lib_path)
    a.enable_statistics(0.0)
    b.enable_statistics(0.0)
    t_end = float(g["t_end"])
    try:
        a.run_steps(steps, t_end)
        b.run_steps(steps, t_end)
        rng = np.random.default_rng(3)
        perm = rng.permutation(M).astype(np.int32)
        b.permute_members(perm)
        a.run_steps(steps, t_end)
        b.run_steps(steps, t_end)
        for f in ("SWB_NODE_NEW_DEPTH", "SWB_LINK_NEW_FLOW", "SWB_NODE_NEW_QUAL", "SWB_LINK_TOTAL_LOAD", "SWB_COND_Q1"):
            if f.endswith("QUAL") and net.n_pollut == 0:
                continue
            assert np.array_equal(a.get_field(f)[perm], b.get_field(f)), f
        sa, sb = a.stats(), b.stats()
        for i in range(M):
            x, y = sa[perm[i]], sb[i]
            assert (x.sim_time, x.iterations, x.steps, x.non_converged, x.next_dt) == (y.sim_time, y.iterations, y.steps, y.non_converged, y.next_dt)
        na, la_, ya = a.statistics()
        nb, lb, yb = b.statistics()
        assert np.array_equal(na[perm], nb) and np.array_equal(la_[perm], lb) and np.array_equal(ya[perm], yb)
        ta, tb = a.routing_totals(), b.routing_totals()
        for da, db in zip(ta, tb):          # (sums of unordered atomic adds: equal to rounding)
            for k in da:
                np.testing.assert_allclose(np.asarray(da[k])[perm], np.asarray(db[k]), rtol=1e-11, atol=1e-9, err_msg=k)
        with pytest.raises(solver.SwbError):
            b.permute_members(np.zeros(M, dtype=np.int32))
    finally:
        a.close()
        b.close()


def test_permute_members_emulated(emul_lib):
    _permute_check(emul_lib, 32, 40)


@pytest.mark.gpu
def test_permute_members_cuda(cuda_lib):
    _permute_check(None, 256, 150)


@pytest.mark.gpu
def test_two_solvers_on_two_devices_in_one_process(cuda_lib):
    """Every entry point makes its solver's device current: two ensembles on two GPUs of one
    process, calls interleaved, both equal a run on their own."""
    lib = solver.load_library()
    if lib.swb_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    M, steps = 64, 60
    scale = np.linspace(0.5, 2.0, M)
    net, g = pc.load_golden("c2_grid12_slot")
    nP = net.n_pollut

    def make(dev):
        s = solver.Solver(net, M, device=dev)
        s.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
        s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"], ts_q=g["inf_ts_q"],
                      sfactor=g["inf_sfactor"], baseline=g["inf_baseline"], concen=g["inf_concen"] if nP else None,
                      member_scale=scale, start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
        return s
    a, b = make(0), make(1)
    ref = make(0)
    t_end = float(g["t_end"])
    try:
        ref.run_steps(steps, t_end)
        for _ in range(steps // 10):          # interleaved: the current device flips between every call
            a.run_steps(10, t_end)
            b.run_steps(10, t_end)
            a.get_field("SWB_NODE_NEW_DEPTH", 3, 1)
        for f in ("SWB_NODE_NEW_DEPTH", "SWB_LINK_NEW_FLOW"):
            want = ref.get_field(f)
            assert np.array_equal(a.get_field(f), want) and np.array_equal(b.get_field(f), want), f
        assert a.last_kernel_ms() > 0.0 and b.last_kernel_ms() > 0.0
    finally:
        a.close(); b.close(); ref.close()


@pytest.mark.gpu
def test_solver_larger_than_device_memory_fails_cleanly(cuda_lib):
    """allocations are checked -- an ensemble that does not fit the device returns SWB_ERR_CUDA
    (nothing half-built, no sticky error) and the device stays usable."""
    from swmm_b200 import network, scenarios
    case = network.build_grid(scenarios.GridSpec(nx=230, ny=230, hours=1.0))        # ~25 MB of state per member
    with pytest.raises(solver.SwbError, match="(?i)memory|alloc"):
        solver.Solver(case.net, 8192)                                               # ~205 GB > 180 GB of HBM
    s = solver.Solver(case.net, 32)
    s.load_state(case.state0)
    s.set_inflows(**case.inflows)
    s.run_steps(5, case.t_end)
    assert all(x.steps == 5 for x in s.stats())
    s.close()
