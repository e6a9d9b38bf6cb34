"""BASELINE config 5 (1000 x 500 looped grid, 998 501 conduits) against the REFERENCE: tests/golden/
c5_grid_1000x500.npz holds the reference's time steps and Picard counts of every step of the first 10 simulated
minutes and, at seven steps, every 97th node depth / node volume / link flow / link depth plus the sum and the sum
of squares of the complete arrays (tests/golden/make_c5_golden.py, oracle/_ref).  The device replays the run as
ONE model on one GPU and striped over two (when the box has two): 1e-6 on the samples, the checksums and the
clock; identical Picard counts."""
import os

import numpy as np
import pytest

import parity_common as pc
from swmm_b200 import network, scenarios, solver

TOL = 1e-6
FIELDS = ["SWB_NODE_NEW_DEPTH", "SWB_NODE_NEW_VOLUME", "SWB_LINK_NEW_FLOW", "SWB_LINK_NEW_DEPTH"]
FLOOR = {"SWB_NODE_NEW_DEPTH": 1e-4, "SWB_NODE_NEW_VOLUME": 1e-3, "SWB_LINK_NEW_FLOW": 1e-4, "SWB_LINK_NEW_DEPTH": 1e-4}


def _fixture():
    path = os.path.join(pc.GOLDEN, "c5_grid_1000x500.npz")
    assert os.path.exists(path), "run tests/golden/make_c5_golden.py"
    return np.load(path)


def test_fixture_is_the_stated_config():
    g = _fixture()
    assert int(g["nx"]) == 1000 and int(g["ny"]) == 500
    assert abs(float(g["series_time"][-1]) - 600.0) < 1e-9
    assert g["sample_SWB_LINK_NEW_FLOW"].shape[1] == len(range(0, 998501, int(g["stride"])))
    assert g["sample_SWB_NODE_NEW_DEPTH"].shape[1] == len(range(0, 500001, int(g["stride"])))


def _compare(get_full, g, k, worst):
    stride = int(g["stride"])
    for f in FIELDS:
        a = get_full(f)
        ref = g["sample_" + f][k]
        worst[f] = max(worst.get(f, 0.0), pc.rel_err(a[::stride], ref, FLOOR[f]))
        s1, s2 = g["sums_" + f][k]
        worst[f + "_sum"] = max(worst.get(f + "_sum", 0.0), abs(float(np.sum(a)) - s1) / max(abs(s1), 1.0))
        worst[f + "_sumsq"] = max(worst.get(f + "_sumsq", 0.0), abs(float(np.sum(a * a)) - s2) / max(abs(s2), 1.0))


def _replay(lib_path, n_snaps=None):
    g = _fixture()
    spec = scenarios.GridSpec(nx=1000, ny=500, hours=float(g["sim_min"]) / 60.0, pollutants=False, surcharge="SLOT")
    case = network.build_grid(spec, lib_path=lib_path)
    s = solver.Solver(case.net, 1, lib_path=lib_path)
    s.load_state(case.state0)
    s.set_inflows(**case.inflows)
    times, iters = g["series_time"], g["series_iters"]
    worst, done = {}, 0
    try:
        for k, step in enumerate(g["snap_steps"][:n_snaps]):
            s.run_steps(int(step) - done, case.t_end)
            done = int(step)
            st = s.stats()[0]
            assert st.steps == done
            assert abs(st.sim_time - times[done - 1]) < 1e-9, (done, st.sim_time, times[done - 1])
            assert st.iterations == int(np.sum(iters[:done])), (done, st.iterations)
            _compare(lambda f: s.get_field(f)[0], g, k, worst)
    finally:
        s.close()
    print(worst)
    return done, len(times), worst


def test_emulated_config5_first_steps_equal_reference_fixture(emul_lib):
    """The host build of the device engine on the full 1M-link model, first 20 routing steps: samples and
    whole-array checksums equal the reference's to the last bit."""
    done, _, worst = _replay(emul_lib, n_snaps=5)
    assert done == 20 and max(worst.values()) == 0.0, worst


@pytest.mark.gpu
def test_cuda_config5_single_gpu_vs_reference_fixture(cuda_lib):
    done, total, worst = _replay(None)
    assert done == total
    assert max(worst.values()) <= TOL, worst


@pytest.mark.gpu
def test_cuda_config5_two_gpus_vs_reference_fixture(cuda_lib):
    """The same replay striped over two GPUs, one process per GPU (tests/partition_worker.py pattern)."""
    if solver.load_library().swb_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "c5_golden_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "C5 GOLDEN OK" in r.stdout
