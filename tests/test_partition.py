"""One network cut over several ranks: every rank must reproduce, for the objects
it owns, exactly what the unpartitioned solver computes -- same variable steps, same Picard trip
counts, bit-identical depths / flows / concentrations -- with the border exchange running inside the
kernel.  CPU suite: the host emulation of the engine, ranks as threads of one process and as two
`torch.distributed` (gloo) processes; the GPU test runs the same comparison on the device."""
import ctypes as C
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import parity_common as pc
from swmm_b200 import partition, solver

FIELDS = ["SWB_NODE_NEW_DEPTH", "SWB_NODE_NEW_VOLUME", "SWB_NODE_OVERFLOW", "SWB_NODE_INFLOW",
          "SWB_NODE_NEW_QUAL", "SWB_LINK_NEW_FLOW", "SWB_LINK_NEW_DEPTH", "SWB_LINK_NEW_VOLUME",
          "SWB_LINK_NEW_QUAL", "SWB_LINK_FLOW_CLASS", "SWB_COND_CAPACITY_LIMITED"]


def free_port() -> str:
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return str(sk.getsockname()[1])


def golden_setup(case):
    net, g = pc.load_golden(case)
    state = {k[3:]: g[k] for k in g if k.startswith("s0_")}
    inflows = dict(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"], ts_q=g["inf_ts_q"],
                   sfactor=g["inf_sfactor"], baseline=g["inf_baseline"],
                   concen=g["inf_concen"] if net.n_pollut else None,
                   start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
    return net, state, inflows, float(g["t_end"])


def run_parts_in_threads(parts, state, inflows, nP, lib_path, chunks, t_end, timeout_s=60.0):
    """All ranks in this process, one Python thread per rank per launch (ctypes drops the GIL)."""
    ps = [partition.PartitionedSolver(p, lib_path=lib_path, timeout_s=timeout_s) for p in parts]
    handles = [s.export_handle() for s in ps]
    for s in ps:
        s.connect(handles)
        s.load_state(partition.split_state(s.part, state, nP))
        s.set_inflows(**partition.split_inflows(s.part, inflows, nP))
    for n in chunks:
        errs = []

        def go(s):
            try:
                s.run_steps(n, t_end)
            except Exception as e:       # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=go, args=(s,)) for s in ps]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs
        yield ps


def compare_with_single(ps, single, net, tag):
    st0 = single.stats()[0]
    for s in ps:
        st = s.stats()[0]
        assert st.sim_time == st0.sim_time and st.iterations == st0.iterations, (tag, s.part.rank)
        assert st.next_dt == st0.next_dt and st.non_converged == st0.non_converged, (tag, s.part.rank)
    for f in FIELDS:
        fid = single._fid(f)
        w = 1 if f not in ("SWB_NODE_NEW_QUAL", "SWB_LINK_NEW_QUAL") else net.n_pollut
        if w == 0:
            continue
        n_items = net.n_nodes if f.startswith("SWB_NODE") else net.n_links
        got = partition.assemble([s.owned_field(f) for s in ps], n_items, w)
        ref = single.get_field(fid)[0]
        assert np.array_equal(got, ref), (tag, f, float(np.max(np.abs(got - ref))))
    # ghosts hold their owner's depth
    ref = single.get_field("SWB_NODE_NEW_DEPTH")[0]
    for s in ps:
        assert np.array_equal(s.get_field("SWB_NODE_NEW_DEPTH")[0], ref[s.part.node_gid]), (tag, "ghosts")


@pytest.mark.parametrize("case,n_ranks", [("c2_grid12_slot", 2), ("c2_grid12_slot", 3), ("c2_grid12_extran", 2)])
def test_partitioned_emulation_equals_single_domain(case, n_ranks, emul_lib):
    net, state, inflows, t_end = golden_setup(case)
    owner = partition.stripes(12, 12, n_ranks, extra_nodes=1)
    parts = partition.split_network(net, owner, n_ranks)
    assert sum(p.n_owned for p in parts) == net.n_nodes
    assert sum(int(p.link_owned.sum()) for p in parts) == net.n_links
    assert all(p.recv_node.size > 0 for p in parts)
    single = solver.Solver(net, 1, lib_path=emul_lib)
    single.load_state(state)
    single.set_inflows(**inflows)
    chunks = [1, 1, 3, 20, 50, 75, 100]
    for k, ps in enumerate(run_parts_in_threads(parts, state, inflows, net.n_pollut, emul_lib, chunks, t_end)):
        single.run_steps(chunks[k], t_end)
        compare_with_single(ps, single, net, (case, n_ranks, k))
        assert ps[0].exchanges() == ps[-1].exchanges() > 0
    # mass-balance terms: every loss is reported by exactly one rank
    mb = single.massbal()
    tot = {k: sum(s.massbal()[k] for s in ps) for k in mb}
    for k in mb:
        assert np.allclose(tot[k], mb[k], rtol=1e-12, atol=1e-300), k
    single.close()
    for s in ps:
        s.close()


def test_cut_regulator_is_rejected(emul_lib):
    net, _ = pc.load_golden("c2_grid12_slot")
    net.arrays["link_type"][5] = 2          # an orifice ...
    owner = np.zeros(net.n_nodes, dtype=np.int32)
    owner[net.arrays["link_node2"][5]] = 1  # ... whose ends land on different ranks
    with pytest.raises(ValueError):
        partition.split_network(net, owner, 2)


def test_lost_peer_times_out_instead_of_hanging(emul_lib):
    net, state, inflows, t_end = golden_setup("c2_grid12_slot")
    parts = partition.split_network(net, partition.stripes(12, 12, 2, extra_nodes=1), 2)
    ps = [partition.PartitionedSolver(p, lib_path=emul_lib, timeout_s=0.5) for p in parts]
    handles = [s.export_handle() for s in ps]
    for s in ps:
        s.connect(handles)
        s.load_state(partition.split_state(s.part, state, net.n_pollut))
        s.set_inflows(**partition.split_inflows(s.part, inflows, net.n_pollut))
    with pytest.raises(solver.SwbError, match="halo exchange timed out"):
        ps[0].run_steps(1, t_end)          # rank 1 never launches
    for s in ps:
        s.close()


def test_two_gloo_processes(emul_lib, tmp_path):
    """world_size 2 over torch.distributed/gloo: handles by all_gather_object, results gathered to
    rank 0 and compared there with the single-domain run."""
    script = os.path.join(pc.ROOT, "tests", "partition_worker.py")
    env = dict(os.environ, SWB_LIB=emul_lib, PYTHONPATH=os.pathsep.join([pc.ROOT, os.path.join(pc.ROOT, "tests")]))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", free_port(), script, "--backend", "gloo",
                        "--case", "c2_grid12_slot", "--steps", "120"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "partition parity ok" in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
def test_partitioned_gpu_two_processes_share_one_or_two_devices():
    """The CUDA path: two processes (own device each when the box has two, else both on device 0 --
    the windows are CUDA IPC mappings either way) against the single-domain CUDA run, bit for bit."""
    assert pc.cuda_available()
    script = os.path.join(pc.ROOT, "tests", "partition_worker.py")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([pc.ROOT, os.path.join(pc.ROOT, "tests")]))
    env.pop("SWB_LIB", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", free_port(), script, "--backend", "gloo",
                        "--case", "c2_grid12_slot", "--steps", "300", "--cuda"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "partition parity ok" in r.stdout, r.stdout[-2000:]


def test_partition_argument_checks(emul_lib):
    """Error behaviour of the partition entry points (all SWB_ERR_ARG / SWB_ERR_UNSUPP, no crash)."""
    net, state, inflows, t_end = golden_setup("c2_grid12_slot")
    parts = partition.split_network(net, partition.stripes(12, 12, 2, extra_nodes=1), 2)
    # a partitioned solver holds one member
    s = solver.Solver(parts[0].net, 32, lib_path=emul_lib)
    d = partition.PartitionDesc()
    d.rank, d.n_ranks, d.n_owned_nodes = 0, 2, parts[0].n_owned
    s.lib.swb_partition_attach.argtypes = [C.c_void_p, C.POINTER(partition.PartitionDesc)]
    with pytest.raises(solver.SwbError, match="one member"):
        s._chk(s.lib.swb_partition_attach(s._h, C.byref(d)))
    s.close()
    # stepping before every peer is connected is refused
    ps = partition.PartitionedSolver(parts[0], lib_path=emul_lib)
    ps.load_state(partition.split_state(ps.part, state, net.n_pollut))
    ps.set_inflows(**partition.split_inflows(ps.part, inflows, net.n_pollut))
    with pytest.raises(solver.SwbError, match="swb_partition_connect"):
        ps.run_steps(1, t_end)
    # a second attach, a bad peer rank and connecting twice are refused
    other = partition.PartitionedSolver(parts[1], lib_path=emul_lib)
    with pytest.raises(solver.SwbError, match="bad peer rank"):
        ps._chk(ps.lib.swb_partition_connect(ps._h, 0, C.create_string_buffer(other.export_handle(), 64)))
    ps.connect([None, other.export_handle()])
    with pytest.raises(solver.SwbError, match="already connected"):
        ps.connect([None, other.export_handle()])
    ps.close()
    other.close()
