"""Shared helpers of the parity tests.

Three ways to obtain the truth, all derived from the UNMODIFIED reference engine:
  * live: oracle/_ref/libswmm5.so stepped side by side with the solver (needs oracle/_ref);
  * golden: tests/golden/*.npz written by tests/golden/make_golden.py from the same engine;
  * oracle: oracle/libswmm_oracle.so, the C restatement (pinned against the two above).
The solver under test is loaded through the C-ABI: lib_path=None -> the CUDA library,
lib_path=EMUL_LIB -> the host emulation of the same kernels (CPU-only suite).
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import swmm_b200  # noqa: E402,F401
from swmm_b200 import abi, scenarios, solver  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
EMUL_LIB = os.path.join(EMUL_DIR, "libswb_emul.so")
CSRC = os.path.join(ROOT, "stormwater-management-model_b200", "csrc")
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

SNAP_FIELDS = ["SWB_NODE_NEW_DEPTH", "SWB_LINK_NEW_FLOW", "SWB_NODE_NEW_VOLUME", "SWB_NODE_OVERFLOW",
               "SWB_LINK_NEW_DEPTH", "SWB_LINK_NEW_VOLUME", "SWB_NODE_NEW_QUAL", "SWB_LINK_NEW_QUAL"]
FLOOR = {"SWB_NODE_NEW_DEPTH": 1e-4, "SWB_LINK_NEW_FLOW": 1e-4, "SWB_NODE_NEW_VOLUME": 1e-3,
         "SWB_NODE_OVERFLOW": 1e-4, "SWB_LINK_NEW_DEPTH": 1e-4, "SWB_LINK_NEW_VOLUME": 1e-3,
         "SWB_NODE_NEW_QUAL": 1e-6, "SWB_LINK_NEW_QUAL": 1e-6}


def build_emul() -> str:
    """Compile the host emulation of the device engine (test scaffolding) if it is stale."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    srcs += [os.path.join(EMUL_DIR, "emul_backend.cpp"), os.path.join(ROOT, "include", "swmm_b200.h")]
    if os.path.exists(EMUL_LIB) and all(os.path.getmtime(s) <= os.path.getmtime(EMUL_LIB) for s in srcs):
        return EMUL_LIB
    subprocess.run([HOST_CXX, "-O2", "-std=c++20", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                    "-Wno-unused-function", f"-I{CSRC}", f"-I{ROOT}/include",
                    os.path.join(EMUL_DIR, "emul_backend.cpp"), "-o", EMUL_LIB], check=True)
    return EMUL_LIB


def reference_available() -> bool:
    import refengine
    return refengine.available()


def cuda_available() -> bool:
    if not os.path.exists(solver.CUDA_LIB):
        return False
    try:
        return solver.load_library().swb_device_count() > 0
    except Exception:
        return False


# ---- named cases -----------------------------------------------------------------------------------
def case_inp(name: str) -> str:
    if name == "c1_tree":
        return scenarios.c1_tree_inp(scenarios.TreeSpec())
    if name == "c1_tree_slot":
        return scenarios.c1_tree_inp(scenarios.TreeSpec(surcharge="SLOT", peak_cfs=3.0, hours=3.0))
    if name.startswith("c2_grid"):
        # c2_grid<n>_<slot|extran>[_noq]
        parts = name.split("_")
        n = int(parts[1][4:])
        sur = parts[2].upper()
        hours = 2.0
        return scenarios.c2_grid_inp(scenarios.GridSpec(nx=n, ny=n, hours=hours, surcharge=sur,
                                                        pollutants="noq" not in parts))
    if name == "c3_mixed":
        return scenarios.c3_mixed_inp()
    if name == "c3_large":
        return scenarios.c3_large_inp()
    if name == "c3_large_6h":
        return scenarios.c3_large_inp(scenarios.C3Spec(hours=6.0))
    if name == "c3_large_nocontrols":
        return scenarios.c3_large_inp(scenarios.C3Spec(controls=False))
    if name == "c3b_shapes":
        return scenarios.c3b_shapes_inp()
    if name == "c3c_culverts_hw":
        return scenarios.c3c_culverts_inp("H-W")
    if name == "c3c_culverts_dw":
        return scenarios.c3c_culverts_inp("D-W")
    raise KeyError(name)


def rel_err(a: np.ndarray, ref: np.ndarray, floor: float) -> float:
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor)))


def grab_state(engine) -> dict:
    st = {}
    for f in solver.Solver.STATE_FIELDS:
        try:
            st[f] = engine.field(f)
        except KeyError:
            pass
    return st


def open_reference(inp_text: str, lib: str = "libswmm5.so"):
    import refengine
    d = tempfile.mkdtemp(prefix="swb_")
    path = os.path.join(d, "model.inp")
    with open(path, "w") as f:
        f.write(inp_text)
    e = refengine.RefEngine(lib)
    e.open(path)
    e.start()
    return e, d


def make_solver_from_engine(e, lib_path, n_members=1, member_scale=None, member_shift=None, full=False):
    """full=True: everything swb_run_steps can evaluate on the device comes from the shipped flatteners
    (patterns, pollutant inflow records, dry-weather flow, control rules, pump depths, timed outfall
    stages) instead of the plain FLOW-hydrograph extraction."""
    net = e.network()
    s = solver.Solver(net, n_members, lib_path=lib_path)
    s.load_state(grab_state(e))
    if full:
        s.set_inflows_desc(e.inflows_desc(), member_scale=member_scale, member_shift=member_shift)
        s.set_controls_desc(e.controls_desc())
    else:
        inf = e.inflows()
        s.set_inflows(member_scale=member_scale, member_shift=member_shift, **inf)
    return s


def lockstep_vs_reference(inp_text: str, lib_path, max_steps: int | None = None, every: int = 1,
                          n_members: int = 1, continuity: bool = False, full: bool = False) -> dict:
    """Step the reference and the solver (ensemble driver, one step per launch) side by side.
    continuity=True (whole runs only) adds the flow / quality continuity errors of the solver's
    device-side routing totals next to the reference's own swmm_getMassBalErr."""
    e, _ = open_reference(inp_text)
    out = None
    try:
        s = make_solver_from_engine(e, lib_path, n_members, full=full)
        init_storage = s.storage() if continuity else None
        ref_iters = 0
        t_end = e.total_duration_s()
        worst = {f: 0.0 for f in SNAP_FIELDS}
        worst_dt = 0.0
        steps = 0
        crit_checked = crit_bad = 0
        prev_crit = None
        while True:
            t = e.step()
            ref_iters += e.last_iterations()
            # the reference searched the step it just took on the state BEFORE it: compare with
            # the arg-min our previous step left behind
            if prev_crit is not None and steps >= 1:
                crit_checked += 1
                if prev_crit != e.last_critical():
                    crit_bad += 1
            s.run_steps(1, t_end)
            st0 = s.stats()[0]
            prev_crit = (st0.crit_node, st0.crit_link)
            steps += 1
            st = s.stats()
            t_ref = e.routing_time_ms() / 1000.0
            for k in range(n_members):
                worst_dt = max(worst_dt, abs(st[k].sim_time - t_ref))
            if steps % every == 0 or t == 0:
                for f in SNAP_FIELDS:
                    try:
                        r = e.field(f)
                    except KeyError:
                        continue
                    m = s.get_field(f)
                    for k in range(n_members):
                        worst[f] = max(worst[f], rel_err(m[k], r, FLOOR[f]))
            if t == 0 or (max_steps and steps >= max_steps):
                break
        out = {"steps": steps, "time_err_s": worst_dt, "crit_checked": crit_checked, "crit_mismatch": crit_bad, "iterations": int(st[0].iterations),
               "non_converged": int(st[0].non_converged), "ref_non_converged": e.non_converge_count()}
        out.update({"rel_" + f: v for f, v in worst.items()})
        out["max_rel"] = max(worst.values())
        out["ref_iterations"] = ref_iters
        if continuity:
            flow, qual = s.continuity(init_storage)
            out["flow_error_pct"] = float(flow[0])
            out["qual_error_pct"] = float(qual[0][np.argmax(np.abs(qual[0]))]) if qual.size else 0.0
        s.close()
    finally:
        e.end()
        if continuity and out is not None:
            _, out["ref_flow_error_pct"], out["ref_qual_error_pct"] = e.mass_bal_err()
        e.close()
    return out


# ---- golden fixtures ---------------------------------------------------------------------------------
def golden_path(name: str) -> str:
    return os.path.join(GOLDEN, name + ".npz")


def load_golden(name: str):
    net, extra = abi.Network.load(golden_path(name))
    return net, extra


def run_golden_case(name: str, lib_path, max_steps: int | None = None, n_members: int = 1) -> dict:
    """Replay a committed reference trajectory: every step's dt and the sampled snapshots."""
    net, g = load_golden(name)
    s = solver.Solver(net, n_members, lib_path=lib_path)
    state = {k[3:]: g[k] for k in g if k.startswith("s0_")}
    s.load_state(state)
    nP = net.n_pollut
    s.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"], ts_q=g["inf_ts_q"],
                  sfactor=g["inf_sfactor"], baseline=g["inf_baseline"],
                  concen=g["inf_concen"] if nP else None,
                  start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
    t_end = float(g["t_end"])
    ref_t = g["series_time"]
    ref_it = g["series_iters"]
    snap_steps = list(g["snap_steps"])
    n_total = len(ref_t)
    n_run = min(n_total, max_steps) if max_steps else n_total
    worst = {}
    time_err = 0.0
    iters_ok = True
    prev_it = 0
    for step in range(1, n_run + 1):
        s.run_steps(1, t_end)
        st = s.stats()
        for k in range(n_members):
            time_err = max(time_err, abs(st[k].sim_time - ref_t[step - 1]))
            iters_ok = iters_ok and (st[k].iterations - prev_it == ref_it[step - 1])
        prev_it = st[0].iterations
        if step in snap_steps:
            idx = snap_steps.index(step)
            for f in SNAP_FIELDS:
                key = f"snap_{f}"
                if key not in g:
                    continue
                r = g[key][idx]
                m = s.get_field(f)
                for k in range(n_members):
                    worst[f] = max(worst.get(f, 0.0), rel_err(m[k], r, FLOOR[f]))
    s.close()
    return {"steps": n_run, "time_err_s": time_err, "iters_match": iters_ok,
            "max_rel_depth": worst.get("SWB_NODE_NEW_DEPTH", 0.0),
            "max_rel_flow": worst.get("SWB_LINK_NEW_FLOW", 0.0),
            "max_rel_qual": max(worst.get("SWB_NODE_NEW_QUAL", 0.0), worst.get("SWB_LINK_NEW_QUAL", 0.0)),
            "max_rel": max(worst.values()) if worst else 0.0, "snapshots": len([x for x in snap_steps if x <= n_run])}


def batch_step_equals_sequential(lib_path, case="c2_grid12_slot", M=64, blocks=2, warm=25, steps=12):
    """swb_step_host_batch on member blocks cloned from one ensemble gives, block for block, what
    swb_step_host gives on the whole ensemble (host-chosen dt fed back every step)."""
    from swmm_b200 import solver
    net, g = load_golden(case)
    nP = net.n_pollut
    a = solver.Solver(net, M, lib_path=lib_path)
    a.load_state({k[3:]: g[k] for k in g if k.startswith("s0_")})
    a.set_inflows(node=g["inf_node"], ts_start=g["inf_ts_start"], ts_t=g["inf_ts_t"],
                  ts_q=g["inf_ts_q"], sfactor=g["inf_sfactor"], baseline=g["inf_baseline"],
                  concen=g["inf_concen"] if nP else None, member_scale=np.linspace(0.4, 1.6, M),
                  start_day=float(g["inf_start"][0]), start_secs=float(g["inf_start"][1]))
    a.run_steps(warm, 1e9)
    nb = M // blocks
    subs = [a.clone_members(b * nb, nb) for b in range(blocks)]
    lat = a.host_array((M, net.n_nodes))
    lat[:] = a.get_field("SWB_NODE_NEW_LATFLOW")
    conc = np.zeros((net.n_nodes, nP))
    conc[g["inf_node"]] = g["inf_concen"].reshape(-1, nP)
    load = a.host_array((M, net.n_nodes, max(nP, 1)))
    load[:] = 0.0
    if nP:
        load[:] = np.maximum(lat, 0.0)[:, :, None] * conc[None]
    dt = a.host_array((M,))
    dt[:] = [x.next_dt for x in a.stats()]
    out = {}
    for tag in ("a", "b"):
        out[tag] = dict(depth=a.host_array((M, net.n_nodes)), flow=a.host_array((M, net.n_links)),
                        next_dt=a.host_array((M,)), iters=a.host_array((M,), dtype=np.int32))
    for step in range(steps):
        oa, ob = out["a"], out["b"]
        a.step_host(lat, dt=dt, qual_load=load if nP else None, node_depth=oa["depth"],
                    link_flow=oa["flow"], next_dt=oa["next_dt"], iters=oa["iters"])
        ios = []
        for b in range(blocks):
            sl = slice(b * nb, (b + 1) * nb)
            ios.append(dict(latflow=lat[sl], dt=dt[sl], qual_load=load[sl] if nP else None,
                            node_depth=ob["depth"][sl], link_flow=ob["flow"][sl],
                            next_dt=ob["next_dt"][sl], iters=ob["iters"][sl]))
        solver.step_host_batch(subs, ios)
        for k in oa:
            assert np.array_equal(oa[k], ob[k]), (step, k)
        dt[:] = oa["next_dt"]
    assert len(set(out["a"]["iters"].tolist())) >= 1
    for b in range(blocks):
        sl = slice(b * nb, (b + 1) * nb)
        assert np.array_equal(subs[b].get_field("SWB_NODE_NEW_QUAL"),
                              a.get_field("SWB_NODE_NEW_QUAL", b * nb, nb))
    for s in subs:
        s.close()
    a.close()
