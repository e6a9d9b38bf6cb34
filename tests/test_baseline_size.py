"""Parity at BASELINE size, and the ensemble's equivalence contract:

  * the network builder bench.py uses equals the engine's own flattening of the same .inp;
  * member m of a config-4 ensemble == the reference run on the correspondingly perturbed .inp
    (scale factor and shifted hydrograph written into the file), for a random sample of members;
  * config 2 at its full size (100 x 100, 2 pollutants, SLOT) stepped side by side with the live
    reference for the whole simulation: same time steps, same Picard counts, 1e-6 on depths /
    flows / concentrations, continuity errors within 0.01 percentage points.

Truth is always the UNMODIFIED reference engine (oracle/_ref, prebuilt, travels to the GPU box).
CPU tests drive the host emulation of the device engine (bit-exact expected); -m gpu tests drive
the CUDA library (1e-6, north_star).
"""
import numpy as np
import pytest

import parity_common as pc
from swmm_b200 import network, scenarios, solver

TOL = 1e-6
TOL_PP = 0.01


# ---- the builder ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n,sur,scale,shift", [(12, "SLOT", 1.0, 0.0), (30, "EXTRAN", 1.0, 0.0),
                                               (12, "SLOT", 1.372911, 0.25)])
def test_network_builder_equals_engine_flattening(n, sur, scale, shift, emul_lib, have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref is not built")
    spec = scenarios.GridSpec(nx=n, ny=n, surcharge=sur, inflow_scale=scale, inflow_shift_h=shift)
    built = network.build_grid(spec, lib_path=emul_lib)
    e, _ = pc.open_reference(scenarios.c2_grid_inp(spec))
    try:
        ref = e.network()
        assert ref.scalars == built.net.scalars
        for k, v in ref.arrays.items():
            assert np.array_equal(np.asarray(v), np.asarray(built.net.arrays[k])), k
        for k, v in ref.options.items():
            if k in ("reserved", "reserved0"):
                continue
            assert built.net.options[k] == v, (k, built.net.options[k], v)
        inf = e.inflows()
        for k, v in inf.items():
            b = built.inflows[k]
            if v is None:
                assert b is None, k
            else:
                assert np.array_equal(np.asarray(v), np.asarray(b)), k
        st = pc.grab_state(e)
        for k, v in built.state0.items():
            assert np.array_equal(st[k], v), k
        for k, v in st.items():
            if k not in built.state0:
                assert not np.any(v), k            # everything else starts at zero
    finally:
        e.end()
        e.close()


# ---- sampled members against the reference on the perturbed .inp ------------------------------------
def reference_member(spec, every):
    """Reference trajectory of one perturbed model: time after every step, Picard count of every
    step, snapshots of the parity fields every `every` steps and at the end."""
    e, _ = pc.open_reference(scenarios.c2_grid_inp(spec))
    times, iters, snaps = [], [], {}
    try:
        step = 0
        while True:
            t = e.step()
            step += 1
            times.append(e.routing_time_ms() / 1000.0)
            iters.append(e.last_iterations())
            if step % every == 0 or t == 0:
                snaps[step] = {f: e.field(f).copy() for f in pc.SNAP_FIELDS if _has(e, f)}
            if t == 0:
                break
        nonconv = e.non_converge_count()
    finally:
        e.end()
        e.close()
    return dict(times=np.array(times), iters=np.array(iters), snaps=snaps, non_converged=nonconv)


def _has(e, f):
    try:
        e.field(f)
        return True
    except KeyError:
        return False


def members_vs_reference(lib_path, nx, M, sample, every, hours=2.0, surcharge="SLOT"):
    scale, shift_h = scenarios.c4_members(4096, 2024)
    base = scenarios.GridSpec(nx=nx, ny=nx, hours=hours, surcharge=surcharge)
    refs = {m: reference_member(scenarios.GridSpec(nx=nx, ny=nx, hours=hours, surcharge=surcharge,
                                                   inflow_scale=float(scale[m]),
                                                   inflow_shift_h=float(shift_h[m])), every)
            for m in sample}
    case = network.build_grid(base, lib_path=lib_path)
    s = solver.Solver(case.net, M, lib_path=lib_path)
    s.load_state(case.state0)
    s.set_inflows(member_scale=scale[:M], member_shift=shift_h[:M] / 24.0, **case.inflows)
    n_max = max(len(r["times"]) for r in refs.values())
    worst = {f: 0.0 for f in pc.SNAP_FIELDS}
    time_err, iters_ok = 0.0, True
    prev = {m: 0 for m in sample}
    for step in range(1, n_max + 1):
        s.run_steps(1, case.t_end)
        st = s.stats()
        for m in sample:
            r = refs[m]
            if step > len(r["times"]):
                continue
            time_err = max(time_err, abs(st[m].sim_time - r["times"][step - 1]))
            iters_ok = iters_ok and (st[m].iterations - prev[m] == r["iters"][step - 1])
            prev[m] = st[m].iterations
            if step in r["snaps"]:
                for f, ref in r["snaps"][step].items():
                    worst[f] = max(worst[f], pc.rel_err(s.get_field(f, m, 1)[0], ref, pc.FLOOR[f]))
    st = s.stats()
    out = dict(steps={m: len(refs[m]["times"]) for m in sample}, time_err_s=time_err, iters_match=iters_ok,
               non_converged_match=all(st[m].non_converged == refs[m]["non_converged"] for m in sample),
               steps_match=all(st[m].steps == len(refs[m]["times"]) for m in sample),
               max_rel=max(worst.values()), worst=worst,
               scales=[float(scale[m]) for m in sample], shifts_h=[float(shift_h[m]) for m in sample])
    s.close()
    return out


def test_emulated_ensemble_members_equal_reference_on_perturbed_inp(emul_lib, have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref is not built")
    r = members_vs_reference(emul_lib, nx=12, M=32, sample=[0, 13, 31], every=150)
    print(r)
    assert r["time_err_s"] < 1e-9 and r["iters_match"] and r["steps_match"], r
    assert r["max_rel"] == 0.0, r          # the host build of the device engine is bit-exact


@pytest.mark.gpu
def test_cuda_ensemble_members_equal_reference_on_perturbed_inp(cuda_lib, have_reference):
    """8 random members of the 4 096-member config-4 ensemble, each against the
    reference's own run of its perturbed .inp, inside a 64-member lockstep ensemble (30 x 30)."""
    assert have_reference, "oracle/_ref must travel to the GPU box"
    sample = sorted(int(x) for x in np.random.default_rng(7).choice(64, 8, replace=False))
    r = members_vs_reference(None, nx=30, M=64, sample=sample, every=100)
    print(r)
    assert r["time_err_s"] < 1e-9 and r["iters_match"] and r["steps_match"], r
    assert r["non_converged_match"], r
    assert r["max_rel"] <= TOL, r


# ---- config 2 at its stated size -----------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_c2_full_size_lockstep_vs_live_reference(cuda_lib, have_reference):
    """100 x 100 grid (19 801 conduits), 2 pollutants, SLOT, the whole 2 h (about 1 500 steps)."""
    assert have_reference, "oracle/_ref must travel to the GPU box"
    r = pc.lockstep_vs_reference(pc.case_inp("c2_grid100_slot"), None, every=100, continuity=True)
    print(r)
    assert r["steps"] > 1000
    assert r["time_err_s"] < 1e-9 and r["max_rel"] <= TOL, r
    assert r["iterations"] == r["ref_iterations"], r
    assert r["non_converged"] == r["ref_non_converged"], r
    assert r["crit_mismatch"] == 0, r
    assert abs(r["flow_error_pct"] - r["ref_flow_error_pct"]) <= TOL_PP, r
    assert abs(r["qual_error_pct"] - r["ref_qual_error_pct"]) <= TOL_PP, r


def test_emulated_c2_mid_size_lockstep_with_continuity(emul_lib, have_reference):
    """The same comparison at 30 x 30 in the host build (bit-exact), so the continuity plumbing of
    the full-size GPU test is exercised on CPU."""
    if not have_reference:
        pytest.skip("oracle/_ref is not built")
    r = pc.lockstep_vs_reference(pc.case_inp("c2_grid30_slot"), emul_lib, every=200, continuity=True,
                                 max_steps=None)
    print(r)
    assert r["time_err_s"] < 1e-9 and r["max_rel"] == 0.0, r
    assert r["iterations"] == r["ref_iterations"], r
    assert abs(r["flow_error_pct"] - r["ref_flow_error_pct"]) <= TOL_PP, r
    assert abs(r["qual_error_pct"] - r["ref_qual_error_pct"]) <= TOL_PP, r
