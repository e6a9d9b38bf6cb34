"""Host side of the C-ABI: thin ctypes wrapper over libswmm_b200.so (include/swmm_b200.h).

There is no CPU fallback: if the CUDA library is missing or no device is present every call
raises.  (tests/emul builds a host emulation of the same ABI for the CPU-only test suite; it is
loaded only when a test passes its path explicitly.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_PKG = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB = os.path.join(_PKG, "csrc", "libswmm_b200.so")

_P_D = C.POINTER(C.c_double)
_P_I = C.POINTER(C.c_int)


class SwbError(RuntimeError):
    pass


def load_library(path: str | None = None) -> C.CDLL:
    path = path or CUDA_LIB
    if not os.path.exists(path):
        raise SwbError(f"{path} not found: build the CUDA extension first "
                       f"(python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(path)
    lib.swb_last_error.restype = C.c_char_p
    lib.swb_network_create.argtypes = [C.POINTER(abi.NetworkDesc), C.POINTER(abi.Options), C.c_int,
                                       C.POINTER(C.c_void_p)]
    lib.swb_network_destroy.argtypes = [C.c_void_p]
    lib.swb_solver_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.swb_solver_destroy.argtypes = [C.c_void_p]
    lib.swb_set_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _P_D]
    lib.swb_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _P_D]
    lib.swb_broadcast_field.argtypes = [C.c_void_p, C.c_int, _P_D]
    lib.swb_set_climate.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.swb_qual_init.argtypes = [C.c_void_p, _P_D]
    lib.swb_old_state_swap.argtypes = [C.c_void_p, _P_D, C.c_int]
    lib.swb_dynwave_execute.argtypes = [C.c_void_p, _P_D, _P_I]
    lib.swb_qualrout_execute.argtypes = [C.c_void_p, _P_D]
    lib.swb_get_routing_step.argtypes = [C.c_void_p, C.c_double, _P_D]
    lib.swb_set_inflows.argtypes = [C.c_void_p, C.POINTER(abi.InflowDesc)]
    lib.swb_set_controls.argtypes = [C.c_void_p, C.POINTER(abi.ControlsDesc)]
    lib.swb_permute_members.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.swb_run_steps.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.swb_get_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.MemberStats)]
    lib.swb_get_massbal.argtypes = [C.c_void_p, C.c_int, C.c_int, _P_D, _P_D, _P_D]
    lib.swb_get_routing_totals.argtypes = [C.c_void_p, C.c_int, C.c_int, _P_D, _P_D]
    lib.swb_conduit_updates.argtypes = [C.c_void_p]
    lib.swb_conduit_updates.restype = C.c_longlong
    lib.swb_set_staged_min_members.argtypes = [C.c_int]
    lib.swb_set_staged_min_members.restype = C.c_int
    lib.swb_launch_count.argtypes = [C.c_void_p]
    lib.swb_launch_count.restype = C.c_longlong
    lib.swb_last_kernel_ms.argtypes = [C.c_void_p]
    lib.swb_last_kernel_ms.restype = C.c_double
    lib.swb_sync.argtypes = [C.c_void_p]
    lib.swb_get_phase_times.argtypes = [C.c_void_p, _P_D, C.c_int, C.c_int]
    lib.swb_enable_statistics.argtypes = [C.c_void_p, C.c_double]
    lib.swb_get_node_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, _P_D]
    lib.swb_get_link_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, _P_D]
    lib.swb_get_system_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, _P_D]
    lib.swb_debug_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.swb_host_alloc.argtypes = [C.c_ulonglong]
    lib.swb_host_alloc.restype = C.c_void_p
    lib.swb_host_free.argtypes = [C.c_void_p]
    lib.swb_step_host.argtypes = [C.c_void_p, C.POINTER(abi.StepIO)]
    lib.swb_step_host_batch.argtypes = [C.POINTER(C.c_void_p), C.POINTER(abi.StepIO), C.c_int]
    lib.swb_get_results.argtypes = [C.c_void_p, _P_D, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.swb_xsect_eval.argtypes = [C.c_int, C.c_int, C.c_int, _P_D, C.c_int, _P_D, _P_D]
    return lib


XS_FN = {"AofY": 0, "WofY": 1, "RofY": 2, "YofA": 3, "RofA": 4, "SofA": 5, "AofS": 6, "dSdA": 7,
         "Ycrit": 8}


def xsect_eval(fn: str, xs_type: int, params, args, device: int = 0, lib_path: str | None = None):
    """Evaluate one geometry function on the device (known-answer hook, swb_xsect_eval)."""
    lib = load_library(lib_path)
    p = np.ascontiguousarray(params, dtype=np.float64)
    a = np.ascontiguousarray(args, dtype=np.float64)
    out = np.zeros_like(a)
    rc = lib.swb_xsect_eval(device, XS_FN[fn], int(xs_type), p.ctypes.data_as(_P_D), a.size,
                            a.ctypes.data_as(_P_D), out.ctypes.data_as(_P_D))
    if rc:
        raise SwbError(f"swb error {rc}: {lib.swb_last_error().decode()}")
    return out


def step_host_batch(solvers, ios):
    """One routing step of several solvers (member blocks of one ensemble) as a pipelined batch
    (swb_step_host_batch): copies of one block overlap the routing kernel of another.
    ios: one dict of step_host keyword arguments (incl. "latflow") per solver."""
    n = len(solvers)
    arr = (abi.StepIO * n)()
    keep = []
    for i, kw in enumerate(ios):
        io, k = Solver._step_io(**kw)
        arr[i] = io
        keep.append(k)
    handles = (C.c_void_p * n)(*[s._h for s in solvers])
    rc = solvers[0].lib.swb_step_host_batch(handles, arr, n)
    if rc:
        raise SwbError(f"swb error {rc}: {solvers[0].lib.swb_last_error().decode()}")


class Solver:
    """M lockstep members of one flat network on one GPU."""

    def __init__(self, net: abi.Network, n_members: int = 1, device: int = 0, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        self._lib_path, self._device = lib_path, device
        self.net = net
        self.M = n_members
        self._desc, self._opt = net.to_c()        # keep alive
        self._hnet = C.c_void_p()
        self._chk(self.lib.swb_network_create(C.byref(self._desc), C.byref(self._opt), device,
                                              C.byref(self._hnet)))
        self._h = C.c_void_p()
        self._chk(self.lib.swb_solver_create(self._hnet, n_members, C.byref(self._h)))
        self._keep = []
        self._pinned = []

    def _chk(self, rc: int):
        if rc:
            raise SwbError(f"swb error {rc}: {self.lib.swb_last_error().decode()}")

    def close(self):
        for p in getattr(self, "_pinned", []):
            self.lib.swb_host_free(p)
        self._pinned = []
        if self._h:
            self.lib.swb_solver_destroy(self._h)
            self._h = C.c_void_p()
        if self._hnet:
            self.lib.swb_network_destroy(self._hnet)
            self._hnet = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- fields ------------------------------------------------------------------------------
    def _fid(self, f):
        return abi.FIELD[f] if isinstance(f, str) else int(f)

    def _items(self, fid: int) -> int:
        n = self.net.n_nodes if abi.is_node_field(fid) else self.net.n_links
        return n * abi.field_width(fid, self.net.n_pollut)

    def set_field(self, field, values, member0: int = 0):
        fid = self._fid(field)
        a = np.ascontiguousarray(values, dtype=np.float64)
        items = self._items(fid)
        if items == 0:
            return
        nm = a.size // items
        assert nm * items == a.size, (field, a.size, items)
        self._chk(self.lib.swb_set_field(self._h, fid, member0, nm, a.ctypes.data_as(_P_D)))

    def broadcast_field(self, field, values):
        fid = self._fid(field)
        a = np.ascontiguousarray(values, dtype=np.float64)
        if self._items(fid) == 0:
            return
        assert a.size == self._items(fid), (field, a.size, self._items(fid))
        self._chk(self.lib.swb_broadcast_field(self._h, fid, a.ctypes.data_as(_P_D)))

    def get_field(self, field, member0: int = 0, n_members: int | None = None) -> np.ndarray:
        fid = self._fid(field)
        nm = self.M - member0 if n_members is None else n_members
        items = self._items(fid)
        out = np.zeros((nm, items), dtype=np.float64)
        if items:
            self._chk(self.lib.swb_get_field(self._h, fid, member0, nm, out.ctypes.data_as(_P_D)))
        return out

    # ---- the reference's per-step calls ---------------------------------------------------------
    def _dt(self, dt):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, dtype=np.float64), (self.M,)))
        return a, a.ctypes.data_as(_P_D)

    def old_state_swap(self, dt, with_quality: bool = False):
        a, p = self._dt(dt)
        self._chk(self.lib.swb_old_state_swap(self._h, p, int(with_quality)))

    def dynwave_execute(self, dt) -> np.ndarray:
        a, p = self._dt(dt)
        it = np.zeros(self.M, dtype=np.int32)
        self._chk(self.lib.swb_dynwave_execute(self._h, p, it.ctypes.data_as(_P_I)))
        return it

    def qualrout_execute(self, dt):
        a, p = self._dt(dt)
        self._chk(self.lib.swb_qualrout_execute(self._h, p))

    def get_routing_step(self, fixed_step: float) -> np.ndarray:
        out = np.zeros(self.M, dtype=np.float64)
        self._chk(self.lib.swb_get_routing_step(self._h, fixed_step, out.ctypes.data_as(_P_D)))
        return out

    def qual_init(self, init_concen):
        a = np.ascontiguousarray(init_concen, dtype=np.float64)
        self._chk(self.lib.swb_qual_init(self._h, a.ctypes.data_as(_P_D)))

    def set_climate(self, evap_rate: float, hydcon_factor: float = 1.0):
        self._chk(self.lib.swb_set_climate(self._h, evap_rate, hydcon_factor))

    # ---- ensemble driver ----------------------------------------------------------------------
    def set_inflows(self, node, ts_start, ts_t, ts_q, sfactor, baseline, concen=None,
                    member_scale=None, member_shift=None, start_day=0.0, start_secs=0.0):
        d = abi.InflowDesc()
        arrs = {
            "node": np.ascontiguousarray(node, dtype=np.int32),
            "ts_start": np.ascontiguousarray(ts_start, dtype=np.int32),
            "ts_t": np.ascontiguousarray(ts_t, dtype=np.float64),
            "ts_q": np.ascontiguousarray(ts_q, dtype=np.float64),
            "sfactor": np.ascontiguousarray(sfactor, dtype=np.float64),
            "baseline": np.ascontiguousarray(baseline, dtype=np.float64),
        }
        n = arrs["node"].size
        if concen is not None:
            arrs["concen"] = np.ascontiguousarray(concen, dtype=np.float64)
            assert arrs["concen"].size == n * self.net.n_pollut
        if member_scale is not None:
            arrs["member_scale"] = np.ascontiguousarray(member_scale, dtype=np.float64)
            assert arrs["member_scale"].size == self.M
        if member_shift is not None:
            arrs["member_shift"] = np.ascontiguousarray(member_shift, dtype=np.float64)
            assert arrs["member_shift"].size == self.M
        d.n_inflow_nodes = n
        d.n_ts_pts = arrs["ts_t"].size
        d.start_day = float(start_day)
        d.start_secs = float(start_secs)
        for name, base in abi.INFLOW_ARRAYS:
            if name in arrs:
                ct = C.c_int if base == "int" else C.c_double
                setattr(d, name, arrs[name].ctypes.data_as(C.POINTER(ct)))
        self._keep.append(arrs)
        self._chk(self.lib.swb_set_inflows(self._h, C.byref(d)))

    def set_inflows_desc(self, desc, member_scale=None, member_shift=None):
        """swb_set_inflows with a ready swb_inflow_desc (e.g. from the seam's flattener); the per-member
        scale / shift of the FLOW hydrographs are filled in here."""
        d = abi.InflowDesc()
        C.memmove(C.byref(d), C.byref(desc), C.sizeof(d))
        keep = {}
        for name, arr in (("member_scale", member_scale), ("member_shift", member_shift)):
            if arr is not None:
                keep[name] = np.ascontiguousarray(arr, dtype=np.float64)
                assert keep[name].size == self.M
                setattr(d, name, keep[name].ctypes.data_as(_P_D))
        self._chk(self.lib.swb_set_inflows(self._h, C.byref(d)))

    def set_controls_desc(self, desc):
        """Control rules / pump depths / timed outfall stages for run_steps (swb_set_controls)."""
        self._chk(self.lib.swb_set_controls(self._h, C.byref(desc)))

    def permute_members(self, perm):
        """Re-enumerate the members: afterwards member i is what member perm[i] was (swb_permute_members)."""
        p = np.ascontiguousarray(perm, dtype=np.int32)
        assert p.size == self.M
        self._chk(self.lib.swb_permute_members(self._h, p.ctypes.data_as(C.POINTER(C.c_int))))

    def run_steps(self, n_steps: int, t_end: float):
        self._chk(self.lib.swb_run_steps(self._h, n_steps, t_end))

    def stats(self, member0: int = 0, n_members: int | None = None):
        nm = self.M - member0 if n_members is None else n_members
        arr = (abi.MemberStats * nm)()
        self._chk(self.lib.swb_get_stats(self._h, member0, nm, arr))
        return arr

    def massbal(self):
        nP = max(self.net.n_pollut, 1)
        r, sp, f = (np.zeros((self.M, nP)) for _ in range(3))
        self._chk(self.lib.swb_get_massbal(self._h, 0, self.M, r.ctypes.data_as(_P_D),
                                           sp.ctypes.data_as(_P_D), f.ctypes.data_as(_P_D)))
        return {"reacted": r, "seepage": sp, "final_storage": f}

    FLOW_TERMS = ["ex_inflow", "flooding", "outflow", "evap_loss", "seep_loss"]
    QUAL_TERMS = ["ex_inflow", "flooding", "outflow", "reacted", "seepage", "final_storage"]

    def routing_totals(self):
        """Per-member routing totals kept on the device (swb_get_routing_totals): flow volumes
        {term: [M]} and pollutant masses {term: [M, P]}."""
        nP = self.net.n_pollut
        fl = np.zeros((self.M, len(self.FLOW_TERMS)))
        ql = np.zeros((self.M, max(nP, 1), len(self.QUAL_TERMS)))
        self._chk(self.lib.swb_get_routing_totals(self._h, 0, self.M, fl.ctypes.data_as(_P_D),
                                                  ql.ctypes.data_as(_P_D) if nP else None))
        return ({k: fl[:, i].copy() for i, k in enumerate(self.FLOW_TERMS)},
                {k: ql[:, :nP, i].copy() for i, k in enumerate(self.QUAL_TERMS)})

    def storage(self):
        """Stored volume [M] and pollutant mass [M, P] in nodes + links (massbal_getStorage
        massbal.c:634-662, massbal_getStoredMass :1025-1047)."""
        nP = self.net.n_pollut
        nv, lv = self.get_field("SWB_NODE_NEW_VOLUME"), self.get_field("SWB_LINK_NEW_VOLUME")
        vol = nv.sum(axis=1) + lv.sum(axis=1)
        mass = np.zeros((self.M, nP))
        if nP:
            nq = self.get_field("SWB_NODE_NEW_QUAL").reshape(self.M, -1, nP)
            lq = self.get_field("SWB_LINK_NEW_QUAL").reshape(self.M, -1, nP)
            mass = (nv[:, :, None] * nq).sum(axis=1) + (lv[:, :, None] * lq).sum(axis=1)
        return vol, mass

    def continuity(self, init_storage):
        """Flow-routing and quality continuity errors in percent for every member, computed like
        massbal_getFlowError / massbal_getQualError (massbal.c:858-960).  init_storage = storage()
        taken before the first step.  Returns (flow_pct[M], qual_pct[M, P])."""
        fl, ql = self.routing_totals()
        v0, w0 = init_storage
        v1, w1 = self.storage()

        def pct(tin, tout, small):
            out = np.zeros_like(tin)
            for idx in np.ndindex(tin.shape):
                a, b = tin[idx], tout[idx]
                if abs(a - b) < small:
                    out[idx] = 1.0e-6
                elif abs(a) > 0.0:
                    out[idx] = 100.0 * (1.0 - b / a)
                elif abs(b) > 0.0:
                    out[idx] = 100.0 * (a / b - 1.0)
            return out
        ex, of = fl["ex_inflow"], fl["outflow"]
        tin = v0 + np.where(ex >= 0.0, ex, 0.0) + np.where(of < 0.0, -of, 0.0)
        tout = v1 + fl["flooding"] + fl["evap_loss"] + fl["seep_loss"] + np.where(ex < 0.0, -ex, 0.0) \
            + np.where(of >= 0.0, of, 0.0)
        flow_pct = pct(tin, tout, 1.0)
        qin = ql["ex_inflow"] + w0
        qout = ql["flooding"] + ql["outflow"] + ql["reacted"] + ql["seepage"] + ql["final_storage"] + w1
        return flow_pct, pct(qin, qout, 0.001)

    PHASES = ["prologue", "links", "regulators", "nodes", "control", "epilogue", "qual_nodes",
              "qual_links", "next_dt", "halo", "halo_wait"]

    def phase_times(self, reset: bool = True) -> dict:
        ms = np.zeros(len(self.PHASES))
        self._chk(self.lib.swb_get_phase_times(self._h, ms.ctypes.data_as(_P_D), ms.size, int(reset)))
        return dict(zip(self.PHASES, ms.tolist()))

    def results(self, f, member0: int = 0, n_members: int | None = None, nodes: bool = True, links: bool = True):
        """Report-time float32 records (swb_get_results): f = weighting factor per member (scalar or
        [M]).  Returns (node[nm, n_nodes, 6 + P], link[nm, n_links, 5 + P]) in the reference's .out
        variable order (enums.h:200-219)."""
        nm = self.M - member0 if n_members is None else n_members
        fa = np.ascontiguousarray(np.broadcast_to(np.asarray(f, dtype=np.float64), (self.M,)))
        P = 0 if self.net.options["ignore_quality"] else self.net.n_pollut
        nd = np.zeros((nm, self.net.n_nodes, 6 + P), dtype=np.float32) if nodes else None
        ld = np.zeros((nm, self.net.n_links, 5 + P), dtype=np.float32) if links else None
        pf = C.POINTER(C.c_float)
        self._chk(self.lib.swb_get_results(self._h, fa.ctypes.data_as(_P_D), member0, nm,
                                           nd.ctypes.data_as(pf) if nodes else None,
                                           ld.ctypes.data_as(pf) if links else None))
        return nd, ld

    # ---- per-object statistics (stats.c), kept per member on the device -----------------------
    def enable_statistics(self, report_start_s: float = 0.0):
        self._chk(self.lib.swb_enable_statistics(self._h, float(report_start_s)))

    def statistics(self, member0: int = 0, n_members: int | None = None):
        """(node[nm, planes, n_nodes], link[nm, planes, n_links], system[nm, planes]); plane ids are
        abi.NODE_STAT / abi.LINK_STAT / abi.SYSTEM_STAT (include/swmm_b200.h)."""
        nm = self.M - member0 if n_members is None else n_members
        nd = np.zeros((nm, abi.NODE_STAT["SWB_NS_PLANES"] + self.net.n_pollut, self.net.n_nodes))
        ld = np.zeros((nm, abi.LINK_STAT["SWB_LS_PLANES"], self.net.n_links))
        sd = np.zeros((nm, abi.SYSTEM_STAT["SWB_SS_PLANES"]))
        self._chk(self.lib.swb_get_node_stats(self._h, member0, nm, nd.ctypes.data_as(_P_D)))
        self._chk(self.lib.swb_get_link_stats(self._h, member0, nm, ld.ctypes.data_as(_P_D)))
        self._chk(self.lib.swb_get_system_stats(self._h, member0, nm, sd.ctypes.data_as(_P_D)))
        return nd, ld, sd

    def debug_run(self, phases: int, n_steps: int, debug: int = 0, profile: bool = False):
        """Profiling aid (swb_debug_run): phase mask + debug switches, see include/swmm_b200.h."""
        self._chk(self.lib.swb_debug_run(self._h, phases, n_steps, debug, int(profile)))

    def conduit_updates(self) -> int:
        return int(self.lib.swb_conduit_updates(self._h))

    def set_staged_min_members(self, n: int) -> int:
        """Process-wide switch between the staged kernel chain (ensembles of >= n members) and the
        persistent kernel (swb_set_staged_min_members); returns the previous value."""
        return int(self.lib.swb_set_staged_min_members(int(n)))

    def launch_count(self) -> int:
        return int(self.lib.swb_launch_count(self._h))

    def last_kernel_ms(self) -> float:
        return float(self.lib.swb_last_kernel_ms(self._h))

    def sync(self):
        self._chk(self.lib.swb_sync(self._h))

    # ---- host-buffer step (ensemble form of the seam's per-step exchange) -----------------------
    def host_array(self, shape, dtype=np.float64) -> np.ndarray:
        """numpy array over pinned host memory (swb_host_alloc); freed with the solver."""
        n = int(np.prod(shape))
        nbytes = max(n * np.dtype(dtype).itemsize, 8)
        ptr = self.lib.swb_host_alloc(nbytes)
        if not ptr:
            raise SwbError("swb_host_alloc failed")
        self._pinned.append(ptr)
        buf = (C.c_char * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    @staticmethod
    def _step_io(latflow, dt=None, node_losses=None, qual_load=None, node_depth=None,
                 link_flow=None, next_dt=None, iters=None):
        io = abi.StepIO()
        keep = []

        def ptr(a, ct):
            if a is None:
                return None
            assert a.flags["C_CONTIGUOUS"]
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(ct))
        io.dt = ptr(dt, C.c_double)
        io.latflow = ptr(latflow, C.c_double)
        io.node_losses = ptr(node_losses, C.c_double)
        io.qual_load = ptr(qual_load, C.c_double)
        io.node_depth = ptr(node_depth, C.c_double)
        io.link_flow = ptr(link_flow, C.c_double)
        io.next_dt = ptr(next_dt, C.c_double)
        io.iters = ptr(iters, C.c_int)
        return io, keep

    def step_host(self, latflow, **kw):
        """One routing step with host buffers (swb_step_host); keywords as in swb_step_io."""
        io, keep = self._step_io(latflow, **kw)
        self._chk(self.lib.swb_step_host(self._h, C.byref(io)))

    def clone_members(self, member0: int, n_members: int) -> "Solver":
        """A new solver holding members [member0, member0 + n_members) of this one (state fields
        only; per-member clocks stay with the parent: drive the clone with explicit dt)."""
        c = Solver(self.net, n_members, device=self._device, lib_path=self._lib_path)
        for f in self.STATE_FIELDS:
            if f == "SWB_COND_Q2":
                continue
            c.set_field(f, self.get_field(f, member0, n_members))
        return c

    # ---- convenience ---------------------------------------------------------------------------
    STATE_FIELDS = [
        "SWB_NODE_NEW_DEPTH", "SWB_NODE_OLD_DEPTH", "SWB_NODE_NEW_VOLUME", "SWB_NODE_OLD_VOLUME",
        "SWB_NODE_NEW_LATFLOW", "SWB_NODE_LOSSES", "SWB_NODE_INFLOW", "SWB_NODE_OUTFLOW",
        "SWB_NODE_OVERFLOW", "SWB_NODE_OLD_NET_INFLOW", "SWB_NODE_OUTFALL_STAGE",
        "SWB_NODE_STORAGE_EVAP_LOSS", "SWB_NODE_STORAGE_EXFIL_LOSS", "SWB_NODE_HRT",
        "SWB_NODE_NEW_QUAL", "SWB_NODE_OLD_QUAL", "SWB_NODE_OLD_LATFLOW", "SWB_NODE_OLD_INFLOW",
        "SWB_LINK_NEW_FLOW", "SWB_LINK_OLD_FLOW", "SWB_LINK_NEW_DEPTH", "SWB_LINK_OLD_DEPTH",
        "SWB_LINK_NEW_VOLUME", "SWB_LINK_OLD_VOLUME", "SWB_LINK_SETTING", "SWB_LINK_TARGET_SETTING",
        "SWB_LINK_DQDH", "SWB_LINK_FROUDE", "SWB_LINK_FLOW_CLASS", "SWB_LINK_SURF_AREA1",
        "SWB_LINK_SURF_AREA2", "SWB_LINK_NORMAL_FLOW", "SWB_LINK_INLET_CONTROL", "SWB_COND_A1",
        "SWB_COND_A2", "SWB_COND_Q1", "SWB_COND_Q2", "SWB_COND_FULL_STATE",
        "SWB_COND_CAPACITY_LIMITED", "SWB_COND_EVAP_LOSS", "SWB_COND_SEEP_LOSS", "SWB_ORIF_CORIF",
        "SWB_ORIF_CWEIR", "SWB_ORIF_HCRIT", "SWB_REG_SURF_AREA", "SWB_WEIR_CSURCHARGE",
        "SWB_LINK_NEW_QUAL", "SWB_LINK_OLD_QUAL", "SWB_LINK_TOTAL_LOAD",
    ]

    def load_state(self, state: dict):
        """Broadcast a single-member state image (field name -> array) to every member."""
        for k, v in state.items():
            self.broadcast_field(k, v)
