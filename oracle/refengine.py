"""TEST INFRASTRUCTURE: ctypes driver for the unmodified reference engine (oracle/_ref).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  Drives swmm_open/start/step/end/close (swmm5.h:129-151) and reads the engine's global
objects through oracle/_ref/librefhook.so.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def _abi():
    name = "swmm_b200"
    if name not in sys.modules:
        root = os.path.dirname(HERE)
        sys.path.insert(0, root)
        import swmm_b200  # noqa: F401  (root shim registers the package)
    return sys.modules["swmm_b200.abi"] if "swmm_b200.abi" in sys.modules else \
        __import__("swmm_b200.abi", fromlist=["x"])


def available(lib: str = "libswmm5.so") -> bool:
    return os.path.exists(os.path.join(REFDIR, lib)) and \
        os.path.exists(os.path.join(REFDIR, "librefhook.so")) and \
        os.path.exists(os.path.join(REFDIR, "librefcount.so"))


class RefEngine:
    """One reference simulation (the engine is global-state: one per process at a time)."""

    def __init__(self, lib: str = "libswmm5.so"):
        self.abi = _abi()
        mode = C.RTLD_GLOBAL
        # the counter interposer must enter the global scope BEFORE the engine (see refcount.c)
        self.count = C.CDLL(os.path.join(REFDIR, "librefcount.so"), mode=mode)
        self.count.refcount_get.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                            C.POINTER(C.c_double)]
        self.lib = C.CDLL(os.path.join(REFDIR, lib), mode=mode)
        self.hook = C.CDLL(os.path.join(REFDIR, "librefhook.so"), mode=mode)
        L = self.lib
        self.count.refcount_bind.argtypes = [C.c_void_p] * 4
        self.count.refcount_bind(*[C.cast(getattr(L, n), C.c_void_p) for n in
                                   ("dynwave_execute", "dynwave_getRoutingStep", "qualrout_execute",
                                    "routing_execute")])
        self.count.refcount_bind_crit.argtypes = [C.c_void_p]
        self.count.refcount_bind_crit(C.cast(L.stats_updateCriticalTimeCount, C.c_void_p))
        L.swmm_getValue.restype = C.c_double
        L.swmm_getValue.argtypes = [C.c_int, C.c_int]
        L.swmm_step.argtypes = [C.POINTER(C.c_double)]
        L.swmm_open.argtypes = [C.c_char_p] * 3
        L.swmm_getMassBalErr.argtypes = [C.POINTER(C.c_float)] * 3
        for fn in ("dynwave_getRoutingStep",):
            getattr(L, fn).restype = C.c_double
            getattr(L, fn).argtypes = [C.c_double]
        L.dynwave_execute.argtypes = [C.c_double]
        L.qualrout_execute.argtypes = [C.c_double]
        H = self.hook
        H.refhook_network.restype = C.POINTER(self.abi.NetworkDesc)
        H.refhook_options.restype = C.POINTER(self.abi.Options)
        H.refhook_get_field.argtypes = [C.c_int, C.POINTER(C.c_double)]
        H.refhook_set_field.argtypes = [C.c_int, C.POINTER(C.c_double)]
        H.refhook_new_routing_time.restype = C.c_double
        H.refhook_inflows.restype = C.POINTER(self.abi.InflowDesc)
        H.refhook_total_duration.restype = C.c_double
        H.refhook_inflows_full.restype = C.POINTER(self.abi.InflowDesc)
        H.refhook_controls.restype = C.POINTER(self.abi.ControlsDesc)
        H.refhook_xsect_eval.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
        H.refhook_xsect_set.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double,
                                        C.POINTER(C.c_double)]
        self.is_open = False

    # ---- swmm5.h API -----------------------------------------------------------------
    def open(self, inp: str, rpt: str | None = None, out: str | None = None):
        inp = os.path.abspath(inp)          # relative paths crash realpath (SURVEY 8c pitfall 1)
        rpt = os.path.abspath(rpt or inp[:-4] + ".rpt")
        out = os.path.abspath(out or inp[:-4] + ".out")
        err = self.lib.swmm_open(inp.encode(), rpt.encode(), out.encode())
        if err:
            self.lib.swmm_close()
            raise RuntimeError(f"swmm_open error {err} (see {rpt})")
        self.is_open = True

    def start(self, save: bool = True):
        err = self.lib.swmm_start(1 if save else 0)
        if err:
            raise RuntimeError(f"swmm_start error {err}")

    def step(self) -> float:
        t = C.c_double(0.0)
        err = self.lib.swmm_step(C.byref(t))
        if err:
            raise RuntimeError(f"swmm_step error {err}")
        return t.value

    def end(self):
        self.lib.swmm_end()

    def report(self):
        self.lib.swmm_report()

    def close(self):
        self.lib.swmm_close()
        self.is_open = False

    def mass_bal_err(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self.lib.swmm_getMassBalErr(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def get_value(self, prop: int, idx: int) -> float:
        return self.lib.swmm_getValue(prop, idx)

    # ---- flat views ------------------------------------------------------------------
    def network(self):
        d = self.hook.refhook_network()
        o = self.hook.refhook_options()
        return self.abi.Network.from_c(d.contents, o.contents)

    def field(self, fid) -> np.ndarray:
        if isinstance(fid, str):
            fid = self.abi.FIELD[fid]
        n = self.hook.refhook_field_len(fid)
        buf = np.zeros(max(n, 1), dtype=np.float64)
        rc = self.hook.refhook_get_field(fid, buf.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise KeyError(f"field {fid} not held by the engine")
        return buf[:n]

    def set_field(self, fid, arr):
        if isinstance(fid, str):
            fid = self.abi.FIELD[fid]
        buf = np.ascontiguousarray(arr, dtype=np.float64)
        self.hook.refhook_set_field(fid, buf.ctypes.data_as(C.POINTER(C.c_double)))

    def inflows(self) -> dict:
        """External FLOW hydrographs of the open model as keyword arguments for Solver.set_inflows."""
        d = self.hook.refhook_inflows().contents
        F = self.abi.FIELD
        n, npts = d.n_inflow_nodes, d.n_ts_pts
        nP = self.hook.refhook_field_len(F['SWB_NODE_NEW_QUAL']) // self.hook.refhook_field_len(F['SWB_NODE_NEW_DEPTH'])
        arr = lambda p, k, dt: np.ctypeslib.as_array(p, shape=(max(k, 1),))[:k].astype(dt).copy()
        return dict(node=arr(d.node, n, np.int32), ts_start=arr(d.ts_start, n + 1, np.int32),
                    ts_t=arr(d.ts_t, npts, np.float64), ts_q=arr(d.ts_q, npts, np.float64),
                    sfactor=arr(d.sfactor, n, np.float64), baseline=arr(d.baseline, n, np.float64),
                    concen=arr(d.concen, n * nP, np.float64) if nP else None,
                    start_day=d.start_day, start_secs=d.start_secs)

    def inflows_desc(self):
        """swb_inflow_desc of the open model from the shipped flattener (patterns, pollutant records, DWF):
        a ctypes struct whose arrays live in the hook library until the next call."""
        p = self.hook.refhook_inflows_full()
        assert p, "swb_flatten_inflows failed"
        return p.contents

    def controls_desc(self):
        """swb_controls_desc of the open model (rules, pump depths, orifice rates, timed outfall stages)."""
        p = self.hook.refhook_controls()
        assert p, "swb_flatten_controls: unsupported rule (named variable, expression or rain gage)"
        return p.contents

    def results(self, f: float, n_nodes: int, n_links: int, n_pollut: int):
        """The engine's own report records: node_getResults (node.c:497) / link_getResults
        (link.c:674) for every object at weighting factor f -> float32 [n][6 + P], [n][5 + P]."""
        L = self.lib
        L.node_getResults.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_float)]
        L.link_getResults.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_float)]
        L.node_getResults.restype = None
        L.link_getResults.restype = None
        nd = np.zeros((n_nodes, 6 + n_pollut), dtype=np.float32)
        ld = np.zeros((n_links, 5 + n_pollut), dtype=np.float32)
        buf = (C.c_float * (16 + n_pollut))()
        for j in range(n_nodes):
            L.node_getResults(j, float(f), buf)
            nd[j] = buf[:6 + n_pollut]
        for j in range(n_links):
            L.link_getResults(j, float(f), buf)
            ld[j] = buf[:5 + n_pollut]
        return nd, ld

    def statistics(self, n_nodes: int, n_links: int, n_pollut: int):
        """(node[planes, n_nodes], link[planes, n_links], MaxOutfallFlow) of the live engine in the plane
        order of include/swmm_b200.h (refhook_node_stats / refhook_link_stats)."""
        nd = np.zeros((self.abi.NODE_STAT["SWB_NS_PLANES"] + n_pollut, n_nodes))
        ld = np.zeros((self.abi.LINK_STAT["SWB_LS_PLANES"], n_links))
        P = C.POINTER(C.c_double)
        self.hook.refhook_node_stats.argtypes = [P]
        self.hook.refhook_link_stats.argtypes = [P]
        self.hook.refhook_max_outfall_flow.restype = C.c_double
        self.hook.refhook_node_stats(nd.ctypes.data_as(P))
        self.hook.refhook_link_stats(ld.ctypes.data_as(P))
        return nd, ld, self.hook.refhook_max_outfall_flow()

    def total_duration_s(self) -> float:
        return self.hook.refhook_total_duration() / 1000.0

    def last_iterations(self) -> int:
        """Return value of the most recent dynwave_execute call (Picard iterations of that step)."""
        return self.count.refcount_last_iterations()

    def last_critical(self):
        """(node, link) passed to the most recent stats_updateCriticalTimeCount (dynwave.c:827)."""
        n, l = C.c_int(-1), C.c_int(-1)
        self.count.refcount_last_critical(C.byref(n), C.byref(l))
        return n.value, l.value

    def seam_totals(self) -> dict:
        st, it = C.c_longlong(), C.c_longlong()
        t = (C.c_double * 4)()
        self.count.refcount_get(C.byref(st), C.byref(it), t)
        return {"steps": st.value, "iterations": it.value, "t_dynwave_execute": t[0],
                "t_get_routing_step": t[1], "t_qualrout_execute": t[2], "t_routing_execute": t[3]}

    def reset_seam_totals(self):
        self.count.refcount_reset()

    def routing_time_ms(self) -> float:
        return self.hook.refhook_new_routing_time()

    def non_converge_count(self) -> int:
        return self.hook.refhook_non_converge_count()

    # ---- geometry known-answer access --------------------------------------------------
    XS_FN = {"AofY": 0, "WofY": 1, "RofY": 2, "YofA": 3, "RofA": 4, "SofA": 5, "AofS": 6,
             "dSdA": 7, "Ycrit": 8}

    def xsect_set(self, xtype: int, geom, ucf: float = 1.0):
        g = (C.c_double * 4)(*[float(v) for v in geom])
        p = (C.c_double * 11)()
        ok = self.hook.refhook_xsect_set(xtype, g, ucf, p)
        return ok, np.array(list(p))

    def xsect_eval(self, fn: str, xtype: int, params, args) -> np.ndarray:
        p = np.ascontiguousarray(params, dtype=np.float64)
        a = np.ascontiguousarray(args, dtype=np.float64)
        out = np.zeros_like(a)
        self.hook.refhook_xsect_eval(self.XS_FN[fn], xtype,
                                     p.ctypes.data_as(C.POINTER(C.c_double)), a.size,
                                     a.ctypes.data_as(C.POINTER(C.c_double)),
                                     out.ctypes.data_as(C.POINTER(C.c_double)))
        return out
