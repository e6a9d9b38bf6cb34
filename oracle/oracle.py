"""TEST INFRASTRUCTURE: ctypes wrapper of the C restatement oracle (oracle/swmm_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libswmm_oracle.so")
sys.path.insert(0, os.path.dirname(HERE))
import swmm_b200  # noqa: E402,F401
from swmm_b200 import abi  # noqa: E402

_P_D = C.POINTER(C.c_double)


def build() -> str:
    src = os.path.join(HERE, "swmm_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    return LIB


class Oracle:
    def __init__(self, net: abi.Network):
        self.lib = C.CDLL(build())
        L = self.lib
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(abi.NetworkDesc), C.POINTER(abi.Options)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_inflows.argtypes = [C.c_void_p, C.POINTER(abi.InflowDesc), C.c_double, C.c_double]
        L.oracle_step.argtypes = [C.c_void_p, C.c_double]
        L.oracle_time.argtypes = [C.c_void_p]
        L.oracle_time.restype = C.c_double
        L.oracle_total_iterations.argtypes = [C.c_void_p]
        L.oracle_total_iterations.restype = C.c_longlong
        L.oracle_non_converged.argtypes = [C.c_void_p]
        L.oracle_non_converged.restype = C.c_longlong
        L.oracle_set_field.argtypes = [C.c_void_p, C.c_int, _P_D]
        L.oracle_get_field.argtypes = [C.c_void_p, C.c_int, _P_D]
        self.net = net
        self._d, self._o = net.to_c()
        self.h = L.oracle_create(C.byref(self._d), C.byref(self._o))
        if not self.h:
            raise NotImplementedError("network uses elements outside the oracle's coverage")
        self._keep = []

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def load_state(self, state: dict):
        for k, v in state.items():
            a = np.ascontiguousarray(v, dtype=np.float64)
            if a.size:
                self.lib.oracle_set_field(self.h, abi.FIELD[k], a.ctypes.data_as(_P_D))

    def set_inflows(self, node, ts_start, ts_t, ts_q, sfactor, baseline, concen=None, scale=1.0,
                    shift_days=0.0, start_day=0.0, start_secs=0.0):
        d = abi.InflowDesc()
        arrs = dict(node=np.ascontiguousarray(node, dtype=np.int32),
                    ts_start=np.ascontiguousarray(ts_start, dtype=np.int32),
                    ts_t=np.ascontiguousarray(ts_t, dtype=np.float64),
                    ts_q=np.ascontiguousarray(ts_q, dtype=np.float64),
                    sfactor=np.ascontiguousarray(sfactor, dtype=np.float64),
                    baseline=np.ascontiguousarray(baseline, dtype=np.float64))
        if concen is not None and self.net.n_pollut:
            arrs["concen"] = np.ascontiguousarray(concen, dtype=np.float64)
        d.n_inflow_nodes = arrs["node"].size
        d.n_ts_pts = arrs["ts_t"].size
        d.start_day, d.start_secs = float(start_day), float(start_secs)
        for name, base in abi.INFLOW_ARRAYS:
            if name in arrs:
                ct = C.c_int if base == "int" else C.c_double
                setattr(d, name, arrs[name].ctypes.data_as(C.POINTER(ct)))
        self._keep.append(arrs)
        self.lib.oracle_set_inflows(self.h, C.byref(d), float(scale), float(shift_days))

    def step(self, t_end: float) -> int:
        return self.lib.oracle_step(self.h, float(t_end))

    @property
    def time(self) -> float:
        return self.lib.oracle_time(self.h)

    def get_field(self, field: str) -> np.ndarray:
        fid = abi.FIELD[field]
        n = (self.net.n_nodes if abi.is_node_field(fid) else self.net.n_links) * \
            abi.field_width(fid, self.net.n_pollut)
        out = np.zeros(max(n, 1))
        self.lib.oracle_get_field(self.h, fid, out.ctypes.data_as(_P_D))
        return out[:n]
