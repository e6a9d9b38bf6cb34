"""ctypes mirror of include/swmm_b200.h.

The struct layouts and the field ids are parsed from the header itself so the Python side cannot
drift from the C-ABI.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "swmm_b200.h")

_CT = {"int": C.c_int, "double": C.c_double, "long long": C.c_longlong}


def _strip_comments(text: str) -> str:
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def _parse_struct(text: str, name: str):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S)
    if not m:
        raise RuntimeError(f"struct {name} not found in {HEADER}")
    fields = []
    for decl in m.group(1).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        const = decl.startswith("const ")
        if const:
            decl = decl[6:]
        for base in ("long long", "double", "int"):
            if decl.startswith(base + " "):
                rest = decl[len(base):]
                break
        else:
            raise RuntimeError(f"cannot parse declaration '{decl}' in {name}")
        for item in rest.split(","):
            item = item.strip()
            ptr = item.startswith("*")
            fname = item.lstrip("* ").strip()
            ctype = C.POINTER(_CT[base]) if ptr else _CT[base]
            fields.append((fname, ctype, base, ptr))
    return fields


def _parse_enum(text: str, name: str) -> dict:
    m = re.search(r"enum %s \{(.*?)\};" % name, text, flags=re.S)
    out, val = {}, -1
    for item in m.group(1).split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            k, v = item.split("=")
            v = v.strip()
            # an enumerator may be defined by another one (+ a constant)
            val = sum(out[t.strip()] if t.strip() in out else int(t) for t in v.split("+"))
            out[k.strip()] = val
        else:
            val += 1
            out[item] = val
    return out


_TEXT = _strip_comments(open(HEADER).read())
_DESC_FIELDS = _parse_struct(_TEXT, "swb_network_desc")
_OPT_FIELDS = _parse_struct(_TEXT, "swb_options")
_INFLOW_FIELDS = _parse_struct(_TEXT, "swb_inflow_desc")
_STATS_FIELDS = _parse_struct(_TEXT, "swb_member_stats")
_STEPIO_FIELDS = _parse_struct(_TEXT, "swb_step_io")
FIELD = _parse_enum(_TEXT, "swb_field")
FIELD_NAME = {v: k for k, v in FIELD.items()}
NODE_STAT = _parse_enum(_TEXT, "swb_node_stat")
LINK_STAT = _parse_enum(_TEXT, "swb_link_stat")
SYSTEM_STAT = _parse_enum(_TEXT, "swb_system_stat")


class NetworkDesc(C.Structure):
    _fields_ = [(n, t) for (n, t, _, _) in _DESC_FIELDS]


class Options(C.Structure):
    _fields_ = [(n, t) for (n, t, _, _) in _OPT_FIELDS]


class InflowDesc(C.Structure):
    _fields_ = [(n, t) for (n, t, _, _) in _INFLOW_FIELDS]


class ControlsDesc(C.Structure):
    _fields_ = [(n, t) for (n, t, _, _) in _parse_struct(_TEXT, "swb_controls_desc")]


class StepIO(C.Structure):
    _fields_ = [(n, t) for (n, t, _, _) in _STEPIO_FIELDS]


class MemberStats(C.Structure):
    _fields_ = [(n, t) for (n, t, _, _) in _STATS_FIELDS]


DESC_ARRAYS = [(n, base) for (n, _, base, ptr) in _DESC_FIELDS if ptr]
DESC_SCALARS = [n for (n, _, _, ptr) in _DESC_FIELDS if not ptr]
OPT_NAMES = [n for (n, _, _, _) in _OPT_FIELDS]
INFLOW_ARRAYS = [(n, base) for (n, _, base, ptr) in _INFLOW_FIELDS if ptr]

# length of every descriptor array in terms of the descriptor's counts
def desc_array_len(name: str, sc: dict) -> int:
    if name.startswith(("node_", "outfall_", "storage_")):
        return sc["n_nodes"]
    if name == "curve_start":
        return sc["n_curves"] + 1
    if name == "curve_type":
        return sc["n_curves"]
    if name in ("curve_x", "curve_y"):
        return sc["n_curve_pts"]
    if name == "shape_tbl_n":
        return sc["n_shape_tbls"]
    if name.startswith("shape_"):
        return sc["n_shape_tbls"] * sc["shape_tbl_len"]
    if name == "pollut_kdecay":
        return sc["n_pollut"]
    return sc["n_links"]


def is_node_field(fid: int) -> bool:
    return fid < FIELD["SWB_LINK_NEW_FLOW"]


def field_width(fid: int, n_pollut: int) -> int:
    if FIELD_NAME[fid] in ("SWB_NODE_NEW_QUAL", "SWB_NODE_OLD_QUAL", "SWB_LINK_NEW_QUAL",
                           "SWB_LINK_OLD_QUAL", "SWB_LINK_TOTAL_LOAD"):
        return n_pollut
    return 1


class Network:
    """A flat network held as numpy arrays (the Python twin of swb_network_desc + swb_options)."""

    def __init__(self, scalars: dict, arrays: dict, options: dict):
        self.scalars = dict(scalars)
        self.arrays = {}
        for name, base in DESC_ARRAYS:
            dt = np.int32 if base == "int" else np.float64
            n = desc_array_len(name, self.scalars)
            a = arrays.get(name)
            if a is None:
                fill = -1 if name in ("xs_table", "pump_curve", "weir_cd_curve", "outlet_curve",
                                      "storage_curve") else 0
                a = np.full(n, fill, dtype=dt)
                if name in ("cond_barrels", "link_direction"):
                    a[:] = 1
            a = np.ascontiguousarray(a, dtype=dt)
            if a.size != n:
                raise ValueError(f"{name}: expected {n} entries, got {a.size}")
            self.arrays[name] = a
        self.options = {k: options[k] for k in OPT_NAMES}

    @property
    def n_nodes(self): return self.scalars["n_nodes"]
    @property
    def n_links(self): return self.scalars["n_links"]
    @property
    def n_pollut(self): return self.scalars["n_pollut"]

    def true_conduit_mask(self) -> np.ndarray:
        a = self.arrays
        return (a["link_type"] == 0) & (a["xs_type"] != 0)

    def to_c(self):
        d = NetworkDesc()
        for k in DESC_SCALARS:
            setattr(d, k, int(self.scalars.get(k, 0)))
        for name, base in DESC_ARRAYS:
            ct = C.c_int if base == "int" else C.c_double
            setattr(d, name, self.arrays[name].ctypes.data_as(C.POINTER(ct)))
        o = Options()
        for k in OPT_NAMES:
            setattr(o, k, self.options[k])
        return d, o

    @classmethod
    def from_c(cls, d: NetworkDesc, o: Options) -> "Network":
        sc = {k: getattr(d, k) for k in DESC_SCALARS}
        arrays = {}
        for name, base in DESC_ARRAYS:
            n = desc_array_len(name, sc)
            p = getattr(d, name)
            dt = np.int32 if base == "int" else np.float64
            arrays[name] = np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy() if n else \
                np.zeros(0, dtype=dt)
        opts = {k: getattr(o, k) for k in OPT_NAMES}
        return cls(sc, arrays, opts)

    def save(self, path: str, **extra):
        np.savez_compressed(path, __scalars=np.array([self.scalars[k] for k in DESC_SCALARS]),
                            __options=np.array([float(self.options[k]) for k in OPT_NAMES]),
                            **self.arrays, **extra)

    @classmethod
    def load(cls, path: str):
        z = np.load(path)
        sc = {k: int(v) for k, v in zip(DESC_SCALARS, z["__scalars"])}
        ov = z["__options"]
        opts = {}
        for (n, _, base, _), v in zip(_OPT_FIELDS, ov):
            opts[n] = int(v) if base == "int" else float(v)
        arrays = {name: z[name] for name, _ in DESC_ARRAYS if name in z.files}
        net = cls(sc, arrays, opts)
        extra = {k: z[k] for k in z.files if k not in net.arrays and not k.startswith("__")}
        return net, extra
