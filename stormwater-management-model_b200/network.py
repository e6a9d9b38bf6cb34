"""Host-side network builder: flat networks for the synthetic BASELINE configs WITHOUT the engine.

bench.py and the ensemble driver use this on the GPU box, where only this repo exists.  It restates
the reference's SETUP arithmetic for the element types those configs use, so that a built network
equals the one the engine produces from the equivalent ``.inp`` bit for bit
(tests/test_baseline_size.py::test_network_builder_equals_engine_flattening, against oracle/_ref):
  xsect_setParams       xsect.c:216-330   (CIRCULAR, RECT_CLOSED)
  conduit_validate      link.c:992-1154   (slope, reversal, roughFactor, beta, qFull, qMax)
  conduit_getSlope      link.c:1258-1300
  dynwave_init          dynwave.c:137-160 (crown elevations)
  toposort / layout     toposort.c:76-92, flowrout.c:274-333 (Node.degree and its sign)
  external inflows      inflow.c:41-134, table.c:113-204 (time series breakpoints as DateTime)
Scalars use math.pow / math.sqrt (libm, like the reference), never vectorised numpy pow.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import abi, scenarios

PI = 3.141592654          # consts.h
PHI = 1.486
GRAVITY = 32.2
MIN_DELTA_Z = 0.001       # link.c:63
RECT_ALFMAX = 0.97

XS_CIRCULAR, XS_RECT_CLOSED = 1, 3

DEFAULT_OPTIONS = dict(
    surcharge_method=0, inert_damping=1, normal_flow_ltd=2, allow_ponding=0, max_trials=8,
    force_main_eqn=0, unit_system=0, ignore_quality=0, head_tol=0.005, min_surf_area=12.566,
    courant_factor=0.75, min_route_step=0.5, route_step=5.0, ucf_length=1.0, ucf_volume=1.0,
    ucf_flow=1.0, evap_rate=0.0, hydcon_factor=1.0)


def xsect_params(xs_type: int, g1: float, g2: float = 0.0) -> dict:
    """xsect_setParams for the two closed shapes of configs 1/2/4/5 (US units, ucf = 1)."""
    if xs_type == XS_CIRCULAR:
        y = g1 / 1.0
        a = PI / 4.0 * y * y
        r = 0.2500 * y
        s = a * math.pow(r, 2. / 3.)
        return dict(yfull=y, wmax=y, afull=a, rfull=r, sfull=s, smax=1.08 * s, ywmax=0.5 * y)
    if xs_type == XS_RECT_CLOSED:
        y = g1 / 1.0
        w = g2 / 1.0
        a = y * w
        r = a / (2.0 * (y + w))
        s = a * math.pow(r, 2. / 3.)
        amax = RECT_ALFMAX * a
        # rect_closed_getRofA (xsect.c:1793-1803) at alpha == RECT_ALFMAX exactly: no top term
        p = w + 2. * amax / w
        if amax / a > RECT_ALFMAX:
            p += (amax / a - RECT_ALFMAX) / (1.0 - RECT_ALFMAX) * w
        smax = amax * math.pow(amax / p, 2. / 3.)
        return dict(yfull=y, wmax=w, afull=a, rfull=r, sfull=s, smax=smax, ywmax=y)
    raise NotImplementedError(f"xsect type {xs_type}: use the engine flattening (seam/flatten.c)")


def encode_date(year: int, month: int, day: int) -> float:
    """datetime_encodeDate (datetime.c:84-105): days since 12/31/1899."""
    days_per_month = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    leap = (year % 4 == 0) and ((year % 100 != 0) or (year % 400 == 0))
    if leap:
        days_per_month[1] = 29
    for i in range(month - 1):
        day += days_per_month[i]
    i = year - 1
    return float(i * 365 + i // 4 - i // 100 + i // 400 + day - 693594)


def encode_time(h: int, m: int, s: int) -> float:
    """datetime_encodeTime (datetime.c:142-153)."""
    return float(h * 3600 + m * 60 + s) / 86400.


@dataclass
class BuiltCase:
    net: abi.Network
    state0: dict            # field name -> single-member array (all zero for a dry start)
    inflows: dict           # keyword arguments of Solver.set_inflows
    t_end: float            # seconds


class _Builder:
    def __init__(self, options: dict, n_pollut: int = 0, kdecay=()):
        self.opt = dict(DEFAULT_OPTIONS)
        self.opt.update(options)
        self.nodes = []      # (type, invert, full_depth, outfall_type)
        self.links = []      # dicts
        self.n_pollut = n_pollut
        self.kdecay = list(kdecay)

    def junction(self, invert: float, max_depth: float) -> int:
        self.nodes.append((0, invert, max_depth, 0))
        return len(self.nodes) - 1

    def outfall_free(self, invert: float) -> int:
        self.nodes.append((1, invert, 0.0, 0))
        return len(self.nodes) - 1

    def conduit(self, n1: int, n2: int, length: float, n_manning: float, xs_type: int, g1: float,
                g2: float = 0.0):
        self.links.append(dict(n1=n1, n2=n2, length=length, n=n_manning, xs=xs_type, g1=g1, g2=g2))

    def finish(self) -> abi.Network:
        nN, nL = len(self.nodes), len(self.links)
        A = {}
        f64 = lambda: np.zeros(nL)
        i32 = lambda: np.zeros(nL, dtype=np.int32)
        node_type = np.array([t for t, *_ in self.nodes], dtype=np.int32)
        invert = np.array([z for _, z, *_ in self.nodes])
        full_depth = np.array([d for _, _, d, _ in self.nodes])
        n1 = np.array([l["n1"] for l in self.links], dtype=np.int32)
        n2 = np.array([l["n2"] for l in self.links], dtype=np.int32)
        direction = np.ones(nL, dtype=np.int32)
        keys = ("yfull", "wmax", "ywmax", "afull", "rfull", "sfull", "smax")
        xs = {k: f64() for k in keys}
        xs_type = i32()
        length, slope_a, rough, beta, qmax, qfull = f64(), f64(), f64(), f64(), f64(), f64()
        cache = {}
        for j, l in enumerate(self.links):
            key = (l["xs"], l["g1"], l["g2"])
            if key not in cache:
                cache[key] = xsect_params(*key)
            p = cache[key]
            xs_type[j] = l["xs"]
            for k in keys:
                xs[k][j] = p[k]
            # conduit_getSlope (link.c:1258-1300), offsets are zero in these configs
            L = l["length"]
            e1, e2 = invert[n1[j]], invert[n2[j]]
            delta = abs(e1 - e2)
            if delta < MIN_DELTA_Z:
                delta = MIN_DELTA_Z
            if delta >= L:
                sl = delta / L
            else:
                sl = delta / math.sqrt(L * L - delta * delta)
            if e1 < e2:
                sl = -sl
            # conduit_reverse for adverse slopes under DW (link.c:1082-1087, 1158-1191)
            if sl < 0.0:
                n1[j], n2[j] = n2[j], n1[j]
                sl = -sl
                direction[j] = -1
            length[j] = L
            slope_a[j] = sl
            rough[j] = GRAVITY * ((l["n"] / PHI) * (l["n"] / PHI))
            beta[j] = PHI * math.sqrt(abs(sl)) / l["n"]
            qfull[j] = p["sfull"] * beta[j]
            qmax[j] = p["smax"] * beta[j]
        # link_validate (link.c:440-463): end nodes are at least as deep as the link crown
        for j in range(nL):
            for nd in (n1[j], n2[j]):
                full_depth[nd] = max(full_depth[nd], 0.0 + xs["yfull"][j])
        # crown elevations (dynwave.c:137-153)
        crown = invert.copy()
        for j in range(nL):
            for nd in (n1[j], n2[j]):
                z = invert[nd] + 0.0 + xs["yfull"][j]
                crown[nd] = max(crown[nd], z)
        # Node.degree (toposort.c:76-92) and its sign (flowrout.c:283-331)
        degree = np.zeros(nN, dtype=np.int32)
        inflow_links = np.zeros(nN, dtype=np.int32)
        for j in range(nL):
            n = n1[j] if direction[j] > 0 else n2[j]
            if node_type[n] == 1:
                n = n1[j] if direction[j] < 0 else n2[j]
            degree[n] += 1
        for j in range(nL):
            i = n1[j]
            if node_type[i] != 1:
                i = n2[j]
            inflow_links[i] += 1
        degree = np.where(inflow_links == 0, -degree, degree).astype(np.int32)

        A.update(node_type=node_type, node_degree=degree, node_invert=invert,
                 node_full_depth=full_depth, node_crown_elev=crown,
                 link_type=i32(), link_node1=n1, link_node2=n2, link_direction=direction,
                 xs_type=xs_type, xs_yfull=xs["yfull"], xs_wmax=xs["wmax"], xs_ywmax=xs["ywmax"],
                 xs_afull=xs["afull"], xs_rfull=xs["rfull"], xs_sfull=xs["sfull"], xs_smax=xs["smax"],
                 cond_barrels=np.ones(nL, dtype=np.int32), cond_length=length.copy(),
                 cond_user_length=length.copy(), cond_mod_length=length.copy(),
                 cond_rough_factor=rough, cond_slope=slope_a, cond_beta=beta, cond_q_max=qmax,
                 link_q_full=qfull, pollut_kdecay=np.array(self.kdecay, dtype=np.float64))
        sc = dict(n_nodes=nN, n_links=nL, n_pollut=self.n_pollut, n_curves=0, n_curve_pts=0,
                  n_shape_tbls=0, shape_tbl_len=51, reserved0=0)
        return abi.Network(sc, A, self.opt)


def _dry_state(net: abi.Network, lib_path=None) -> dict:
    """State after swmm_start for a dry network (flowrout.c:337-510): every conduit starts at the
    FUDGE depth, with the matching area and volume; nodes are empty."""
    from . import solver
    A = net.arrays
    nL = net.n_links
    fudge = 0.0001
    a1 = np.zeros(nL)
    keys = ("xs_yfull", "xs_wmax", "xs_ywmax", "xs_afull", "xs_rfull", "xs_sfull", "xs_smax",
            "xs_ybot", "xs_abot", "xs_sbot", "xs_rbot")
    cache = {}
    for j in range(nL):
        key = (int(A["xs_type"][j]),) + tuple(float(A[k][j]) for k in keys)
        if key not in cache:
            cache[key] = float(solver.xsect_eval("AofY", key[0], key[1:], [fudge], lib_path=lib_path)[0])
        a1[j] = cache[key]
    vol = a1 * A["cond_length"] * A["cond_barrels"]
    stage = np.where(A["node_type"] == 1, A["node_invert"], 0.0)
    return {"SWB_LINK_SETTING": np.ones(nL), "SWB_LINK_TARGET_SETTING": np.ones(nL),
            "SWB_LINK_NEW_DEPTH": np.full(nL, fudge), "SWB_LINK_NEW_VOLUME": vol,
            "SWB_LINK_OLD_VOLUME": vol.copy(), "SWB_COND_A1": a1, "SWB_COND_A2": a1.copy(),
            "SWB_NODE_OUTFALL_STAGE": stage}


def _options_from_spec(spec) -> dict:
    return dict(surcharge_method=1 if spec.surcharge.upper() == "SLOT" else 0,
                route_step=float(spec.route_step), courant_factor=float(spec.variable_step))


def build_grid(spec: scenarios.GridSpec | None = None, lib_path=None) -> BuiltCase:
    """Config 2 / 5 network, identical to parsing scenarios.c2_grid_inp(spec) with the engine."""
    s = spec or scenarios.GridSpec()
    nP = 2 if s.pollutants else 0
    b = _Builder(_options_from_spec(s), n_pollut=nP, kdecay=[0.5 / 86400., 0.0] if nP else [])
    ids = {}
    for i in range(s.nx):
        for j in range(s.ny):
            ids[(i, j)] = b.junction(float(f"{s.elev(i, j):.4f}"), s.max_depth)
    out = b.outfall_free(float(f"{s.elev(s.nx - 1, s.ny - 1) - 0.4:.4f}"))
    k = 0
    for i in range(s.nx):
        for j in range(s.ny):
            for (di, dj) in ((1, 0), (0, 1)):
                ii, jj = i + di, j + dj
                if ii >= s.nx or jj >= s.ny:
                    continue
                if k % 2 == 0:
                    b.conduit(ids[(i, j)], ids[(ii, jj)], s.length, s.roughness, XS_CIRCULAR, s.size)
                else:
                    b.conduit(ids[(i, j)], ids[(ii, jj)], s.length, s.roughness, XS_RECT_CLOSED, s.size, s.size)
                k += 1
    b.conduit(ids[(s.nx - 1, s.ny - 1)], out, s.length, s.roughness, XS_CIRCULAR, 2.0 * s.size)
    net = b.finish()
    # external inflows: node order = node index order (routing.c:445), series HYDRO shared
    start = encode_date(2020, 1, 1)
    pts = s.hydrograph()
    ts_t = [start + encode_time(*_hms_int(t)) for t, _ in pts]
    ts_q = [v for _, v in pts]
    nodes, sf = [], []
    for (i, j, sc) in s.inflow_nodes():
        nodes.append(ids[(i, j)])
        sf.append(s.sfactor(sc))
    order = np.argsort(nodes, kind="stable")
    nodes = np.array(nodes, dtype=np.int32)[order]
    sf = np.array(sf)[order]
    n = len(nodes)
    m = len(ts_t)
    inflows = dict(node=nodes, ts_start=np.arange(n + 1, dtype=np.int32) * m,
                   ts_t=np.tile(np.array(ts_t), n), ts_q=np.tile(np.array(ts_q), n),
                   sfactor=sf, baseline=np.zeros(n),
                   concen=np.tile(np.array([100.0, 50.0]) * CONCEN_MGL, n) if nP else None,
                   start_day=start, start_secs=0.0)
    return BuiltCase(net, _dry_state(net, lib_path), inflows, s.hours * 3600.0)


# CONCEN inflows keep the user's concentration units (cFactor = 1, inflow.c:84-118)
CONCEN_MGL = 1.0


def _hms_int(hours: float):
    s = int(round(hours * 3600))
    return s // 3600, (s % 3600) // 60, s % 60
