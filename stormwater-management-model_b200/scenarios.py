"""Deterministic generators for the BASELINE.json configs.

Each generator returns the text of a SWMM 5.2 ``.inp`` (formats per SURVEY.md Appendix E, parsers
cited there: node.c:606-809, link.c:933-988,162-311, inflow.c:41-134, table.c:67-204).  The same
scenario can also be built directly as flat arrays by ``network.build_*`` (no engine needed on the
GPU box); tests check the two agree bit for bit.

Configs:
  c1  100-junction dendritic tree, circular pipes, 6 h triangular storm
  c2  n x n looped grid (n=100 -> 10k nodes / 19 801 conduits), circular + rect_closed,
      surcharge, 2 pollutants (TSS decays, DYE conservative); SLOT (strict gate) or EXTRAN
  c3  mixed-element network (storage, pumps, orifices, weirs, outlet, outfalls, rules)
  c4  members of c2 = (inflow scale, time shift) pairs, rng 2024
  c5  c2 at 1000 x 500 (~1M links), no pollutants
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field

import numpy as np


def _hms(hours: float) -> str:
    s = int(round(hours * 3600))
    return f"{s // 3600}:{(s % 3600) // 60:02d}:{s % 60:02d}"


def _options(*, hours: float, route_step: float, variable_step: float, surcharge: str,
             report_step_s: int, threads: int, extra: dict | None = None) -> list[str]:
    days = int(hours // 24)
    rem = hours - 24 * days
    end_date = f"01/{1 + days:02d}/2020"
    o = {
        "FLOW_UNITS": "CFS",
        "FLOW_ROUTING": "DYNWAVE",
        "START_DATE": "01/01/2020",
        "START_TIME": "00:00:00",
        "REPORT_START_DATE": "01/01/2020",
        "REPORT_START_TIME": "00:00:00",
        "END_DATE": end_date,
        "END_TIME": _hms(rem) if rem > 0 else "00:00:00",
        "REPORT_STEP": _hms(report_step_s / 3600.0),
        "ROUTING_STEP": f"{route_step:g}",
        "VARIABLE_STEP": f"{variable_step:g}",
        "SURCHARGE_METHOD": surcharge,
        "MAX_TRIALS": "8",
        "HEAD_TOLERANCE": "0.005",
        "THREADS": str(threads),
    }
    if hours >= 24 and rem == 0:
        o["END_TIME"] = "00:00:00"
    if extra:
        o.update(extra)
    return ["[OPTIONS]"] + [f"{k} {v}" for k, v in o.items()]


# --------------------------------------------------------------------------------------------
# config 1: dendritic tree
# --------------------------------------------------------------------------------------------
@dataclass
class TreeSpec:
    n_junctions: int = 100
    length: float = 400.0
    roughness: float = 0.013
    drop: float = 2.0            # invert drop per conduit, ft
    max_depth: float = 8.0
    peak_cfs: float = 0.6
    hours: float = 6.0
    route_step: float = 5.0
    variable_step: float = 0.75
    surcharge: str = "EXTRAN"
    report_step_s: int = 300
    threads: int = 1

    def level(self, i: int) -> int:           # heap index i >= 1
        return int(math.floor(math.log2(i)))

    def diameter(self, i: int) -> float:
        top = self.level(self.n_junctions)
        return max(1.0, 4.0 - 3.0 * self.level(i) / max(top, 1))

    def leaves(self) -> list[int]:
        n = self.n_junctions
        return [i for i in range(1, n + 1) if 2 * i > n]


def c1_tree_inp(spec: TreeSpec | None = None) -> str:
    s = spec or TreeSpec()
    n = s.n_junctions
    L = _options(hours=s.hours, route_step=s.route_step, variable_step=s.variable_step,
                 surcharge=s.surcharge, report_step_s=s.report_step_s, threads=s.threads)
    L += ["[JUNCTIONS]"]
    for i in range(1, n + 1):
        elev = 10.0 + s.drop * (s.level(i) + 1)
        L.append(f"J{i} {elev:.4f} {s.max_depth:g} 0 0 0")
    L += ["[OUTFALLS]", "OUT 10 FREE NO"]
    L += ["[CONDUITS]"]
    for i in range(1, n + 1):
        dn = f"J{i // 2}" if i > 1 else "OUT"
        L.append(f"C{i} J{i} {dn} {s.length:g} {s.roughness:g} 0 0 0 0")
    L += ["[XSECTIONS]"]
    for i in range(1, n + 1):
        L.append(f"C{i} CIRCULAR {s.diameter(i):.6g} 0 0 0 1")
    L += ["[INFLOWS]"]
    for k, i in enumerate(s.leaves()):
        scale = 0.5 + 0.5 * ((k * 7919) % 100) / 100.0
        L.append(f"J{i} FLOW STORM FLOW 1.0 {scale:.4f}")
    L += ["[TIMESERIES]", "STORM 0:00 0", f"STORM 2:00 {s.peak_cfs:g}", "STORM 4:00 0",
          "STORM 6:00 0"]
    L += ["[REPORT]", "NODES ALL", "LINKS ALL"]
    return "\n".join(L) + "\n"


# --------------------------------------------------------------------------------------------
# config 2 / 5: looped urban grid
# --------------------------------------------------------------------------------------------
@dataclass
class GridSpec:
    nx: int = 100
    ny: int = 100
    length: float = 300.0
    roughness: float = 0.013
    size: float = 1.5            # circular diameter and rect_closed side, ft
    max_depth: float = 6.0
    hours: float = 2.0
    route_step: float = 5.0
    variable_step: float = 0.75
    surcharge: str = "SLOT"
    pollutants: bool = True
    report_step_s: int = 300
    threads: int = 1
    seed: int = 1
    inflow_scale: float = 1.0    # member scale (config 4)
    inflow_shift_h: float = 0.0  # member time shift, hours (config 4)
    # hydrograph breakpoints (hours, multiplier)
    hydro: tuple = ((0.0, 0.0), (1.5, 1.0), (3.0, 4.0), (4.5, 1.0), (6.0, 0.0))

    def node_id(self, i: int, j: int) -> str:
        return f"N{i}_{j}"

    def elev(self, i: int, j: int) -> float:
        return 100.0 + 0.4 * ((self.nx - 1 - i) + (self.ny - 1 - j))

    def inflow_nodes(self) -> list[tuple[int, int, float]]:
        """(i, j, scale) for every 2nd node, scale ~ U(0.2, 0.5), random.seed(seed)."""
        rng = random.Random(self.seed)
        out = []
        k = 0
        for i in range(self.nx):
            for j in range(self.ny):
                if k % 2 == 0:
                    out.append((i, j, rng.uniform(0.2, 0.5)))
                k += 1
        return out

    def sfactor(self, sc: float) -> float:
        """Scale factor written for an inflow node: the base model's 6-decimal value, times the
        member scale in ONE double multiplication -- exactly what the device forms from
        (sfactor, member_scale), so a perturbed .inp and an ensemble member see the same bits."""
        base = float(f"{sc:.6f}")
        return base if self.inflow_scale == 1.0 else base * float(self.inflow_scale)

    def sfactor_text(self, sc: float) -> str:
        return f"{sc:.6f}" if self.inflow_scale == 1.0 else repr(self.sfactor(sc))

    def hydrograph(self) -> list[tuple[float, float]]:
        """Breakpoints after the member time shift (hours >= 0, value)."""
        pts = [(t + self.inflow_shift_h, v) for t, v in self.hydro]
        return pts


def c2_grid_inp(spec: GridSpec | None = None) -> str:
    s = spec or GridSpec()
    L = _options(hours=s.hours, route_step=s.route_step, variable_step=s.variable_step,
                 surcharge=s.surcharge, report_step_s=s.report_step_s, threads=s.threads)
    L += ["[JUNCTIONS]"]
    for i in range(s.nx):
        for j in range(s.ny):
            L.append(f"{s.node_id(i, j)} {s.elev(i, j):.4f} {s.max_depth:g} 0 0 0")
    out_elev = s.elev(s.nx - 1, s.ny - 1) - 0.4
    L += ["[OUTFALLS]", f"OUT {out_elev:.4f} FREE NO"]
    cond, xs = [], []
    k = 0
    for i in range(s.nx):
        for j in range(s.ny):
            for (di, dj) in ((1, 0), (0, 1)):
                ii, jj = i + di, j + dj
                if ii >= s.nx or jj >= s.ny:
                    continue
                cid = f"C{k}"
                cond.append(f"{cid} {s.node_id(i, j)} {s.node_id(ii, jj)} {s.length:g} "
                            f"{s.roughness:g} 0 0 0 0")
                if k % 2 == 0:
                    xs.append(f"{cid} CIRCULAR {s.size:g} 0 0 0 1")
                else:
                    xs.append(f"{cid} RECT_CLOSED {s.size:g} {s.size:g} 0 0 1")
                k += 1
    cond.append(f"COUT {s.node_id(s.nx - 1, s.ny - 1)} OUT {s.length:g} {s.roughness:g} 0 0 0 0")
    xs.append(f"COUT CIRCULAR {2.0 * s.size:g} 0 0 0 1")
    L += ["[CONDUITS]"] + cond + ["[XSECTIONS]"] + xs
    if s.pollutants:
        L += ["[POLLUTANTS]", "TSS MG/L 0 0 0 0.5", "DYE MG/L 0 0 0 0"]
    L += ["[INFLOWS]"]
    for (i, j, sc) in s.inflow_nodes():
        nid = s.node_id(i, j)
        L.append(f"{nid} FLOW HYDRO FLOW 1.0 {s.sfactor_text(sc)}")
        if s.pollutants:
            L.append(f"{nid} TSS CTSS CONCEN 1.0 1.0")
            L.append(f"{nid} DYE CDYE CONCEN 1.0 1.0")
    L += ["[TIMESERIES]"]
    for (t, v) in s.hydrograph():
        L.append(f"HYDRO {_hms(t)} {v:g}")
    if s.pollutants:
        L += ["CTSS 0:00 100", "CTSS 48:00 100", "CDYE 0:00 50", "CDYE 48:00 50"]
    L += ["[REPORT]", "NODES ALL", "LINKS ALL"]
    return "\n".join(L) + "\n"


def c4_members(n_members: int = 4096, seed: int = 2024):
    """(scale, shift_hours) per member: LogNormal(0, 0.3) scale, U(-30, 30) min shift (8d).

    Shifts are made non-negative by a common +30 min offset so every hydrograph starts at t >= 0.
    """
    rng = np.random.default_rng(seed)
    scale = rng.lognormal(0.0, 0.3, n_members)
    shift_min = rng.uniform(-30.0, 30.0, n_members)
    # quantise to what an .inp can express exactly (6 decimals / whole seconds)
    scale = np.round(scale, 6)
    shift_h = np.round((shift_min + 30.0) * 60.0) / 3600.0
    return scale, shift_h


# --------------------------------------------------------------------------------------------
# config 3: mixed elements (SURVEY.md Appendix E seed model, verified against the reference)
# --------------------------------------------------------------------------------------------
C3_MIXED_INP = """[OPTIONS]
FLOW_UNITS CFS
FLOW_ROUTING DYNWAVE
START_DATE 01/01/2020
START_TIME 00:00:00
END_DATE 01/02/2020
END_TIME 00:00:00
REPORT_STEP 00:15:00
ROUTING_STEP 5
VARIABLE_STEP 0.75
ALLOW_PONDING YES
THREADS 1
[JUNCTIONS]
J1 110 8 0 0 500
J2 108 8 0 0 0
J3 100 8 0 0 0
J4 99 8 0 0 0
J5 98 8 0 0 0
WW 90 15 0 0 0
[OUTFALLS]
O1 97 FREE NO
O2 96 FIXED 97.5 YES
O3 96 TIDAL TIDE NO
[STORAGE]
S1 104 12 2 FUNCTIONAL 1000 0 500 0 0
S2 102 10 1 TABULAR SCURVE 0 0
[CONDUITS]
C1 J1 J2 400 0.013 0 0 0 0
C2 J2 S1 400 0.013 0 1 0 0
C3 J3 J4 400 0.013 0 0 0 0
C4 J4 J5 300 0.015 0 0.5 0 0
C5 J5 O1 300 0.013 0 0 0 0
C6 J4 O2 300 0.013 0 0 0 0
C7 S2 O3 300 0.013 0 0 0 0
C8 J3 WW 200 0.013 0 0 0 0
[PUMPS]
P1 WW J4 PC3 ON 4 1
[ORIFICES]
OR1 S1 J3 SIDE 0 0.65 NO 0
OR2 S1 S2 BOTTOM 0 0.6 NO 0.05
[WEIRS]
W1 S1 S2 TRANSVERSE 8 3.33 NO 0 0 YES
W2 S2 J3 V-NOTCH 5 2.5 NO 0 0 YES
[OUTLETS]
OL1 S2 J3 0 FUNCTIONAL/DEPTH 2.0 0.5 NO
[XSECTIONS]
C1 CIRCULAR 2 0 0 0 1
C2 RECT_CLOSED 2 3 0 0 1
C3 CIRCULAR 3 0 0 0 1
C4 TRAPEZOIDAL 4 3 1 1 1
C5 EGG 3 0 0 0 1
C6 CIRCULAR 2 0 0 0 2
C7 CIRCULAR 2.5 0 0 0 1
C8 CIRCULAR 2 0 0 0 1
OR1 CIRCULAR 1 0 0 0
OR2 RECT_CLOSED 1 1 0 0
W1 RECT_OPEN 3 10 0 0
W2 TRIANGULAR 4 6 0 0
[LOSSES]
C3 0.5 0.5 0 NO 0
[CONTROLS]
RULE R1
IF NODE S1 DEPTH > 6
THEN ORIFICE OR1 SETTING = 1.0
ELSE ORIFICE OR1 SETTING = 0.3
PRIORITY 1
RULE R2
IF SIMULATION TIME > 12
THEN PUMP P1 STATUS = OFF
[POLLUTANTS]
TSS MG/L 0 0 0 0.5
DYE MG/L 0 0 0 0
[INFLOWS]
J1 FLOW TS1 FLOW 1.0 1.0
J1 TSS TSC CONCEN 1.0 1.0
J2 FLOW TS1 FLOW 1.0 0.5 0.2
J2 DYE TSC CONCEN 1.0 0.5
[DWF]
J3 FLOW 0.3 DAILYP
[PATTERNS]
DAILYP HOURLY 0.5 0.5 0.5 0.5 0.6 0.8 1.2 1.5 1.4 1.2 1.1 1.0 1.0 1.0 1.0 1.1 1.2 1.4 1.5 1.3 1.0 0.8 0.6 0.5
[CURVES]
SCURVE STORAGE 0 800 5 1500 10 3000
PC3 PUMP3 0 8 5 6 10 3 15 0
TIDE TIDAL 0 96.5 6 98 12 96.5 18 98 24 96.5
[TIMESERIES]
TS1 0:00 0
TS1 2:00 10
TS1 4:00 30
TS1 8:00 5
TS1 24:00 1
TSC 0:00 100
TSC 24:00 100
[REPORT]
NODES ALL
LINKS ALL
CONTROLS YES
"""


def c3_mixed_inp() -> str:
    return C3_MIXED_INP


# The same model with a rule base that exercises the rest of controls.c: conflicting rules resolved by
# priority, CLOCKTIME windows, a right-hand-side variable, OR clauses, TIMEOPEN, conduit status, and the
# three modulated settings (CURVE of the control value, TIMESERIES, PID), optionally with RULE_STEP.
C3_RULES = """[CONTROLS]
RULE R1
IF NODE S1 DEPTH > 6
THEN ORIFICE OR1 SETTING = 1.0
ELSE ORIFICE OR1 SETTING = 0.3
PRIORITY 1
RULE R1B
IF SIMULATION CLOCKTIME >= 06:00:00
AND SIMULATION CLOCKTIME < 07:30:00
THEN ORIFICE OR1 SETTING = 0.6
PRIORITY 3
RULE R2
IF SIMULATION TIME > 12
THEN PUMP P1 STATUS = OFF
RULE R3
IF NODE S2 DEPTH >= 0
THEN WEIR W2 SETTING = CURVE WCURVE
RULE R4
IF NODE S1 DEPTH <> 5
THEN ORIFICE OR2 SETTING = PID -0.5 10 0
RULE R5
IF SIMULATION TIME >= 0
THEN WEIR W1 SETTING = TIMESERIES WTS
RULE R6
IF NODE J3 DEPTH > NODE J4 DEPTH
OR PUMP P1 TIMEOPEN > 2:00
THEN CONDUIT C8 STATUS = CLOSED
ELSE CONDUIT C8 STATUS = OPEN
RULE R7
IF SIMULATION MONTH = 1
AND SIMULATION DAY = 4
AND LINK C3 VELOCITY > 3
THEN OUTLET OL1 SETTING = 0.5
ELSE OUTLET OL1 SETTING = 1.0
"""


def c3_rules_inp(rule_step: str | None = None, pid: bool = True) -> str:
    """pid=False drops rule R4: the PID controller zeroes updates below 1e-4 (controls.c:1141), a
    discontinuity that turns a last-bit difference in a depth into a 1e-4 difference in a setting."""
    a = C3_MIXED_INP.index("[CONTROLS]")
    b = C3_MIXED_INP.index("[POLLUTANTS]")
    rules = C3_RULES
    if not pid:
        r4 = rules.index("RULE R4")
        rules = rules[:r4] + rules[rules.index("RULE R5"):]
    txt = C3_MIXED_INP[:a] + rules + C3_MIXED_INP[b:]
    txt = txt.replace("[CURVES]\n", "[CURVES]\nWCURVE CONTROL 0 1.0 2 0.8 5 0.4\n")
    txt = txt.replace("[TIMESERIES]\n", "[TIMESERIES]\nWTS 0:00 1.0\nWTS 6:00 0.5\nWTS 12:00 1.0\n")
    if rule_step:
        txt = txt.replace("THREADS 1\n", f"THREADS 1\nRULE_STEP {rule_step}\n", 1)
    return txt


# A second mixed-element model that exercises what C3_MIXED_INP does not: pump types 1, 2, 4 and
# ideal, SIDEFLOW / TRAPEZOIDAL weirs, flap-gated and slowly closing orifices, a TABULAR/HEAD outlet,
# CYLINDRICAL / CONICAL / PYRAMIDAL storage with evaporation, open channels (trapezoid, rect_open,
# parabolic, power, triangular), filled circular, horseshoe, arch, gothic, catenary, semi-elliptical,
# basket-handle, semi-circular, modified basket-handle, rect_round, rect_triangular and horizontal /
# vertical ellipse sections, an irregular transect, a custom shape, conduit seepage + evaporation,
# local losses, a flow limit, NORMAL and TIMESERIES outfalls, SI-free CFS units.
C3B_SHAPES_INP = """[OPTIONS]
FLOW_UNITS CFS
FLOW_ROUTING DYNWAVE
START_DATE 01/01/2020
START_TIME 00:00:00
END_DATE 01/01/2020
END_TIME 12:00:00
REPORT_STEP 00:10:00
ROUTING_STEP 5
VARIABLE_STEP 0.75
ALLOW_PONDING YES
SURCHARGE_METHOD SLOT
THREADS 1
[EVAPORATION]
CONSTANT 0.5
[JUNCTIONS]
A1 120 10 0 0 200
A2 118 10 0 0 0
A3 116 10 0 0 0
A4 114 10 0 0 0
A5 112 10 0 0 0
A6 110 10 0 0 0
A7 108 10 0 0 0
A8 106 10 0 0 0
A9 104 10 0 0 0
B1 119 10 0 0 0
B2 117 10 0 0 0
B3 115 10 0 0 0
B4 113 10 0 0 0
B5 111 10 0 0 0
B6 109 10 0 0 0
B7 107 10 0 0 0
B8 105 10 0 0 0
W1 95 12 0 0 0
W2 95 12 0 0 0
W3 95 12 0 0 0
W4 95 12 0 0 0
[OUTFALLS]
OA 100 NORMAL NO
OB 100 TIMESERIES STG YES
OC 99 FREE NO
[STORAGE]
T1 103 10 1 CYLINDRICAL 30 20 0 0 0.5
T2 102 10 0.5 CONICAL 20 10 0.5 0 1.0
T3 101 10 0 PYRAMIDAL 40 20 1 0 0
[CONDUITS]
CA1 A1 A2 300 0.013 0 0 0 0
CA2 A2 A3 300 0.013 0 0 0 0
CA3 A3 A4 300 0.013 0 0 0 0
CA4 A4 A5 300 0.013 0 0 0 0
CA5 A5 A6 300 0.013 0 0 0 0
CA6 A6 A7 300 0.013 0 0 0 0
CA7 A7 A8 300 0.013 0 0 0 0
CA8 A8 A9 300 0.013 0 0 0 0
CA9 A9 T1 300 0.013 0 1 0 0
CB1 B1 B2 300 0.02 0 0 0 0
CB2 B2 B3 300 0.02 0 0 0 0
CB3 B3 B4 300 0.02 0 0 0 0
CB4 B4 B5 300 0.02 0 0 0 0
CB5 B5 B6 300 0.02 0 0 0 0
CB6 B6 B7 300 0.02 0 0 0 3
CB7 B7 B8 300 0.03 0 0 0 0
CB8 B8 T2 300 0.02 0 0.5 0 0
CX1 A5 B5 200 0.013 0.5 0 0 0
CX2 T3 OC 300 0.013 0 0 0 0
CO1 W1 OA 200 0.013 0 0 0 0
CO2 W3 OB 200 0.013 0 0 0 0
CP4 W2 W4 200 0.013 0 0 0 0
[PUMPS]
P1 T1 W1 PC1 ON 0 0
P2 T2 W2 PC2 ON 0 0
P4 W2 W3 PC4 ON 2 0.5
PI W4 T3 * ON 0 0
[ORIFICES]
OR1 T1 T3 SIDE 1 0.65 YES 0
OR2 T2 T3 BOTTOM 0 0.6 NO 0.2
[WEIRS]
WS T1 T2 SIDEFLOW 6 3.0 NO 1 0 NO
WT T2 T3 TRAPEZOIDAL 5 3.2 YES 0 2.8 YES
[OUTLETS]
OL T3 W3 1 TABULAR/HEAD RATE NO
[XSECTIONS]
CA1 FILLED_CIRCULAR 3 0.5 0 0 1
CA2 HORSESHOE 3 0 0 0 1
CA3 ARCH 3 4.5 0 0 1
CA4 GOTHIC 3 0 0 0 1
CA5 CATENARY 3 0 0 0 1
CA6 SEMIELLIPTICAL 3 0 0 0 1
CA7 BASKETHANDLE 3 0 0 0 1
CA8 SEMICIRCULAR 3 0 0 0 1
CA9 MODBASKETHANDLE 4 3 1.5 0 1
CB1 TRAPEZOIDAL 4 3 1.5 2 1
CB2 RECT_OPEN 4 5 0 0 1
CB3 PARABOLIC 4 8 0 0 1
CB4 POWER 4 8 1.5 0 1
CB5 TRIANGULAR 4 8 0 0 1
CB6 IRREGULAR TR1
CB7 RECT_ROUND 4 3 2 0 1
CB8 RECT_TRIANGULAR 4 3 1 0 1
CX1 HORIZ_ELLIPSE 2.5 4 0 0 1
CX2 VERT_ELLIPSE 4 2.5 0 0 2
CO1 CUSTOM 3 SHP 0 0 1
CO2 EGG 3 0 0 0 1
CP4 CIRCULAR 2 0 0 0 1
OR1 CIRCULAR 1.5 0 0 0
OR2 RECT_CLOSED 1 1.5 0 0
WS RECT_OPEN 3 8 0 0
WT TRAPEZOIDAL 3 6 1 1
[TRANSECTS]
NC 0.05 0.05 0.03
X1 TR1 6 10 40 0 0 0 0 0
GR 6 0 4 10 0 15 0.5 35 4 40 5 45 6 50
[LOSSES]
CA2 0.3 0.2 0.1 NO 0
CB1 0 0 0 NO 0.5
CB2 0.2 0.2 0 NO 0.2
CX1 0 0 0 YES 0
[CURVES]
PC1 PUMP1 100 2 300 4 600 6 2000 8
PC2 PUMP2 1 1 2 2 4 3 8 4
PC4 PUMP4 0 0 1 1 3 3 6 4
RATE RATING 0 0 1 2 3 5 6 7
SHP SHAPE 0 0.2 0.25 0.7 0.5 1.0 0.75 0.8 1.0 0.0
[TIMESERIES]
QA 0:00 0
QA 1:00 12
QA 3:00 25
QA 6:00 4
QA 12:00 1
QB 0:00 0
QB 2:00 30
QB 5:00 8
QB 12:00 2
STG 0:00 100.5
STG 6:00 103
STG 12:00 100.5
CC 0:00 80
CC 12:00 80
[POLLUTANTS]
TSS MG/L 0 0 0 1.0
SALT MG/L 0 0 0 0
[INFLOWS]
A1 FLOW QA FLOW 1.0 1.0
A1 TSS CC CONCEN 1.0 1.0
B1 FLOW QB FLOW 1.0 1.0
B1 SALT CC CONCEN 1.0 0.5
A5 FLOW QA FLOW 1.0 0.3
[REPORT]
NODES ALL
LINKS ALL
"""


def c3b_shapes_inp() -> str:
    return C3B_SHAPES_INP


# Config 3c: SURVEY.md §8 row a28 -- culvert inlet control (four FHWA codes incl. a mitered one and a
# Form-2 box), three force mains under a pump, and three ROADWAY weirs (paved / gravel / fixed Cd)
# overtopped by the flood peak and later submerged by an outfall stage wave.
C3C_CULVERTS_INP = """[OPTIONS]
FLOW_UNITS CFS
FLOW_ROUTING DYNWAVE
START_DATE 01/01/2020
START_TIME 00:00:00
END_DATE 01/01/2020
END_TIME 08:00:00
REPORT_STEP 00:05:00
ROUTING_STEP 5
VARIABLE_STEP 0.75
ALLOW_PONDING NO
SURCHARGE_METHOD @SUR@
FORCE_MAIN_EQUATION @FME@
THREADS 1
[JUNCTIONS]
J1 110 12 0 0 0
J2 108 14 0 0 0
J3 106 14 0 0 0
J4 104 16 0 0 0
F1 115 1 0 100 0
F2 112 1 0 100 0
F3 109 1 0 100 0
[OUTFALLS]
O1 102 TIMESERIES STG NO
[STORAGE]
S1 95 12 1 FUNCTIONAL 0 0 200 0 0
[CONDUITS]
CH1 J1 J2 400 0.03 0 0 0 0
CU1 J2 J3 80 0.013 0 0 0 0
CU2 J2 J3 80 0.013 0.5 0 0 0
CU3 J2 J3 80 0.024 0 0 0 0
CU4 J2 J3 80 0.024 1 0 0 0
CH2 J3 J4 400 0.03 0 0 0 0
CH3 J4 O1 400 0.03 0 0 0 0
FM1 F1 F2 400 @RGH@ 0 0 0 0
FM2 F2 F3 400 @RGH@ 0 0 0 0
FM3 F3 J4 400 @RGH@ 0 3 0 0
[PUMPS]
P1 S1 F1 PC2 ON 0 0
[WEIRS]
RW1 J2 J3 ROADWAY 7 3.0 NO 0 0 NO 40 PAVED
RW2 J2 J3 ROADWAY 7.5 3.0 NO 0 0 NO 30 GRAVEL
RW3 J2 J3 ROADWAY 8 2.8
[XSECTIONS]
CH1 TRAPEZOIDAL 12 20 2 2 1
CH2 TRAPEZOIDAL 14 20 2 2 1
CH3 TRAPEZOIDAL 16 20 2 2 1
CU1 CIRCULAR 3 0 0 0 2 1
CU2 RECT_CLOSED 3 4 0 0 1 12
CU3 CIRCULAR 3 0 0 0 1 5
CU4 ARCH 3 4.5 0 0 1 37
FM1 FORCE_MAIN 1 @FMR@ 0 0 1
FM2 FORCE_MAIN 1 @FMR@ 0 0 1
FM3 FORCE_MAIN 1.25 @FMR@ 0 0 1
RW1 RECT_OPEN 5 100 0 0
RW2 RECT_OPEN 5 60 0 0
RW3 RECT_OPEN 5 40 0 0
[CURVES]
PC2 PUMP2 1 2 2 4 4 8 8 10
[TIMESERIES]
QC 0:00 0
QC 0:30 40
QC 1:00 150
QC 1:30 400
QC 2:00 1400
QC 2:30 600
QC 3:00 100
QC 4:00 20
QC 5:00 10
QC 8:00 5
QS 0:00 0
QS 0:20 3
QS 1:00 9
QS 2:00 6
QS 3:00 0.2
QS 4:00 0.05
QS 5:00 7
QS 8:00 1
STG 0:00 102.5
STG 4:00 103
STG 5:00 117
STG 6:00 117.5
STG 7:00 104
STG 8:00 103
[INFLOWS]
J1 FLOW QC FLOW 1.0 1.0
S1 FLOW QS FLOW 1.0 1.0
[REPORT]
NODES ALL
LINKS ALL
"""


def c3c_culverts_inp(eqn: str = "H-W") -> str:
    """eqn "H-W": Hazen-Williams + SLOT; "D-W": Darcy-Weisbach + EXTRAN (forcmain.c:95-118)."""
    sur, rough = ("SLOT", "120") if eqn == "H-W" else ("EXTRAN", "0.03")
    return (C3C_CULVERTS_INP.replace("@SUR@", sur).replace("@FME@", eqn)
            .replace("@RGH@", "0.012").replace("@FMR@", rough))


def c5_mega_spec(hours: float = 1.0) -> GridSpec:
    return GridSpec(nx=1000, ny=500, hours=hours, pollutants=False, surcharge="SLOT")


# --------------------------------------------------------------------------------------------
# config 3 at its stated size: ~1 000 links, mixed elements, 24 h, DWF patterns,
# control rules
# --------------------------------------------------------------------------------------------
@dataclass
class C3Spec:
    n: int = 22                  # n x n collector grid: 2 n (n - 1) conduits
    facilities: int = 8          # storage + orifice + weir + outlet + wet well + pump groups
    hours: float = 24.0
    surcharge: str = "SLOT"
    controls: bool = True
    pollutants: bool = True
    inflow_scale: float = 1.0    # member scale on the storm hydrographs (ensembles)
    seed: int = 3
    threads: int = 1


def c3_large_inp(spec: C3Spec | None = None) -> str:
    """Mixed-element network: an n x n sloping collector grid (circular / rect_closed / egg pipes and
    trapezoidal open channels along the low edge) and `facilities` treatment groups hung off it.  Each
    group: a storage unit (FUNCTIONAL or TABULAR) fed from the grid, emptied by a side or bottom
    orifice, a transverse or V-notch weir and a rating outlet into a lower grid node, plus a wet well
    with a pump (types 1-4 in turn) lifting to a higher grid node.  FREE, FIXED (flap gate) and TIDAL
    outfalls, DWF with hourly patterns, storm hydrographs, two pollutants, five control rules on pump
    status and orifice settings."""
    s = spec or C3Spec()
    n, F = s.n, s.facilities
    rng = random.Random(s.seed)
    L = ["[OPTIONS]", "FLOW_UNITS CFS", "FLOW_ROUTING DYNWAVE", "START_DATE 01/01/2020", "START_TIME 00:00:00",
         "END_DATE " + ("01/02/2020" if s.hours >= 24 else "01/01/2020"),
         "END_TIME " + ("00:00:00" if s.hours >= 24 else _hms(s.hours)), "REPORT_STEP 00:15:00", "ROUTING_STEP 5",
         "VARIABLE_STEP 0.75", "ALLOW_PONDING YES", f"SURCHARGE_METHOD {s.surcharge}", f"THREADS {s.threads}"]
    elev = lambda i, j: 100.0 + 0.5 * ((n - 1 - i) + (n - 1 - j))
    nid = lambda i, j: f"N{i}_{j}"
    L += ["[JUNCTIONS]"]
    for i in range(n):
        for j in range(n):
            pond = 300 if (i + j) % 7 == 0 else 0
            L.append(f"{nid(i, j)} {elev(i, j):.3f} 8 0 0 {pond}")
    for k in range(F):
        L.append(f"WW{k} {elev(*_fac_low(k, n, F)) - 6.0:.3f} 14 0 0 0")
    base = elev(n - 1, n - 1)
    L += ["[OUTFALLS]", f"OF {base - 1.0:.3f} FREE NO", f"OX {base - 1.5:.3f} FIXED {base + 0.5:.3f} YES",
          f"OT {base - 1.5:.3f} TIDAL TIDE NO"]
    L += ["[STORAGE]"]
    for k in range(F):
        z = elev(*_fac_out(k, n, F)) + 0.5           # above the node it drains to, below its feeder
        if k % 2 == 0:
            L.append(f"S{k} {z:.3f} 12 1 FUNCTIONAL {800 + 100 * k} 0 300 0 0")
        else:
            L.append(f"S{k} {z:.3f} 12 0.5 TABULAR SCURVE 0 0")
    cond, xs, losses = [], [], []
    shapes = ["CIRCULAR 2 0 0 0 1", "RECT_CLOSED 2 2.5 0 0 1", "EGG 2.5 0 0 0 1", "CIRCULAR 2.5 0 0 0 1"]
    c = 0
    for i in range(n):
        for j in range(n):
            for (di, dj) in ((1, 0), (0, 1)):
                ii, jj = i + di, j + dj
                if ii >= n or jj >= n:
                    continue
                cid = f"C{c}"
                off2 = 0.2 if c % 37 == 5 else 0
                cond.append(f"{cid} {nid(i, j)} {nid(ii, jj)} 300 0.013 0 {off2} 0 0")
                if i == n - 1 or j == n - 1:
                    xs.append(f"{cid} TRAPEZOIDAL 6 4 1 1 1")       # open collector along the low edges
                else:
                    xs.append(f"{cid} {shapes[c % len(shapes)]}")
                if c % 53 == 7:
                    losses.append(f"{cid} 0.3 0.3 0 NO 0")
                c += 1
    cond.append(f"COF {nid(n - 1, n - 1)} OF 300 0.013 0 0 0 0")
    xs.append("COF TRAPEZOIDAL 6 6 1 1 1")
    cond.append(f"COX {nid(n - 1, n - 3)} OX 300 0.013 0 0 0 0")
    xs.append("COX CIRCULAR 4 0 0 0 2")
    cond.append(f"COT {nid(n - 3, n - 1)} OT 300 0.013 0 0 0 0")
    xs.append("COT CIRCULAR 4 0 0 0 1")
    pumps, orifs, weirs, outlets = [], [], [], []
    for k in range(F):
        a, b, lo, hi = _fac_in(k, n, F), _fac_out(k, n, F), _fac_low(k, n, F), _fac_high(k, n, F)
        cond.append(f"CS{k} {nid(*a)} S{k} 200 0.013 0 1 0 0")
        xs.append(f"CS{k} CIRCULAR 3 0 0 0 1")
        if k % 2 == 0:
            orifs.append(f"OR{k} S{k} {nid(*b)} SIDE 0.5 0.65 NO 0")
            xs.append(f"OR{k} CIRCULAR 1.5 0 0 0")
            weirs.append(f"W{k} S{k} {nid(*b)} TRANSVERSE 8 3.33 NO 0 0 YES")
            xs.append(f"W{k} RECT_OPEN 3 8 0 0")
        else:
            orifs.append(f"OR{k} S{k} {nid(*b)} BOTTOM 0 0.6 NO 0.1")
            xs.append(f"OR{k} RECT_CLOSED 1 1.5 0 0")
            weirs.append(f"W{k} S{k} {nid(*b)} V-NOTCH 6 2.5 NO 0 0 YES")
            xs.append(f"W{k} TRIANGULAR 4 6 0 0")
        if k % 4 == 1:
            outlets.append(f"OL{k} S{k} {nid(*b)} 2 FUNCTIONAL/DEPTH 2.0 0.5 NO")
        elif k % 4 == 3:
            outlets.append(f"OL{k} S{k} {nid(*b)} 2 TABULAR/HEAD RATE NO")
        cond.append(f"CW{k} {nid(*lo)} WW{k} 150 0.013 0 0 0 0")
        xs.append(f"CW{k} CIRCULAR 2 0 0 0 1")
        ptype = k % 4
        if ptype == 0:            # TYPE1 needs a storage inlet: pump straight from the storage unit
            pumps.append(f"P{k} S{k} {nid(*hi)} PC1 ON 0 0")
            pumps.append(f"PW{k} WW{k} {nid(*hi)} PC3 ON 4 1")
        elif ptype == 1:
            pumps.append(f"P{k} WW{k} {nid(*hi)} PC2 ON 0 0")
        elif ptype == 2:
            pumps.append(f"P{k} WW{k} {nid(*hi)} PC3 ON 4 1")
        else:
            pumps.append(f"P{k} WW{k} {nid(*hi)} PC4 ON 2 0.5")
    L += ["[CONDUITS]"] + cond + ["[PUMPS]"] + pumps + ["[ORIFICES]"] + orifs + ["[WEIRS]"] + weirs
    L += ["[OUTLETS]"] + outlets + ["[XSECTIONS]"] + xs + ["[LOSSES]"] + losses
    if s.controls:
        L += ["[CONTROLS]",
              "RULE R1", "IF NODE S0 DEPTH > 6", "THEN ORIFICE OR0 SETTING = 1.0", "ELSE ORIFICE OR0 SETTING = 0.3",
              "PRIORITY 1",
              "RULE R2", "IF SIMULATION TIME > 12", "THEN PUMP P2 STATUS = OFF",
              "RULE R3", "IF NODE S1 DEPTH > 5", "AND LINK OR1 FLOW < 20", "THEN ORIFICE OR1 SETTING = 1.0",
              "ELSE ORIFICE OR1 SETTING = 0.5", "PRIORITY 2",
              "RULE R4", f"IF NODE {nid(n - 1, n - 1)} DEPTH > 3", "OR SIMULATION TIME > 20",
              "THEN PUMP P1 STATUS = OFF", "ELSE PUMP P1 STATUS = ON",
              "RULE R5", "IF LINK COF FLOW > 60", "THEN WEIR W2 SETTING = 0.5", "ELSE WEIR W2 SETTING = 1.0"]
    if s.pollutants:
        L += ["[POLLUTANTS]", "TSS MG/L 0 0 0 0.5", "DYE MG/L 0 0 0 0"]
    L += ["[INFLOWS]"]
    dwf = []
    q = 0
    for i in range(n):
        for j in range(n):
            if (i * n + j) % 3 == 0 and i + j < 2 * n - 6:
                sc = rng.uniform(0.1, 0.3) * s.inflow_scale
                txt = f"{sc:.6f}" if s.inflow_scale == 1.0 else repr(float(f"{sc / s.inflow_scale:.6f}") * s.inflow_scale)
                L.append(f"{nid(i, j)} FLOW STORM FLOW 1.0 {txt}")
                if s.pollutants and q % 2 == 0:
                    L.append(f"{nid(i, j)} TSS CTSS CONCEN 1.0 1.0")
                if s.pollutants and q % 5 == 0:
                    L.append(f"{nid(i, j)} DYE CDYE CONCEN 1.0 1.0")
                q += 1
            if (i * n + j) % 11 == 4:
                dwf.append(f"{nid(i, j)} FLOW 0.05 DAILYP")
    L += ["[DWF]"] + dwf
    L += ["[PATTERNS]", "DAILYP HOURLY 0.5 0.5 0.5 0.5 0.6 0.8 1.2 1.5 1.4 1.2 1.1 1.0 1.0 1.0 1.0 1.1 1.2 1.4 1.5 1.3 1.0 "
          "0.8 0.6 0.5"]
    L += ["[CURVES]", "SCURVE STORAGE 0 800 5 1500 10 3000 12 3500", "PC1 PUMP1 100 2 300 4 600 6 2000 8",
          "PC2 PUMP2 1 1 2 2 4 3 8 4", "PC3 PUMP3 0 8 5 6 10 3 15 0", "PC4 PUMP4 0 0 1 1 3 3 6 4",
          "RATE RATING 0 0 1 2 3 5 6 7", f"TIDE TIDAL 0 {base - 1.0:.2f} 6 {base + 0.5:.2f} 12 {base - 1.0:.2f} "
          f"18 {base + 0.5:.2f} 24 {base - 1.0:.2f}"]
    L += ["[TIMESERIES]", "STORM 0:00 0", "STORM 2:00 0.5", "STORM 5:00 3.0", "STORM 9:00 0.8", "STORM 16:00 0.2",
          "STORM 24:00 0.1", "CTSS 0:00 100", "CTSS 48:00 100", "CDYE 0:00 50", "CDYE 48:00 50"]
    L += ["[REPORT]", "NODES ALL", "LINKS ALL", "CONTROLS YES"]
    return "\n".join(L) + "\n"


def _fac_in(k, n, F):       # grid node feeding storage k (upper part of the grid)
    return (2 + (k * (n - 6)) // max(F - 1, 1), 3 + (k * 5) % (n - 8))


def _fac_out(k, n, F):      # lower grid node the storage drains to
    i, j = _fac_in(k, n, F)
    return (min(i + 3, n - 2), min(j + 3, n - 2))


def _fac_low(k, n, F):      # grid node draining into wet well k
    i, j = _fac_out(k, n, F)
    return (min(i + 1, n - 2), j)


def _fac_high(k, n, F):     # grid node the pump lifts to
    i, j = _fac_in(k, n, F)
    return (max(i - 1, 0), max(j - 1, 0))
