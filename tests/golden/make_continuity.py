#!/usr/bin/env python
"""Writes tests/golden/continuity.json from the UNMODIFIED reference engine (oracle/_ref): for every
golden case the flow-routing and quality continuity errors of swmm_getMassBalErr (swmm5.c) after a
complete run, and the engine's routing totals (massbal.c: FlowTotals / QualTotals, ft3 and mass).
    python tests/golden/make_continuity.py
"""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity_common as pc  # noqa: E402

CASES = ["c1_tree", "c1_tree_slot", "c2_grid12_slot", "c2_grid12_extran"]
FIELDS = ["dwInflow", "wwInflow", "gwInflow", "iiInflow", "exInflow", "flooding", "outflow", "evapLoss",
          "seepLoss", "reacted", "initStorage", "finalStorage", "pctError"]


class Totals(C.Structure):            # objects.h:921-936
    _fields_ = [(f, C.c_double) for f in FIELDS]


out = {}
for name in CASES:
    e, _ = pc.open_reference(pc.case_inp(name))
    nP = e.network().n_pollut
    steps = 0
    while e.step() != 0:
        steps += 1
    qrec = None
    if nP:                             # QualTotals is freed by swmm_end: read it first (its
        qt = C.POINTER(Totals).in_dll(e.lib, "QualTotals")        # finalStorage lacks the stored mass then)
        qrec = [{"ex_inflow": qt[p].exInflow, "flooding": qt[p].flooding, "outflow": qt[p].outflow,
                 "reacted": qt[p].reacted, "seepage": qt[p].seepLoss, "init_storage": qt[p].initStorage,
                 "final_storage": qt[p].finalStorage} for p in range(nP)]
    e.end()
    _, flow_err, qual_err = e.mass_bal_err()
    ft = Totals.in_dll(e.lib, "FlowTotals")
    rec = {"steps": steps + 1, "flow_error_pct": flow_err, "qual_error_pct": qual_err,
           "flow_totals_ft3": {"ex_inflow": ft.exInflow, "flooding": ft.flooding, "outflow": ft.outflow,
                               "evap_loss": ft.evapLoss, "seep_loss": ft.seepLoss},
           "init_storage_ft3": ft.initStorage, "final_storage_ft3": ft.finalStorage}
    if qrec:
        rec["qual_totals"] = qrec
    e.close()
    out[name] = rec
    print(name, rec["steps"], flow_err, qual_err)
json.dump(out, open(os.path.join(HERE, "continuity.json"), "w"), indent=1)
