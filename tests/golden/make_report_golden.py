#!/usr/bin/env python
"""Writes tests/golden/report_<case>.npz from the UNMODIFIED reference engine (oracle/_ref): the flat
network, state snapshots at a few routing steps (every field swb_get_results reads) and the engine's
own report records for them -- node_getResults (node.c:497-528) / link_getResults (link.c:674-724)
called on the live engine at weighting factors 0, 0.37 and 1.
    python tests/golden/make_report_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity_common as pc  # noqa: E402

FIELDS = ["SWB_NODE_NEW_DEPTH", "SWB_NODE_OLD_DEPTH", "SWB_NODE_NEW_VOLUME", "SWB_NODE_OLD_VOLUME",
          "SWB_NODE_NEW_LATFLOW", "SWB_NODE_OLD_LATFLOW", "SWB_NODE_INFLOW", "SWB_NODE_OLD_INFLOW",
          "SWB_NODE_OVERFLOW", "SWB_NODE_NEW_QUAL", "SWB_NODE_OLD_QUAL", "SWB_LINK_NEW_FLOW",
          "SWB_LINK_OLD_FLOW", "SWB_LINK_NEW_DEPTH", "SWB_LINK_OLD_DEPTH", "SWB_LINK_NEW_VOLUME",
          "SWB_LINK_OLD_VOLUME", "SWB_LINK_SETTING", "SWB_LINK_NEW_QUAL", "SWB_LINK_OLD_QUAL"]
F = [0.0, 0.37, 1.0]
CASES = {"c2_grid12_slot": [200, 700, 1200], "c3_mixed": [300, 900, 1500], "c3b_shapes": [300, 900]}

for name, steps in CASES.items():
    e, _ = pc.open_reference(pc.case_inp(name))
    net = e.network()
    extra = {"f": np.array(F), "steps": np.array(steps)}
    n = 0
    for k, target in enumerate(steps):
        while n < target:
            if e.step() == 0:
                break
            n += 1
        for fld in FIELDS:
            extra[f"s{k}_{fld}"] = e.field(fld).copy()
        for i, f in enumerate(F):
            nd, ld = e.results(f, net.n_nodes, net.n_links, net.n_pollut)
            extra[f"node_{k}_{i}"], extra[f"link_{k}_{i}"] = nd, ld
    e.end()
    e.close()
    path = os.path.join(HERE, f"report_{name}.npz")
    net.save(path, **extra)
    print(name, n, os.path.getsize(path), "bytes")
