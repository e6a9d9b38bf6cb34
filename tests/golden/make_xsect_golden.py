#!/usr/bin/env python
"""Writes tests/golden/xsect_golden.npz: dense sweeps of the reference's xsect_* functions
(called directly in oracle/_ref/libswmm5.so) for every analytic / tabulated shape."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity_common as pc  # noqa: E402
import refengine  # noqa: E402

SHAPES = {
    "CIRCULAR": (1, [1.5, 0, 0, 0]), "FILLED_CIRCULAR": (2, [3, 0.5, 0, 0]),
    "RECT_CLOSED": (3, [1.5, 1.5, 0, 0]), "RECT_OPEN": (4, [3, 10, 0, 0]),
    "TRAPEZOIDAL": (5, [4, 3, 1, 1]), "TRIANGULAR": (6, [4, 6, 0, 0]), "PARABOLIC": (7, [4, 6, 0, 0]),
    "POWERFUNC": (8, [4, 6, 1.5, 0]), "RECT_TRIANG": (9, [5, 4, 1.5, 0]), "RECT_ROUND": (10, [5, 4, 2.5, 0]),
    "MOD_BASKET": (11, [5, 4, 2.5, 0]), "HORIZ_ELLIPSE": (12, [3, 5, 0, 0]), "VERT_ELLIPSE": (13, [5, 3, 0, 0]),
    "ARCH": (14, [3, 5, 0, 0]), "EGG": (15, [3, 0, 0, 0]), "HORSESHOE": (16, [3, 0, 0, 0]),
    "GOTHIC": (17, [3, 0, 0, 0]), "CATENARY": (18, [3, 0, 0, 0]), "SEMIELLIPTICAL": (19, [3, 0, 0, 0]),
    "BASKETHANDLE": (20, [3, 0, 0, 0]), "SEMICIRCULAR": (21, [3, 0, 0, 0]),
}

e = refengine.RefEngine()
rng = np.random.default_rng(0)
out = {"shapes": np.array(list(SHAPES))}
for name, (t, g) in SHAPES.items():
    ok, p = e.xsect_set(t, g)
    assert ok, name
    yF, aF, sF, sM = p[0], p[3], p[5], p[6]
    ys = np.concatenate([np.linspace(0, yF, 101), rng.uniform(0, yF, 300), [1e-4, 1e-6, yF * 0.999999]])
    As = np.concatenate([np.linspace(0, aF, 101), rng.uniform(0, aF, 300),
                         aF * np.array([1e-7, 1e-5, 1e-3, 0.039, 0.041, 0.97, 0.9999])])
    Ss = np.concatenate([np.linspace(0, sM, 101), rng.uniform(0, sM, 200),
                         sF * np.array([1e-7, 1e-5, 0.014, 0.016, 1.0])])
    Qs = np.concatenate([np.linspace(0, 50, 80), rng.uniform(0, 5, 120)])
    out[f"{name}_type"] = np.array(t)
    out[f"{name}_params"] = p
    for fn, args in (("AofY", ys), ("WofY", ys), ("RofY", ys), ("YofA", As), ("RofA", As), ("SofA", As),
                     ("AofS", Ss), ("dSdA", As), ("Ycrit", Qs)):
        out[f"{name}_{fn}_x"] = args
        out[f"{name}_{fn}_y"] = e.xsect_eval(fn, t, p, args)
np.savez_compressed(os.path.join(pc.GOLDEN, "xsect_golden.npz"), **out)
print("wrote xsect_golden.npz", os.path.getsize(os.path.join(pc.GOLDEN, "xsect_golden.npz")), "bytes")
