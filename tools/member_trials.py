#!/usr/bin/env python
"""Which members need many Picard trials?  Runs the config-4 ensemble (as generated), spins it up like the
bench, steps it and prints the correlation of the per-member trial counts with hydrograph scale / shift and
their step-to-step persistence (input for choosing a member enumeration)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from swmm_b200 import scenarios  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=1024)
ap.add_argument("--out", default=None)
a = ap.parse_args()
args = argparse.Namespace(grid=100, hours=6.0, surcharge="SLOT", members=a.members, members_total=a.members,
                          member_order="generated")
s, case, spec = bench.make_ensemble(args, 0, 0)
scale, shift = scenarios.c4_members(a.members, 2024)
bench.spin_up(s, 6000.0)
prev = np.array([x.iterations for x in s.stats(0, s.M)])
hist = []
for k in range(6):
    s.run_steps(10, case.t_end)
    cur = np.array([x.iterations for x in s.stats(0, s.M)])
    hist.append(cur - prev)
    prev = cur
h = np.array(hist, dtype=float)          # [block][member] trials per 10 steps
tot = h.sum(0)
def rank(v): return np.argsort(np.argsort(v)).astype(float)
def rcorr(x, y): return float(np.corrcoef(rank(x), rank(y))[0, 1])
rec = {"members": a.members, "trials_per_step_mean": float(tot.mean() / 60), "min": float(tot.min() / 60), "max": float(tot.max() / 60),
       "rank_corr_scale": rcorr(tot, scale), "rank_corr_shift": rcorr(tot, shift),
       "rank_corr_block_to_next": [rcorr(h[i], h[i + 1]) for i in range(len(h) - 1)],
       "rank_corr_first_to_last_block": rcorr(h[0], h[-1])}
print(json.dumps(rec))
if a.out:
    np.savez(a.out, trials=h, scale=scale, shift=shift)
s.close()
