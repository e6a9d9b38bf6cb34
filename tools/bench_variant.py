#!/usr/bin/env python
"""Runs bench.py against an alternative build of libswmm_b200.so (register / block-size sweeps).
    python tools/bench_variant.py variants/libswmm_b200_b128x5.so [bench.py args...]"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swmm_b200  # noqa: E402,F401
from swmm_b200 import solver  # noqa: E402

solver.CUDA_LIB = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
