#!/bin/bash
# Round 2 call Y: config 5 on the GPU against the reference fixture
( time timeout 170 python -m pytest tests/test_c5_golden.py -m gpu -x -q ) 2>&1 | tail -n 8
