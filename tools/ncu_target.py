#!/usr/bin/env python
"""Smallest program that runs the default hot path on the metric config (8192 x 8192): 24 pivots of phase 1 with the library's
default options, then exit.  This is the command the ncu captures under profiles/ are taken from (tools/README in profiles/).
    python tools/ncu_target.py [--no-skip]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplexoncuda_b200 as S

skip = "--no-skip" not in sys.argv
with S.Solver(skip_zero_rows=skip, use_graph=False) as s:
    s.generate(8192, 8192, S.seed_triplet(827392, S.RAND_MSVC), 1, 100)
    s.build_phase1(); s.price_out(); s.select_entering()
    st, done = s.iterate(24)
    assert done == 24, (st, done)
    print("ok", s.loop_mode(), s.trace()[2])
