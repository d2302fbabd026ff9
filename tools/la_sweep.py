#!/usr/bin/env python
"""Pivots/s of the pivot loop at one size for a few loop settings (GPU box).  python tools/la_sweep.py [n m pivots]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplexoncuda_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
K = int(sys.argv[3]) if len(sys.argv) > 3 else 500
seeds = S.seed_triplet(n * 100 + m, S.RAND_MSVC)
rows = []
for skip in (True, False):
    for la, helpers in ((False, 8), (True, 2), (True, 4), (True, 8), (True, 16)):
        os.environ["B2S_LA_HELPERS"] = str(helpers)
        with S.Solver(skip_zero_rows=skip, lookahead=la, persistent=False) as s:
            s.generate(n, m, seeds, 1, 100)
            s.build_phase1(); s.price_out(); s.select_entering()
            s.iterate(300)
            t0 = s.stats().seconds_phase1
            st, done = s.iterate(K)
            dt = s.stats().seconds_phase1 - t0
            _, _, h = s.trace()
            rows.append({"skip_zero_rows": skip, "lookahead": la, "helpers": helpers if la else None,
                         "pivots_per_s": done / dt, "us_per_pivot": 1e6 * dt / done, "hash": str(h)})
            print(json.dumps(rows[-1]), flush=True)
assert len({r["hash"] for r in rows}) == 1, "pivot sequences differ"
