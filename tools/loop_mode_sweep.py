#!/usr/bin/env python
"""Complete solves of the reference's published instances with each loop body (GPU box): persistent cooperative loop kernel,
three launches per pivot, look-ahead kernel (one launch per pivot).  Prints pivots/s (device time of the pivot loops) per
size and mode; all modes must give the same pivot sequence.   python tools/loop_mode_sweep.py [n,m ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplexoncuda_b200 as S

sizes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(1024, 1024), (2048, 1024), (2048, 2048), (4096, 2048),
                                                                         (4096, 4096), (8192, 2048), (8192, 4096)]
modes = {"persistent": dict(persistent=True), "launches": dict(persistent=False, lookahead=False),
         "lookahead": dict(persistent=False, lookahead=True), "auto": dict()}
for n, m in sizes:
    seeds = S.seed_triplet(n * 100 + m, S.RAND_MSVC)
    row = {"n": n, "m": m}
    hashes = set()
    for name, opts in modes.items():
        for helpers in ((8, 16) if name == "lookahead" else (0,)):
            if helpers:
                os.environ["B2S_LA_HELPERS"] = str(helpers)
            else:
                os.environ.pop("B2S_LA_HELPERS", None)
            with S.Solver(**opts) as s:
                s.generate(n, m, seeds, 1, 100)
                r = s.solve()
                st = r["stats"]
                piv = st.pivots_phase1 + st.pivots_phase2
                key = name + (f"_h{helpers}" if helpers else "")
                row[key] = round(piv / (st.seconds_phase1 + st.seconds_phase2), 1)
                if name == "auto":
                    row["auto_mode"] = s.loop_mode()[:24]
                    row["stored_MB"] = round(s.dims()["rows_stored"] * m * 8 / 1e6, 1)
                hashes.add(int(st.trace_hash))
    assert len(hashes) == 1, "pivot sequences differ between loop modes"
    print(json.dumps(row), flush=True)
