"""Sweep the rank-1 update kernel variants (b2s_bench_update) at a given size.  GPU only.
    python tools/kbench.py [n] [m] [launches] [variants comma list] [tile_groups comma list]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import simplexoncuda_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 30
variants = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else list(range(15))
tgs = sys.argv[5].split(",") if len(sys.argv) > 5 else [""]
names = {0: "v16 u8", 1: "v16 u8 .cs", 2: "v32 u4", 3: "v32 u4 .cs", 4: "v32 u8", 5: "v16 u4", 6: "v16 u8 noalloc",
         7: "v32 u8 .cs", 8: "v32 u8 dyn", 9: "v32 u4 dyn", 10: "v16 u4 dyn", 11: "v32 u8 .cs dyn", 12: "v32 u8 noalloc dyn",
         13: "v16 u8 dyn", 14: "bulk-copy (TMA) ring"}
for tg in tgs:
    if tg:
        os.environ["B2S_TILE_GROUPS"] = tg
    for var in variants:
        with S.Solver(update_variant=var, persistent=False) as s:
            s.generate(n, m, (1, 2, 3), 1, 100)
            ms, nbytes = s.bench_update(launches, flush_l2=False)
        ms = ms[3:]
        med = float(np.median(ms)); best = float(ms.min())
        print(json.dumps({"n": n, "m": m, "tg": tg, "variant": var, "name": names.get(var), "median_ms": round(med, 4),
                          "best_ms": round(best, 4), "GBps_median": round(nbytes / med / 1e6, 1),
                          "GBps_best": round(nbytes / best / 1e6, 1)}), flush=True)
