"""Extract the reference's published per-phase pivot counts into a small fixture.

Run once in the build container (needs /root/reference):
    python tests/golden/make_published_counts.py

Source: /root/reference/data/measures/<gpu>/benchmark_<vars>_<cons>.txt, written by the reference's
-D TIMER build (src/chrono.cu:35-50): one `solve` line per simplex iteration, the first CSV
column is the active device-row count (1+n+2m in phase 1, 1+n+m in phase 2).  The last `solve`
line of a phase is the terminating optimality check, so pivots = lines - 1.
Seeds: main.cu:63 (vars*100+cons, +1 for 1024x8192); range [1,100] (main.cu:64).
"""
import json
import os
import re

REF = "/root/reference/data/measures"
out = {"source": "rik1599/SimplexOnCuda data/measures/{rtx2070super,mx250_2}/benchmark_<vars>_<cons>.txt",
       "seed_flavour": "msvc", "range": [1, 100], "instances": []}
for gpu in ("rtx2070super", "mx250_2"):
    for fn in sorted(os.listdir(os.path.join(REF, gpu))):
        mt = re.match(r"benchmark_(\d+)_(\d+)\.txt", fn)
        if not mt:
            continue
        n, m = int(mt.group(1)), int(mt.group(2))
        r1, r2 = 1 + n + 2 * m, 1 + n + m
        c1 = c2 = 0
        for line in open(os.path.join(REF, gpu, fn)):
            f = line.strip().split(",")
            if len(f) == 4 and f[2] == "solve":
                if int(f[0]) == r1:
                    c1 += 1
                elif int(f[0]) == r2:
                    c2 += 1
        seed = n * 100 + m + (1 if (n == 1024 and m == 8192 and gpu == "rtx2070super") else 0)
        out["instances"].append({"gpu": gpu, "vars": n, "constraints": m, "seed": seed,
                                 "pivots_phase1": c1 - 1, "pivots_phase2": max(c2 - 1, -1),
                                 "phase2_ran": c2 > 0})
with open(os.path.join(os.path.dirname(__file__), "published_pivot_counts.json"), "w") as f:
    json.dump(out, f, indent=1)
print(len(out["instances"]), "instances")
