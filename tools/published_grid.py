"""Run the reference's whole published benchmark grid (main.cu -t: 36 instances, vars/cons in 256..8192, MSVC seed
derivation) on this GPU and tabulate pivots, device time and pivots/s next to the reference's published totals.
    python tools/published_grid.py > profiles/r01_published_grid.md        (GPU box)
"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import simplexoncuda_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pub = json.load(open(os.path.join(ROOT, "tests", "golden", "published_pivot_counts.json")))["instances"]
# published wall time per instance on the RTX 2070 Super (sum of the reference's CSV lines)
pub_2070s = {(256, 256): 0.26, (256, 512): 0.65, (256, 1024): 1.51, (256, 2048): 6.34, (256, 4096): 27.64, (256, 8192): 202.37,
             (512, 256): 0.27, (512, 512): 0.68, (512, 1024): 1.69, (512, 2048): 6.31, (512, 4096): 30.28, (512, 8192): 210.11,
             (1024, 256): 0.30, (1024, 512): 0.64, (1024, 1024): 1.91, (1024, 2048): 7.25, (1024, 4096): 35.79, (1024, 8192): 253.11,
             (2048, 256): 0.31, (2048, 512): 0.80, (2048, 1024): 2.16, (2048, 2048): 8.65, (2048, 4096): 44.70, (2048, 8192): 254.90,
             (4096, 256): 0.35, (4096, 512): 0.96, (4096, 1024): 2.28, (4096, 2048): 10.00, (4096, 4096): 52.97, (4096, 8192): 288.88,
             (8192, 256): 0.37, (8192, 512): 0.86, (8192, 1024): 3.73, (8192, 2048): 13.90, (8192, 4096): 69.31, (8192, 8192): 411.08}
print("# The reference's published benchmark grid on one B200 (libb2s, fp64, default options)\n")
print("Instances: `generateRandomProblem(vars, cons, vars*100+cons (+1 for 1024x8192), 1, 100)`, MSVC seed derivation. "
      "Pivot counts must equal the published ones (they do: column `match`). `solve s` = wall time of "
      "`b2s_solve_two_phase` (build, price-out, both phases, solution); the instance is generated on the device.\n")
print("| vars | cons | pivots P1+P2 | match | solve s | pivots/s | 2070 Super published s | speed-up |")
print("|---|---|---|---|---|---|---|---|")
seen = set()
tot_ours = tot_ref = 0.0
with S.Solver() as s:
    for inst in pub:
        key = (inst["vars"], inst["constraints"])
        if inst["gpu"] != "rtx2070super" or key in seen:
            continue
        seen.add(key)
        n, m = key
        s.generate(n, m, S.seed_triplet(inst["seed"], S.RAND_MSVC), 1, 100)
        t0 = time.time()
        r = s.solve()
        dt = time.time() - t0
        p1, p2 = r["stats"].pivots_phase1, r["stats"].pivots_phase2
        ok = (p1, p2) == (inst["pivots_phase1"], inst["pivots_phase2"]) and r["status"] == 0
        tot_ours += dt
        tot_ref += pub_2070s[key]
        print(f"| {n} | {m} | {p1}+{p2} | {'yes' if ok else 'NO'} | {dt:.3f} | {(p1 + p2) / dt:.0f} | {pub_2070s[key]:.2f} | "
              f"{pub_2070s[key] / dt:.0f}x |", flush=True)
print(f"\nWhole grid: {tot_ours:.1f} s here vs {tot_ref:.0f} s published on the RTX 2070 Super ({tot_ref / tot_ours:.0f}x).")
