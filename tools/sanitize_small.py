"""Small solves through every code path, meant to be run under compute-sanitizer on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

import oracle_py as O
import simplexoncuda_b200 as S

for opts in (dict(), dict(persistent=False), dict(persistent=False, use_graph=False), dict(fold_artificials=False),
             dict(skip_zero_rows=True), dict(persistent=True, skip_zero_rows=True), dict(dtype=S.F32), dict(pivot_rule=2)):
    for (n, m, lo) in ((40, 24, -100), (130, 70, 1), (64, 520, 1)):
        A, b, c = O.generate(n, m, O.seed_triplet(3, 0), lo, 100)
        with S.Solver(max_pivots=5000, **opts) as s:
            s.load(A, b, c)
            r = s.solve()
        print(opts, n, m, r["status"], r["stats"].pivots_phase1, r["stats"].pivots_phase2, flush=True)
with S.Solver() as s:
    s.generate(100, 200, (1, 2, 3), 1, 100)
    s.copy_problem()
    print("tournament", s.tournament(np.random.default_rng(0).normal(size=5000)))
print("done")
