"""Run the serial oracle on the published benchmark grid (MSVC seed flavour) and store its
status / pivot counts / objective / pivot-sequence hash as fixtures for the GPU parity tests.

    make -C oracle && python tests/golden/make_oracle_results.py [max_constraints] [threads]

The oracle is deterministic, so the fixture can be regenerated at will; sizes above
max_constraints (default 2048) are skipped because a single host core needs minutes to hours.
"""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
maxc = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
threads = sys.argv[2] if len(sys.argv) > 2 else "4"
pub = json.load(open(os.path.join(root, "tests/golden/published_pivot_counts.json")))
res_path = os.path.join(root, "tests/golden/oracle_results.json")
res = json.load(open(res_path)) if os.path.exists(res_path) else {}
for inst in pub["instances"]:
    n, m, seed = inst["vars"], inst["constraints"], inst["seed"]
    key = f"{n}_{m}_{seed}"
    if m > maxc or n > maxc * 4 or key in res:
        continue
    out = subprocess.run([os.path.join(root, "oracle/serial_tableau"), str(n), str(m), str(seed), "1", "100",
                          "1", "0", threads, "-1"], capture_output=True, text=True, check=True).stdout.split()
    res[key] = {"vars": n, "constraints": m, "seed": seed, "flavour": "msvc", "status": int(out[0]),
                "pivots_phase1": int(out[1]), "pivots_phase2": int(out[2]), "objective": float(out[3]),
                "objective_repr": out[3], "trace_hash": out[4], "seconds": float(out[5])}
    ok = (res[key]["pivots_phase1"] == inst["pivots_phase1"]
          and (not inst["phase2_ran"] or res[key]["pivots_phase2"] == inst["pivots_phase2"]))
    print(key, out, "MATCH" if ok else f"MISMATCH vs {inst}", flush=True)
    json.dump(res, open(res_path, "w"), indent=1, sort_keys=True)
