#!/usr/bin/env python
"""Per-stage timing of the look-ahead pivot kernel (b2s_profile_lookahead), single GPU or under torchrun (sharded).

    python tools/la_profile.py [n m count]                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N tools/la_profile.py   # N ranks, constraint slabs

Prints one JSON line per rank: mean / p50 / max of the kernel time and of the chain's milestones (us from kernel start)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import simplexoncuda_b200 as S
from simplexoncuda_b200 import sharding

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
count = int(sys.argv[3]) if len(sys.argv) > 3 else 200
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = S.Solver(device=local, skip_zero_rows=os.environ.get("B2S_SKIP", "1") != "0")
if world > 1:
    sharding.init_sharded_solver(s, dist)
s.generate(n, m, S.seed_triplet(n * 100 + m, S.RAND_MSVC), 1, 100)
s.build_phase1(); s.price_out(); s.select_entering()
s.iterate(300)
t0 = s.stats().seconds_phase1
s.iterate(400)
free_run = (s.stats().seconds_phase1 - t0) / 400 * 1e6
pr = s.profile_lookahead(count)
out = {"rank": rank, "world": world, "n": n, "m": m, "loop": s.loop_mode(), "free_running_us_per_pivot": free_run,
       "pivots_profiled": int(pr["pivots"])}
for k, v in pr.items():
    if k == "pivots":
        continue
    v = np.asarray(v, dtype=np.float64) * (1e3 if k == "kernel_ms" else 1.0)
    v = v[v >= 0]
    out[k.replace("kernel_ms", "kernel_us")] = {"mean": float(v.mean()), "p50": float(np.median(v)), "p90": float(np.percentile(v, 90)),
                                               "max": float(v.max())}
print(json.dumps(out), flush=True)
s.close()
if dist is not None:
    dist.destroy_process_group()
