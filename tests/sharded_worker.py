"""Worker of tests/test_sharded_gpu.py and tests/test_sharded_same_gpu.py: launched by torchrun.  Solves the LPs listed in
argv sharded over the ranks through the C ABI and prints one JSON line per LP on rank 0.

Default: one rank per GPU, NCCL bootstrap.  B2S_TEST_SAME_GPU=1: every rank uses cuda:0, torch.distributed runs on gloo and
the library is bootstrapped through the host layer (b2s_dist_init_host) -- the peer-memory kernels then exchange through plain
device memory mapped with CUDA IPC, which is how a single-GPU box exercises the sharded code paths.
B2S_EXPECT_PEER_ERROR=1: the solve must fail with B2S_ERR_PEER on every rank (fault injection, B2S_FAULT_RANK/B2S_FAULT_PIVOT)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import simplexoncuda_b200 as S
from simplexoncuda_b200 import sharding


def main():
    same_gpu = os.environ.get("B2S_TEST_SAME_GPU") == "1"
    expect_peer_error = os.environ.get("B2S_EXPECT_PEER_ERROR") == "1"
    local = 0 if same_gpu else int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = json.loads(sys.argv[1])
    # B2S_TEST_PERSISTENT=1 forces the persistent loop kernel (with peer-memory exchanges) on the sharded solve
    s = S.Solver(device=local, persistent=True if os.environ.get("B2S_TEST_PERSISTENT") == "1" else "auto")
    if same_gpu or os.environ.get("B2S_TEST_HOST_BOOTSTRAP") == "1":
        sharding.init_sharded_solver_host(s, dist)
    else:
        sharding.init_sharded_solver(s, dist)
    for cs in cases:
        seeds = S.seed_triplet(cs["seed"], cs["flavour"])
        if cs.get("load") == "host":   # the twoPhaseMethod path: every rank loads its slab from the host arrays
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_py as O      # instance generation only (test infrastructure)
            A, b, c = O.generate(cs["n"], cs["m"], seeds, cs["lo"], cs["hi"])
            s.load(A, b, c)
        else:
            s.generate(cs["n"], cs["m"], seeds, cs["lo"], cs["hi"])
        if expect_peer_error:
            try:
                s.solve()
                out = {"case": cs, "error": None}
            except S.B2SError as exc:
                out = {"case": cs, "error": "peer" if "b2s error 7" in str(exc) else str(exc)[:200]}
            box = [None] * dist.get_world_size()
            dist.all_gather_object(box, out["error"])
            out["all_ranks"] = box
            if dist.get_rank() == 0:
                print("RESULT " + json.dumps(out), flush=True)
            continue
        r = s.solve()
        out = {"case": cs, "status": r["status"], "pivots": [r["stats"].pivots_phase1, r["stats"].pivots_phase2],
               "hash": str(r["stats"].trace_hash), "objective": r["objective"], "basis": r["basis"].tolist(),
               "x_nonzero": int((r["x"] != 0).sum()), "seconds": [r["stats"].seconds_phase1, r["stats"].seconds_phase2],
               "loop": s.loop_mode()}
        # every rank must hold the same replicated result
        box = [None] * dist.get_world_size()
        dist.all_gather_object(box, (out["status"], out["hash"], out["objective"]))
        assert all(b == box[0] for b in box), box
        if dist.get_rank() == 0:
            print("RESULT " + json.dumps(out), flush=True)
    s.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
