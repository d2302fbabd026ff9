"""CPU test of the N>1 path (world_size 2, gloo): the sharded per-pivot protocol of DESIGN.md section 5
-- constraint slabs, all-gather of ratio-test stage-1 block winners, stage 2 replayed on every rank,
pivot constraint published by the owner through an integer-sum all-reduce of bit patterns, price-out
chained through the ranks -- executed with torch.distributed over gloo and numpy stand-ins for the
kernels.  The pivot sequence, basis and objective must equal the unsharded serial oracle's bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIG = np.finfo(np.float64).max


def _cmp3(x, y=0.0):
    if abs(x - y) < 1e-9:
        return 0
    return -1 if x < y else 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, m, seeds, lo, hi, max_pivots, q):
    import ctypes as C
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from simplexoncuda_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = O.lib()
    lib.orc_stage1_block.restype = C.c_double
    lib.orc_stage1_block.argtypes = [C.POINTER(C.c_double), C.c_long, C.c_long, C.c_long, C.POINTER(C.c_int)]
    lib.orc_stage2.restype = C.c_double
    lib.orc_stage2.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_long, C.POINTER(C.c_int)]
    dp = C.POINTER(C.c_double)

    # unique-id plumbing (the id itself is opaque bytes here)
    uid = sharding.exchange_unique_id(dist, lambda: bytes(range(128)))
    assert uid == bytes(range(128))

    A, b, c = O.generate(n, m, seeds, lo, hi)
    c0, c1 = sharding.slab(rank, world, m)
    w = c1 - c0
    G = m // 512
    R = 1 + n + 2 * m
    # ---- build (local slab, reference layout) ----
    T = np.zeros((R, w))
    T[0] = b[c0:c1]
    T[1:1 + n] = A[:, c0:c1]
    for li in range(w):
        gi = c0 + li
        T[1 + n + gi, li] = 1.0
        T[1 + n + m + gi, li] = 1.0
        if _cmp3(T[0, li]) < 0:
            T[:, li] = -T[:, li]
    cost = np.zeros(R)
    cost[n + m + 1:] = 1.0
    base = n + m + np.arange(m)

    def priceout(Rc):
        coef = cost[1 + base[c0:c1]].copy()
        for r in range(world):                      # running sums chained through the ranks
            if r == rank:
                for y in range(Rc):
                    acc = cost[y]
                    row = T[y]
                    for k0 in range(0, w, 64):
                        for l in range(32):
                            x = k0 + l
                            if x >= w:
                                break
                            if x + 32 < w:
                                term = float(np.float64(row[x + 32]) * np.float64(coef[x + 32]))
                                term = _fma(row[x], coef[x], term)
                            else:
                                term = float(row[x] * coef[x])
                            acc = acc + (-term)
                    cost[y] = acc
            t = torch.from_numpy(cost[:Rc].copy())
            dist.broadcast(t, src=r)
            cost[:Rc] = t.numpy()

    import math
    _fma = math.fma if hasattr(math, "fma") else None
    if _fma is None:                                # Python < 3.13: exact fma through the oracle's libm
        libm = C.CDLL("libm.so.6")
        libm.fma.restype = C.c_double
        libm.fma.argtypes = [C.c_double] * 3
        _fma = libm.fma

    def pivot(Rc):
        idx = C.c_int(-1)
        costs = np.ascontiguousarray(cost[1:Rc])
        cq = lib.orc_tournament(costs.ctypes.data_as(dp), costs.size, C.byref(idx))
        q = idx.value
        if not (_cmp3(cq) < 0) or q < 0:
            return "optimal"
        col = T[1 + q].copy()
        ratio = np.where(np.array([_cmp3(v) > 0 for v in col]), T[0] / np.where(col == 0, 1.0, col), BIG)
        # stage 1 on the local blocks (global block ids), then all-gather
        sv = np.full(G, BIG); si = np.full(G, -1, dtype=np.int32); smax = np.full(G, np.finfo(np.float64).tiny)
        full = np.full(m, BIG)
        full[c0:c1] = ratio                         # only the local part is ever read for local blocks
        for gb in sharding.stage1_blocks(rank, world, m):
            v = lib.orc_stage1_block(full.ctypes.data_as(dp), m, G, gb, C.byref(idx))
            sv[gb] = v; si[gb] = idx.value
            smax[gb] = max(np.finfo(np.float64).tiny, col[gb * 512 - c0:(gb + 1) * 512 - c0].max())
        for arr in (sv, smax, si):
            parts = [torch.zeros(G // world, dtype=torch.from_numpy(arr).dtype) for _ in range(world)]
            mine = torch.from_numpy(arr[c0 // 512:c1 // 512].copy())
            dist.all_gather(parts, mine)
            arr[:] = torch.cat(parts).numpy()
        if _cmp3(smax.max()) <= 0:
            return "unbounded"
        lib.orc_stage2(sv.ctypes.data_as(dp), si.ctypes.data_as(C.POINTER(C.c_int)), G, C.byref(idx))
        p = idx.value
        base[p] = q
        # owner publishes the raw pivot constraint as bit patterns; others add zeros
        owner = sharding.owner_of(p, world, m)
        bits = np.zeros(Rc, dtype=np.int64)
        if owner == rank:
            lp = p - c0
            rowp_local = T[:Rc, lp].copy()
            bits[:] = rowp_local.view(np.int64)
            T[:Rc, lp] = rowp_local / col[lp]
        t = torch.from_numpy(bits)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        rowp = t.numpy().view(np.float64).copy()
        piv = rowp[1 + q]
        s = (-col) / piv
        if owner == rank:
            s[p - c0] = 0.0
        sc = (-cq) / piv
        for r in range(Rc):
            a = rowp[r]
            if a != 0.0:
                T[r] = np.array([_fma(s[i], a, T[r, i]) for i in range(w)])
            cost[r] = _fma(sc, rowp[r], cost[r])
        trace.append((q, p))
        return "continue"

    trace = []
    priceout(R)
    st = "continue"
    while st == "continue" and len(trace) < max_pivots:
        st = pivot(R)
    result = {"phase1": st, "trace": list(trace)}
    infeasible = _cmp3(cost[0]) < 0
    degenerate = bool(np.any((base >= n + m) & (base < n + 2 * m)))
    if st != "continue" and not infeasible and not degenerate:
        R2 = 1 + n + m
        cost[1 + n:1 + n + m] = 0.0
        cost[1:1 + n] = -c
        priceout(R2)
        st = "continue"
        while st == "continue" and len(trace) < max_pivots:
            st = pivot(R2)
        result["phase2"] = st
    result["trace"] = list(trace)
    result["base"] = base.tolist()
    result["objective"] = float(cost[0])
    result["infeasible"] = bool(infeasible)
    if rank == 0:
        q.put(result)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,m,seed,lo", [(6, 1024, 3, 1), (5, 1024, 8, -100)])
def test_sharded_protocol_matches_unsharded_oracle(n, m, seed, lo):
    seeds = O.seed_triplet(seed, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    max_pivots = 60
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, m, seeds, lo, 100, max_pivots, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    A, b, c = O.generate(n, m, seeds, lo, 100)
    ref = O.Oracle(A, b, c).two_phase(max_pivots=max_pivots)
    k = len(res["trace"])
    assert k > 0
    assert [tuple(t) for t in ref["trace"][:k].tolist()] == [tuple(t) for t in res["trace"]]
    if ref["status"] != O.ITER_LIMIT:
        assert len(ref["trace"]) == k
        assert res["base"] == ref["basis"].tolist()
        if ref["status"] == 0:
            assert res["objective"] == ref["objective"]
        if ref["status"] == O.INFEASIBLE:
            assert res["infeasible"]


def test_slab_arithmetic():
    from simplexoncuda_b200 import sharding
    assert sharding.slab(0, 2, 2048) == (0, 1024) and sharding.slab(1, 2, 2048) == (1024, 2048)
    assert [sharding.owner_of(p, 4, 8192) for p in (0, 2047, 2048, 8191)] == [0, 0, 1, 3]
    assert list(sharding.stage1_blocks(3, 8, 65536)) == list(range(48, 64))
    with pytest.raises(ValueError):
        sharding.slab(0, 3, 2048)
    with pytest.raises(ValueError):
        sharding.slab(2, 2, 2048)


# ---- host-layer bootstrap (b2s_dist_init_host): the all-gather the library calls back into ------------------------------
def _ag_worker(rank, world, port, q):
    import ctypes as C
    sys.path.insert(0, ROOT)
    from simplexoncuda_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ag = sharding.make_host_allgather(dist)
    ok = True
    for nbytes in (4, 64, 8 * 1000 + 3):          # a barrier word, a CUDA-IPC handle, a cost vector with an odd length
        send = (C.c_ubyte * nbytes)(*[(rank * 37 + i) & 0xFF for i in range(nbytes)])
        recv = (C.c_ubyte * (nbytes * world))()
        rc = ag(C.addressof(send), C.addressof(recv), nbytes)
        got = bytes(recv)
        want = b"".join(bytes((r * 37 + i) & 0xFF for i in range(nbytes)) for r in range(world))
        ok = ok and rc == 0 and got == want
    q.put((rank, ok))
    dist.destroy_process_group()


def test_host_allgather_callback_world2():
    """The b2s_allgather_fn that sharding.init_sharded_solver_host hands to the library: every rank's bytes, in rank order,
    on every rank (gloo, CPU).  The GPU side of the same bootstrap is tests/test_sharded_same_gpu.py."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ag_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
