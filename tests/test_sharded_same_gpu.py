"""Sharded solves on ONE GPU: world_size ranks as separate processes, all on cuda:0, bootstrapped through the host layer
(b2s_dist_init_host over torch.distributed/gloo, CUDA-IPC arenas).  The peer-memory kernels -- look-ahead kernel, the four-launch
exchanges and the persistent loop kernel -- run exactly the code they run over NVLink; the GPU time-slices the ranks' kernels,
so every flag wait really crosses a process boundary.  Results must equal the oracle's bit for bit.  Also: a rank that stops
publishing must surface as B2S_ERR_PEER on every rank instead of a wrong answer."""
import json
import os
import subprocess
import sys

import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    return port


def _run(world, cases, env_extra, port):
    env = dict(os.environ, B2S_TEST_SAME_GPU="1", B2S_PEER_TIMEOUT_MS="30000")
    env.update(env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "sharded_worker.py"),
           json.dumps(cases)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return [json.loads(l[len("RESULT "):]) for l in out.stdout.splitlines() if l.startswith("RESULT ")]


@pytest.mark.parametrize("mode", ["p2p-lookahead", "p2p-launches", "p2p-persistent"])
def test_two_ranks_on_one_gpu_match_oracle(mode):
    env = {}
    if mode == "p2p-launches":
        env["B2S_LOOKAHEAD"] = "0"
    if mode == "p2p-persistent":
        env["B2S_TEST_PERSISTENT"] = "1"
    cases = [dict(n=40, m=1024, seed=11, flavour=1, lo=1, hi=100),
             dict(n=64, m=1024, seed=5, flavour=0, lo=-100, hi=100, load="host")]
    results = _run(2, cases, env, 29721)
    assert len(results) == len(cases)
    for res in results:
        cs = res["case"]
        A, b, c = O.generate(cs["n"], cs["m"], O.seed_triplet(cs["seed"], cs["flavour"]), cs["lo"], cs["hi"])
        ref = O.Oracle(A, b, c, threads=4).two_phase()
        assert res["status"] == ref["status"]
        assert tuple(res["pivots"]) == tuple(ref["pivots"])
        assert res["hash"] == str(ref["hash"])
        assert res["basis"] == ref["basis"].tolist()
        if ref["status"] == 0:
            assert res["objective"] == ref["objective"]
        if mode == "p2p-lookahead":
            assert "look-ahead" in res["loop"]


def test_four_ranks_on_one_gpu_lookahead():
    cases = [dict(n=32, m=2048, seed=7, flavour=1, lo=1, hi=100)]
    res = _run(4, cases, {}, 29722)[0]
    cs = res["case"]
    A, b, c = O.generate(cs["n"], cs["m"], O.seed_triplet(cs["seed"], cs["flavour"]), cs["lo"], cs["hi"])
    ref = O.Oracle(A, b, c, threads=4).two_phase()
    assert (res["status"], tuple(res["pivots"]), res["hash"]) == (ref["status"], tuple(ref["pivots"]), str(ref["hash"]))
    assert res["objective"] == ref["objective"]


@pytest.mark.parametrize("mode", ["p2p-lookahead", "p2p-launches"])
def test_silent_peer_is_an_error_not_a_wrong_answer(mode):
    """Rank 1 stops publishing its ratio-test winners from pivot 6 on: every rank's bounded wait expires and
    b2s_solve_two_phase returns B2S_ERR_PEER (it must not run on into the phase-1 verdict or phase 2)."""
    env = {"B2S_FAULT_RANK": "1", "B2S_FAULT_PIVOT": "6", "B2S_PEER_TIMEOUT_MS": "3000", "B2S_EXPECT_PEER_ERROR": "1"}
    if mode == "p2p-launches":
        env["B2S_LOOKAHEAD"] = "0"
    res = _run(2, [dict(n=40, m=1024, seed=11, flavour=1, lo=1, hi=100)], env, 29723)[0]
    assert res["all_ranks"] == ["peer", "peer"], res
