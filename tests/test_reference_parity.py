"""GPU parity against the UNMODIFIED reference build (oracle/_ref/libsimplex_ref.so, compiled from
/root/reference by oracle/build_ref.sh), each reference solve in a fresh subprocess: identical
status, per-phase pivot counts, pivot sequence and basis; objective within 1e-9 relative (the
reference's price-out accumulates with fp64 atomics in unspecified order)."""
import io
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libsimplex_ref.so")
RUNNER = os.path.join(ROOT, "oracle", "ref_runner.py")
EXAMPLES = json.load(open(os.path.join(ROOT, "tests", "golden", "examples.json")))

needs_ref = pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (needs /root/reference at build time)")


def run_reference(tmp_path, A, b, c, trace=True):
    prob = str(tmp_path / "prob.npz")
    out = str(tmp_path / "ref")
    np.savez(prob, A=A, b=b, c=c)
    cmd = [sys.executable, RUNNER, prob, out] + (["--trace"] if trace else [])
    subprocess.run(cmd, check=True, capture_output=True, timeout=900)
    res = json.load(open(out + ".json"))
    z = np.load(out + ".npz")
    res.update(x=z["x"], basis=z["basis"], trace=z["trace"])
    return res


def compare(S, tmp_path, A, b, c, **opts):
    ref = run_reference(tmp_path, A, b, c)
    with S.Solver(**opts) as s:
        s.load(A, b, c)
        r = s.solve()
        qp, cnt, h = s.trace()
    assert r["status"] == ref["status"]
    assert (r["stats"].pivots_phase1, r["stats"].pivots_phase2) == (ref["pivots_phase1"], ref["pivots_phase2"])
    assert np.array_equal(qp, ref["trace"])
    if ref["pivots_phase1"] + ref["pivots_phase2"] > 0:
        assert np.array_equal(r["basis"], ref["basis"])
    if r["status"] == 0:
        assert abs(r["objective"] - ref["objective"]) <= 1e-9 * max(1.0, abs(ref["objective"]))
        assert np.allclose(r["x"], ref["x"], rtol=1e-9, atol=1e-12)
    return r, ref


@pytest.fixture(scope="module")
def S():
    import simplexoncuda_b200 as S
    return S


@needs_ref
@pytest.mark.parametrize("name", ["smallProblem", "infeasibleProblem", "unboundedProblem"])
def test_examples_vs_reference(S, tmp_path, name):
    p = S.readProblemFromFile(io.StringIO(EXAMPLES[name]["text"]))
    r, ref = compare(S, tmp_path, p.constraintsMatrix, p.knownTermsVector, p.objectiveFunction)
    assert ref["status"] == EXAMPLES[name]["status"]


@needs_ref
@pytest.mark.parametrize("n,m,seed,flavour,lo", [(256, 256, 25856, 0, 1), (256, 256, 25856, 1, 1), (1024, 256, 102656, 0, 1),
                                                  (256, 512, 26112, 1, 1), (1024, 1024, 103424, 1, 1),
                                                  (48, 32, 3, 0, -100), (32, 48, 4, 0, -100), (200, 100, 5, 0, -100),
                                                  (513, 1025, 6, 0, -100)])
def test_random_vs_reference(S, tmp_path, n, m, seed, flavour, lo):
    A, b, c = O.generate(n, m, O.seed_triplet(seed, flavour), lo, 100)
    compare(S, tmp_path, A, b, c)


@needs_ref
def test_unfolded_layout_vs_reference(S, tmp_path):
    A, b, c = O.generate(300, 200, O.seed_triplet(42, 0), 1, 100)
    compare(S, tmp_path, A, b, c, fold_artificials=False)


@needs_ref
@pytest.mark.parametrize("n,m,seed,lo,hi", [(256, 256, 25856, 1, 100), (100, 300, 7, -100, 100)])
def test_generator_vs_reference(S, tmp_path, n, m, seed, lo, hi):
    """Our device generator (glibc seed flavour) == the reference's generateRandomProblem on this box."""
    out = str(tmp_path / "gen.npz")
    subprocess.run([sys.executable, RUNNER, "--generate", str(n), str(m), str(seed), str(lo), str(hi), out],
                   check=True, capture_output=True, timeout=300)
    z = np.load(out)
    p = S.generateRandomProblem(n, m, seed, lo, hi, rand_flavour=S.RAND_GLIBC)
    assert np.array_equal(p.constraintsMatrix, z["A"])
    assert np.array_equal(p.knownTermsVector, z["b"]) and np.array_equal(p.objectiveFunction, z["c"])
