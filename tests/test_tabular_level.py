"""GPU test of the tabular_t level of the drop-in API (include/compat: newTabular, updateObjectiveFunction,
solve(tabular_t*, int*), minElement x2, isLessOrEqualThanZero) through a C++ client that uses it the way the
reference's twoPhaseMethod.cu does; results compared with the serial oracle bit for bit."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "simplexoncuda_b200", "lib", "test_tabular_level")


def write_lp(path, A, b, c):
    n, m = A.shape
    with open(path, "w") as f:
        f.write(f"{n} {m}\n" + " ".join(repr(float(v)) for v in c) + "\n")
        for i in range(m):
            f.write(" ".join(repr(float(A[j, i])) for j in range(n)) + " " + repr(float(b[i])) + "\n")


@pytest.mark.skipif(not os.path.exists(EXE), reason="C++ client not built")
@pytest.mark.parametrize("n,m,seed,lo", [(3, 2, 0, 0), (40, 24, 3, -100), (130, 70, 4, 1), (64, 600, 5, 1), (300, 200, 6, -100)])
def test_tabular_level_client(tmp_path, n, m, seed, lo):
    if seed == 0:
        A = np.array([[1.0, 1.0], [3.0, 5.0], [2.0, 1.0]]); b = np.array([10.0, 8.0]); c = np.array([8.0, 10.0, 7.0])
    else:
        A, b, c = O.generate(n, m, O.seed_triplet(seed, 0), lo, 100)
    lp = str(tmp_path / "lp.txt")
    write_lp(lp, A, b, c)
    out = subprocess.run([EXE, lp], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    res = json.loads(out.stdout.strip().splitlines()[-1])
    o = O.Oracle(A, b, c)
    o.build_phase1(); o.priceout()
    costs = o.costs()
    assert res["priced_cost0"] == costs[0]
    val, idx = O.tournament(costs[1:])
    assert (res["min_cost"], res["min_cost_index"]) == (val, idx)
    T = o.tableau()
    ratio = np.where(np.array([O.lib().orc_compare(float(v), 0.0) > 0 for v in T[1]]), T[0] / np.where(T[1] == 0, 1, T[1]),
                     np.finfo(np.float64).max)
    rv, ri = O.tournament(ratio)
    assert res["ratio_min"] == rv and (res["ratio_index"] == ri or ri < 0)
    assert res["col1_nonpositive"] == int(T[1].max() < 1e-9)
    st = o.iterate(-1)
    assert res["status"] == st
    assert res["cost0"] == o.costs()[0] and res["base"] == o.basis().tolist()


C_CLIENT = os.path.join(ROOT, "simplexoncuda_b200", "lib", "example_solve_file")


@pytest.mark.skipif(not os.path.exists(C_CLIENT), reason="C client not built")
def test_plain_c_client(tmp_path):
    """examples/solve_file.c (C99, gcc) through the C ABI on the reference's smallProblem: optimum 64 at x=(8,0,0)."""
    lp = tmp_path / "small.txt"
    lp.write_text("3 2\n8 10 7\n1 3 2 10\n1 5 1 8\n")
    out = subprocess.run([C_CLIENT, str(lp)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith("status 0  pivots 2+2")
    assert lines[1] == "optimal value 64" and lines[2] == "x = 8 0 0" and lines[3] == "basis = 3 0"
