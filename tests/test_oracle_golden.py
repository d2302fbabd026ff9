"""CPU tests: the serial oracle against every golden vector the reference ships for this path
(published per-phase pivot counts, statuses of the three example LPs) and against the committed
oracle fixtures.  Runs without a GPU."""
import io
import json
import os
import random

import numpy as np
import pytest

import oracle_py as O

HERE = os.path.dirname(os.path.abspath(__file__))
PUB = json.load(open(os.path.join(HERE, "golden", "published_pivot_counts.json")))
EXAMPLES = json.load(open(os.path.join(HERE, "golden", "examples.json")))
ORC = json.load(open(os.path.join(HERE, "golden", "oracle_results.json")))


def parse_lp(text):
    tok = text.split()
    n, m = int(tok[0]), int(tok[1])
    v = np.array(tok[2:], dtype=np.float64)
    c = v[:n]
    body = v[n:n + m * (n + 1)].reshape(m, n + 1)
    return np.ascontiguousarray(body[:, :n].T), body[:, n].copy(), c.copy()


def test_compare_semantics():
    # include/macro.h:28-42
    l = O.lib()
    assert l.orc_compare(0.0, 0.0) == 0
    assert l.orc_compare(5e-10, 0.0) == 0
    assert l.orc_compare(-5e-10, 0.0) == 0
    assert l.orc_compare(1e-9, 0.0) == 1
    assert l.orc_compare(-1e-9, 0.0) == -1
    assert l.orc_compare(float("nan"), 0.0) == 1
    assert l.orc_compare(1.7976931348623157e308, 1.7976931348623157e308) == 0


def test_seed_triplets_match_survey():
    assert O.seed_triplet(25856, 1) == (18937, 13107, 19527)
    assert O.seed_triplet(25856, 0) == (1600737720, 2071352359, 1594764739)
    assert O.seed_triplet(103424, 1) == (10097, 29269, 14395)
    assert O.seed_triplet(827392, 1) == (14976, 16281, 32033)
    assert O.seed_triplet(827392, 0) == (2120439432, 59594510, 2045403306)


def test_glibc_rand_matches_libc():
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 2, 25856, 413696, 4000000000):
        libc.srand(ctypes.c_uint(seed))
        want = tuple(libc.rand() for _ in range(3))
        assert O.seed_triplet(seed, 0) == want


def test_generator_first_values():
    # first values of random_256_256 with MSVC seeds
    A, b, c = O.generate(256, 256, (18937, 13107, 19527), 1, 100)
    assert b[0] == 93.829211235046387
    assert c[0] == 14.719509437680244
    assert A[0, 0] == 61.735219419002533
    assert A[1, 0] == 4.7217921577394009


def test_tournament_tie_order_is_bit_reversed():
    # all-equal vector of 1024 entries: the winner is index 0; knocking out winners in turn must
    # follow the bit-reversed order the survey documents (256, 272, 264, 280 for a block offset 256).
    v = np.zeros(1024)
    v[256:768] = -1.0
    order = []
    for _ in range(4):
        val, idx = O.tournament(v)
        order.append(idx)
        v[idx] = 0.0
    assert order == [256, 272, 264, 280]


def test_tournament_single_block_and_empty():
    val, idx = O.tournament(np.array([3.0, 1.0, 2.0]))
    assert (val, idx) == (1.0, 1)
    val, idx = O.tournament(np.full(7, 1.7976931348623157e308))
    assert idx == -1


@pytest.mark.parametrize("name", ["smallProblem", "infeasibleProblem", "unboundedProblem"])
def test_example_statuses(name):
    ex = EXAMPLES[name]
    A, b, c = parse_lp(ex["text"])
    r = O.Oracle(A, b, c).two_phase()
    assert r["status"] == ex["status"]
    if ex["status"] == 0:
        assert r["objective"] == ex["objective"]
        assert list(r["x"]) == ex["x"]
        assert r["trace"].tolist() == [[1, 1], [2, 0], [0, 1], [3, 0]]


def _published(max_cons, max_vars):
    seen = set()
    for inst in PUB["instances"]:
        key = (inst["vars"], inst["constraints"], inst["seed"])
        if inst["constraints"] <= max_cons and inst["vars"] <= max_vars and key not in seen:
            seen.add(key)
            yield inst


@pytest.mark.parametrize("inst", list(_published(512, 2048)), ids=lambda i: f"{i['vars']}x{i['constraints']}")
def test_published_pivot_counts(inst):
    """Golden vectors of the reference: data/measures/*/benchmark_<n>_<m>.txt (MSVC seed flavour)."""
    n, m, seed = inst["vars"], inst["constraints"], inst["seed"]
    A, b, c = O.generate(n, m, O.seed_triplet(seed, 1), 1, 100)
    r = O.Oracle(A, b, c, threads=4).two_phase()
    assert r["status"] == 0
    assert r["pivots"] == (inst["pivots_phase1"], inst["pivots_phase2"])
    fx = ORC[f"{n}_{m}_{seed}"]
    assert str(r["hash"]) == fx["trace_hash"]
    assert r["objective"] == fx["objective"]


def test_oracle_fixture_agrees_with_published_counts():
    """Every instance the oracle fixture covers reproduces the published counts (incl. sizes too
    slow for the default CPU suite, generated offline by tests/golden/make_oracle_results.py)."""
    checked = 0
    for inst in PUB["instances"]:
        key = f"{inst['vars']}_{inst['constraints']}_{inst['seed']}"
        if key in ORC:
            assert ORC[key]["pivots_phase1"] == inst["pivots_phase1"], key
            if inst["phase2_ran"]:
                assert ORC[key]["pivots_phase2"] == inst["pivots_phase2"], key
            checked += 1
    assert checked >= 37   # all 36 published instances + the MX250 run of 1024x8192 with the un-bumped seed


def test_stepping_equals_whole_solve():
    A, b, c = O.generate(40, 24, (11, 22, 33), -100, 100)
    whole = O.Oracle(A, b, c).two_phase()
    o = O.Oracle(A, b, c)
    o.build_phase1(); o.priceout()
    while o.pivot() == O.CONTINUE:
        pass
    st = o.phase1_verdict()
    if st == 0:
        o.switch_phase2(); o.priceout()
        st = o.iterate(-1)
    assert st == whole["status"] and o.hash() == whole["hash"]


def test_mixed_sign_suite_terminates_with_all_statuses():
    """min=-100,max=100 instances (the reference's -r range, main.cu:7-8): mostly UNBOUNDED or
    INFEASIBLE; every one terminates."""
    seen = set()
    for n, m in ((48, 32), (32, 48)):
        for seed in range(1, 31):
            A, b, c = O.generate(n, m, O.seed_triplet(seed, 0), -100, 100)
            r = O.Oracle(A, b, c).two_phase(max_pivots=100000)
            assert r["status"] != O.ITER_LIMIT
            seen.add(r["status"])
    assert {O.INFEASIBLE, O.UNBOUNDED, O.FEASIBLE} <= seen


def small_integer_lps(count, seed=0):
    """Tiny integer LPs with negative and zero RHS entries: the suite that exercises negated
    constraints, phase-1 artificials and all four reference statuses (incl. DEGENERATE)."""
    rng = np.random.default_rng(seed)
    for _ in range(count):
        n = int(rng.integers(2, 5)); m = int(rng.integers(2, 5))
        yield (rng.integers(-3, 4, size=(n, m)).astype(float), rng.integers(-3, 4, size=m).astype(float),
               rng.integers(-3, 4, size=n).astype(float))


def test_small_integer_suite_covers_all_statuses():
    seen = set()
    for A, b, c in small_integer_lps(2000):
        seen.add(O.Oracle(A, b, c).two_phase(max_pivots=1000)["status"])
    assert {O.FEASIBLE, O.INFEASIBLE, O.UNBOUNDED, O.DEGENERATE} <= seen


def test_cycling_fixtures():
    cases = json.load(open(os.path.join(HERE, "golden", "cycling.json")))["cases"]
    for cs in cases:
        A, b, c = np.array(cs["A"], float), np.array(cs["b"], float), np.array(cs["c"], float)
        assert O.Oracle(A, b, c, rule=0).two_phase(max_pivots=5000)["status"] == O.ITER_LIMIT
        r = O.Oracle(A, b, c, rule=2).two_phase(max_pivots=5000)
        assert r["status"] == cs["bland_status"] and list(r["pivots"]) == cs["bland_pivots"]


def test_lowest_index_and_bland_rules_reach_same_optimum():
    A, b, c = O.generate(64, 64, O.seed_triplet(7, 1), 1, 100)
    ref = O.Oracle(A, b, c, rule=0).two_phase()
    low = O.Oracle(A, b, c, rule=1).two_phase()
    bland = O.Oracle(A, b, c, rule=2).two_phase(max_pivots=200000)
    assert ref["status"] == low["status"] == bland["status"] == 0
    assert abs(ref["objective"] - low["objective"]) <= 1e-7 * abs(ref["objective"])
    assert abs(ref["objective"] - bland["objective"]) <= 1e-7 * abs(ref["objective"])


def test_relative_infeasibility_mode_oracle():
    """Opt-in mode beyond the reference: the absolute test mis-declares scaled feasible LPs infeasible."""
    A, b, c = O.generate(24, 16, O.seed_triplet(1, 0), 1, 100)
    base = O.Oracle(A, b, c).two_phase()
    assert base["status"] == 0
    assert O.Oracle(A, b * 1e7, c).two_phase(max_pivots=5000)["status"] == O.INFEASIBLE
    rel = O.Oracle(A, b * 1e7, c, relative_infeasibility=True).two_phase(max_pivots=5000)
    assert rel["status"] == 0 and abs(rel["objective"] - base["objective"] * 1e7) <= 1e-9 * abs(rel["objective"])
    # a genuinely infeasible LP stays infeasible in both modes
    p = json.load(open(os.path.join(HERE, "golden", "examples.json")))["infeasibleProblem"]["text"]
    Ai, bi, ci = parse_lp(p)
    assert O.Oracle(Ai, bi, ci, relative_infeasibility=True).two_phase()["status"] == O.INFEASIBLE


def test_oracle_drive_out_mode_matches_highs_fixture():
    """The oracle's opt-in drive-out mode (beyond the reference) on the committed DEGENERATE fixtures: default mode keeps the
    reference verdict, drive-out mode reproduces the committed pivots and the optimum HiGHS found when the fixture was made."""
    import json
    import os
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "degenerate.json")))["cases"]
    for cs in cases:
        A, b, c = np.array(cs["A"], float), np.array(cs["b"], float), np.array(cs["c"], float)
        assert O.Oracle(A, b, c).two_phase(max_pivots=2000)["status"] == -3
        r = O.Oracle(A, b, c, drive_out=True).two_phase(max_pivots=2000)
        assert r["status"] == cs["drive_out_status"] and str(r["hash"]) == cs["trace_hash"] and list(r["pivots"]) == cs["pivots"]
        if r["status"] == 0:
            assert abs(r["objective"] - cs["highs_objective"]) <= 1e-7 * max(1.0, abs(cs["highs_objective"]))
