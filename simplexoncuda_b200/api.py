"""Host-side mirror of the reference's public interface for the solve path, same names and
argument meaning (include/problem.h:37-73, include/twoPhaseMethod.h:5-19), over the C ABI.
The C++ drop-in with the same names lives in include/compat/ + csrc/compat.cu; this module is
what the Python tests and bench.py call."""
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .solver import Solver, seed_triplet


@dataclass
class Problem:
    """problem_t (include/problem.h:10-26): max c.x  s.t.  A x <= b, x >= 0."""
    constraintsMatrix: np.ndarray   # [vars, constraints], variable-major: A[j, i]
    knownTermsVector: np.ndarray    # b[constraints]
    objectiveFunction: np.ndarray   # c[vars]

    @property
    def vars(self):
        return self.constraintsMatrix.shape[0]

    @property
    def constraints(self):
        return self.constraintsMatrix.shape[1]


def readProblemFromFile(file):
    """Text LP: `n m` / c[n] / m lines `a_i1 .. a_in b_i` (src/problem.cu:20-47)."""
    tok = file.read().split()
    n, m = int(tok[0]), int(tok[1])
    vals = np.array(tok[2:2 + n + m * (n + 1)], dtype=np.float64)
    if vals.size != n + m * (n + 1):
        raise ValueError("truncated problem file")
    c = vals[:n].copy()
    body = vals[n:].reshape(m, n + 1)
    return Problem(np.ascontiguousarray(body[:, :n].T), body[:, n].copy(), c)


def generateRandomProblem(nVars, nConstraints, seed, minGenerator=-100, maxGenerator=100,
                          rand_flavour=L.RAND_GLIBC, device=0):
    """src/problem.cu:49-126: the instance is generated on the device and copied to the host."""
    seeds = seed_triplet(seed, rand_flavour)
    with Solver(device=device) as s:
        s.generate(nVars, nConstraints, seeds, minGenerator, maxGenerator)
        A, b, c = s.copy_problem()
    return Problem(A, b, c)


def readRandomProblemFromFile(file, rand_flavour=L.RAND_GLIBC, device=0):
    """Seed file `vars constraints seed min max` (src/problem.cu:128-139)."""
    tok = file.read().split()
    n, m, seed, lo, hi = int(tok[0]), int(tok[1]), int(tok[2]), int(tok[3]), int(tok[4])
    return generateRandomProblem(n, m, seed, lo, hi, rand_flavour=rand_flavour, device=device)


def printProblemToStream(stream, problem):
    """src/problem.cu:141-181."""
    c, A, b = problem.objectiveFunction, problem.constraintsMatrix, problem.knownTermsVector
    stream.write("max " + " ".join(f"{'+' if v >= 0 else '-'} {abs(v):.2f} X{i + 1}" for i, v in enumerate(c)) + " \n")
    stream.write("subject to \n")
    for i in range(problem.constraints):
        stream.write(" ".join(f"{'+' if A[j, i] >= 0 else '-'} {abs(A[j, i]):.2f} X{j + 1}"
                              for j in range(problem.vars)) + f" <= {b[i]:.2f}\n")


def twoPhaseMethod(problem, device=0, **options):
    """include/twoPhaseMethod.h:10-19.  Returns (status, solution, optimalValue); solution and
    optimalValue are None unless status == FEASIBLE (the reference leaves them unwritten)."""
    with Solver(device=device, **options) as s:
        s.load(problem.constraintsMatrix, problem.knownTermsVector, problem.objectiveFunction)
        r = s.solve()
    if r["status"] == L.FEASIBLE:
        return r["status"], r["x"], r["objective"]
    return r["status"], None, None
