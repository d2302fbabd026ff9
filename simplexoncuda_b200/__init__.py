"""simplexoncuda_b200 -- B200-native dense-tableau two-phase simplex (drop-in for the solve path
of rik1599/SimplexOnCuda).  The product is the C-ABI library simplexoncuda_b200/lib/libb2s.so
(sources in simplexoncuda_b200/csrc, interface in include/b2s.h); this package is the thin host
layer used by the tests and the benchmark."""
from ._lib import (DEGENERATE, F32, F64, FEASIBLE, INFEASIBLE, ITER_LIMIT, RAND_GLIBC, RAND_MSVC,
                   RULE_BLAND, RULE_LOWEST, RULE_REFERENCE, RUNNING, STATUS_NAMES, UNBOUNDED, LIB_PATH)
from .solver import B2SError, Solver, dist_unique_id, seed_triplet
from .api import (Problem, generateRandomProblem, printProblemToStream, readProblemFromFile,
                  readRandomProblemFromFile, twoPhaseMethod)

__all__ = ["Solver", "B2SError", "seed_triplet", "dist_unique_id", "Problem", "generateRandomProblem",
           "readProblemFromFile", "readRandomProblemFromFile", "printProblemToStream", "twoPhaseMethod",
           "FEASIBLE", "INFEASIBLE", "UNBOUNDED", "DEGENERATE", "ITER_LIMIT", "RUNNING", "F64", "F32",
           "RULE_REFERENCE", "RULE_LOWEST", "RULE_BLAND", "RAND_GLIBC", "RAND_MSVC", "STATUS_NAMES", "LIB_PATH"]
