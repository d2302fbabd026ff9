"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/b2s.h declares, the Python prototypes cover exactly that set, and the product fails
loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b2s.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from simplexoncuda_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/b2s.h but not exported"


def test_python_prototypes_match_header():
    from simplexoncuda_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    _lib.load()


def test_struct_layouts():
    from simplexoncuda_b200 import _lib
    # b2s_options: 7 ints, pad, 2 long long, 3 named + 5 reserved ints ; b2s_stats: 13 x 8 bytes
    assert ctypes.sizeof(_lib.Options) == 80
    assert ctypes.sizeof(_lib.Stats) == 104
    opt = _lib.Options()
    _lib.load().b2s_default_options(ctypes.byref(opt))
    assert (opt.dtype, opt.pivot_rule, opt.fold_artificials, opt.use_graph, opt.persistent) == (0, 0, 1, 1, 2)


def test_seed_triplets_match_oracle():
    import oracle_py as O
    from simplexoncuda_b200 import RAND_GLIBC, RAND_MSVC, seed_triplet
    for seed in (0, 1, 25856, 827392, 6619136, 4000000000):
        assert seed_triplet(seed, RAND_MSVC) == O.seed_triplet(seed, 1)
        assert seed_triplet(seed, RAND_GLIBC) == O.seed_triplet(seed, 0)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from simplexoncuda_b200 import B2SError, Solver
    with pytest.raises(B2SError, match="no CPU fallback"):
        Solver()


def test_text_problem_reader():
    import io
    import json
    from simplexoncuda_b200 import readProblemFromFile
    ex = json.load(open(os.path.join(ROOT, "tests", "golden", "examples.json")))
    p = readProblemFromFile(io.StringIO(ex["smallProblem"]["text"]))
    assert (p.vars, p.constraints) == (3, 2)
    assert p.constraintsMatrix.tolist() == [[1.0, 1.0], [3.0, 5.0], [2.0, 1.0]]  # variable-major
    assert p.knownTermsVector.tolist() == [10.0, 8.0] and p.objectiveFunction.tolist() == [8.0, 10.0, 7.0]


def test_header_is_plain_c(tmp_path):
    """include/b2s.h must be consumable from C (the drop-in boundary is a C ABI): compile it alone as C99."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "b2s.h"\nint main(void) { b2s_options o; (void)o; return (int)sizeof(b2s_stats) == 0; }\n')
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    out = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                          str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
