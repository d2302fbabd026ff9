"""CPU test of the drop-in surface: every function, macro and struct field the reference's public headers
declare is declared by include/compat too (same names), and every reference header file name exists as a
forwarder.  Needs the read-only reference tree, so it runs in the build container and is skipped elsewhere."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/include"
COMPAT = os.path.join(ROOT, "include", "compat")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference tree not present")


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def test_every_reference_header_has_a_forwarder():
    for name in os.listdir(REF_INC):
        assert os.path.exists(os.path.join(COMPAT, name)), f"include/compat/{name} missing"


def test_every_reference_declaration_is_covered():
    ours = strip_comments(open(os.path.join(COMPAT, "simplex_compat.hpp")).read())
    missing = []
    for name in sorted(os.listdir(REF_INC)):
        ref = strip_comments(open(os.path.join(REF_INC, name), encoding="utf-8-sig").read())
        # function declarations / inline definitions: an identifier followed by '(' at declaration level
        funcs = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*[;{]", ref))
        funcs -= {"if", "for", "while", "switch", "sizeof", "defined", "fprintf", "exit", "fopen", "abs", "HandleError",
                  "checkKernelError"} - {"HandleError", "checkKernelError"}
        macros = set(re.findall(r"#define\s+([A-Za-z_][A-Za-z0-9_]*)", ref))
        fields = set(re.findall(r"\b(?:TYPE\s*\*|int|size_t|problem_t\s*\*)\s*([A-Za-z_][A-Za-z0-9_]*)\s*;", ref))
        for ident in funcs | macros | fields:
            if not re.search(r"\b" + re.escape(ident) + r"\b", ours):
                missing.append(f"{name}: {ident}")
    assert not missing, missing


def test_status_codes_match():
    ref = open(os.path.join(REF_INC, "twoPhaseMethod.h")).read()
    ours = open(os.path.join(COMPAT, "simplex_compat.hpp")).read()
    for name in ("INFEASIBLE", "UNBOUNDED", "DEGENERATE", "FEASIBLE"):
        r = re.search(rf"#define\s+{name}\s+(-?\d+)", ref).group(1)
        o = re.search(rf"#define\s+{name}\s+(-?\d+)", ours).group(1)
        assert r == o
