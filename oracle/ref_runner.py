"""oracle/ref_runner.py -- TEST INFRASTRUCTURE.  Runs the unmodified reference build
(oracle/_ref/libsimplex_ref.so, see oracle/build_ref.sh) on one LP in THIS process and writes the
result as JSON + npz.  Always launched as a fresh subprocess: the reference relies on freshly
allocated device memory being zero (its identity blocks are never zero-filled,
src/twoPhaseMethod.cu:30-39), so it must not share a process with other GPU work.

    python oracle/ref_runner.py <problem.npz> <out_prefix> [--trace]

problem.npz holds A (vars x constraints, variable-major), b, c.  Needs a GPU.
"""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libsimplex_ref.so")


def generate():
    """python oracle/ref_runner.py --generate <vars> <cons> <seed> <min> <max> <out.npz>: the reference's
    own generateRandomProblem (glibc rand() seeds, cuRAND XORWOW kernels)."""
    n, m, seed, lo, hi = (int(v) for v in sys.argv[2:7])
    lib = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    lib.ref_generate.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, dp, dp, dp]
    lib.ref_setup_device()
    A = np.zeros((n, m)); b = np.zeros(m); c = np.zeros(n)
    lib.ref_generate(n, m, seed, lo, hi, A.ctypes.data_as(dp), b.ctypes.data_as(dp), c.ctypes.data_as(dp))
    np.savez(sys.argv[7], A=A, b=b, c=c)


def main():
    if sys.argv[1] == "--generate":
        return generate()
    prob, out = sys.argv[1], sys.argv[2]
    trace_on = "--trace" in sys.argv[3:]
    z = np.load(prob)
    A = np.ascontiguousarray(z["A"], dtype=np.float64)
    b = np.ascontiguousarray(z["b"], dtype=np.float64)
    c = np.ascontiguousarray(z["c"], dtype=np.float64)
    n, m = A.shape
    lib = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    lib.ref_two_phase.restype = C.c_int
    lib.ref_two_phase.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, ip, C.c_int, ip, C.c_longlong,
                                  C.POINTER(C.c_longlong), dp]
    lib.ref_setup_device()
    x = np.zeros(n)
    obj = C.c_double(0.0)
    basis = np.full(m, -1, dtype=np.int32)
    cap = 1 << 22
    trace = np.zeros((cap if trace_on else 1, 2), dtype=np.int32)
    piv = (C.c_longlong * 2)()
    secs = (C.c_double * 3)()
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="refrun_")  # the reference prints to stdout only, but keep cwd clean
    os.chdir(scratch)
    try:
        status = lib.ref_two_phase(n, m, A.ctypes.data_as(dp), b.ctypes.data_as(dp), c.ctypes.data_as(dp),
                                   x.ctypes.data_as(dp), C.byref(obj), basis.ctypes.data_as(ip), int(trace_on),
                                   trace.ctypes.data_as(ip), cap, piv, secs)
    finally:
        os.chdir(cwd)
    total = piv[0] + piv[1]
    res = {"status": int(status), "objective": float(obj.value), "objective_repr": repr(float(obj.value)),
           "pivots_phase1": int(piv[0]), "pivots_phase2": int(piv[1]),
           "seconds_total": secs[0], "seconds_loop_phase1": secs[1], "seconds_loop_phase2": secs[2],
           "vars": n, "constraints": m}
    np.savez(out + ".npz", x=x, basis=basis, trace=trace[:min(total, cap)] if trace_on else trace[:0])
    with open(out + ".json", "w") as f:
        json.dump(res, f)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
