"""ctypes binding of the serial oracle (oracle/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None

FEASIBLE, INFEASIBLE, UNBOUNDED, DEGENERATE, ITER_LIMIT, CONTINUE = 0, -1, -2, -3, -4, -10


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "serial_tableau.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
        l = C.CDLL(LIB)
        l.orc_create.restype = C.c_void_p
        l.orc_create.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int]
        l.orc_destroy.argtypes = [C.c_void_p]
        l.orc_set_trace.argtypes = [C.c_void_p, _ip, C.c_long]
        for name in ("orc_build_phase1", "orc_priceout", "orc_switch_phase2"):
            getattr(l, name).argtypes = [C.c_void_p]
            getattr(l, name).restype = None
        l.orc_pivot.argtypes = [C.c_void_p]
        l.orc_iterate.argtypes = [C.c_void_p, C.c_long]
        l.orc_phase1_verdict.argtypes = [C.c_void_p]
        l.orc_extract.argtypes = [C.c_void_p, _dp, _dp]
        l.orc_two_phase.argtypes = [C.c_void_p, C.c_long, _dp, _dp]
        l.orc_set_relative_infeasibility.argtypes = [C.c_void_p, C.c_int]
        l.orc_set_drive_out.argtypes = [C.c_void_p, C.c_int]
        l.orc_drive_out_artificials.argtypes = [C.c_void_p]
        l.orc_drive_out_artificials.restype = C.c_long
        l.orc_rows.restype = C.c_long
        l.orc_rows.argtypes = [C.c_void_p]
        l.orc_pivots.restype = C.c_long
        l.orc_pivots.argtypes = [C.c_void_p, C.c_int]
        l.orc_trace_len.restype = C.c_long
        l.orc_trace_len.argtypes = [C.c_void_p]
        l.orc_hash.restype = C.c_uint64
        l.orc_hash.argtypes = [C.c_void_p]
        l.orc_tableau.restype = _dp
        l.orc_tableau.argtypes = [C.c_void_p]
        l.orc_costs.restype = _dp
        l.orc_costs.argtypes = [C.c_void_p]
        l.orc_basis.restype = _ip
        l.orc_basis.argtypes = [C.c_void_p]
        l.orc_tournament.restype = C.c_double
        l.orc_tournament.argtypes = [_dp, C.c_long, _ip]
        l.orc_compare.argtypes = [C.c_double, C.c_double]
        l.orc_seed_triplet.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_uint)]
        l.orc_generate.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint), C.c_double, C.c_double, _dp, _dp, _dp]
        l.orc_xorwow_outputs.argtypes = [C.c_uint64, C.c_uint64, C.c_long, C.POINTER(C.c_uint32)]
        _lib = l
    return _lib


def seed_triplet(seed, flavour):
    out = (C.c_uint * 3)()
    lib().orc_seed_triplet(seed & 0xFFFFFFFF, flavour, out)
    return tuple(int(v) for v in out)


def generate(n, m, seeds, lo, hi):
    A = np.empty((n, m)); b = np.empty(m); c = np.empty(n)
    arr = (C.c_uint * 3)(*seeds)
    lib().orc_generate(n, m, arr, float(lo), float(hi), A.ctypes.data_as(_dp), b.ctypes.data_as(_dp), c.ctypes.data_as(_dp))
    return A, b, c


def tournament(vec):
    v = np.ascontiguousarray(vec, dtype=np.float64)
    idx = C.c_int(-2)
    val = lib().orc_tournament(v.ctypes.data_as(_dp), v.size, C.byref(idx))
    return val, idx.value


def xorwow_outputs(seed, offset, count):
    out = np.zeros(count, dtype=np.uint32)
    lib().orc_xorwow_outputs(seed, offset, count, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


class Oracle:
    """Stepping handle over the serial restatement."""

    def __init__(self, A, b, c, rule=0, threads=1, trace_cap=1 << 20, relative_infeasibility=False, drive_out=False):
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.b = np.ascontiguousarray(b, dtype=np.float64)
        self.c = np.ascontiguousarray(c, dtype=np.float64)
        self.n, self.m = self.A.shape
        self.l = lib()
        self.h = self.l.orc_create(self.n, self.m, self.A.ctypes.data_as(_dp), self.b.ctypes.data_as(_dp),
                                   self.c.ctypes.data_as(_dp), rule, threads)
        if relative_infeasibility:
            self.l.orc_set_relative_infeasibility(self.h, 1)
        if drive_out:
            self.l.orc_set_drive_out(self.h, 1)
        self._trace = np.zeros((trace_cap, 2), dtype=np.int32)
        self.l.orc_set_trace(self.h, self._trace.ctypes.data_as(_ip), trace_cap)

    def close(self):
        if self.h:
            self.l.orc_destroy(self.h)
            self.h = None

    __del__ = close

    def build_phase1(self):
        self.l.orc_build_phase1(self.h)

    def priceout(self):
        self.l.orc_priceout(self.h)

    def pivot(self):
        return self.l.orc_pivot(self.h)

    def iterate(self, budget=-1):
        return self.l.orc_iterate(self.h, budget)

    def phase1_verdict(self):
        return self.l.orc_phase1_verdict(self.h)

    def drive_out_artificials(self):
        return self.l.orc_drive_out_artificials(self.h)

    def switch_phase2(self):
        self.l.orc_switch_phase2(self.h)

    def extract(self):
        x = np.zeros(self.n); obj = C.c_double()
        self.l.orc_extract(self.h, x.ctypes.data_as(_dp), C.byref(obj))
        return x, obj.value

    def two_phase(self, max_pivots=-1):
        x = np.zeros(self.n); obj = C.c_double()
        st = self.l.orc_two_phase(self.h, max_pivots, x.ctypes.data_as(_dp), C.byref(obj))
        return {"status": st, "x": x, "objective": obj.value, "basis": self.basis(),
                "pivots": (self.l.orc_pivots(self.h, 1), self.l.orc_pivots(self.h, 2)),
                "hash": int(self.l.orc_hash(self.h)), "trace": self.trace()}

    def rows(self):
        return self.l.orc_rows(self.h)

    def tableau(self):
        R = self.rows()
        return np.ctypeslib.as_array(self.l.orc_tableau(self.h), shape=(R, self.m)).copy()

    def costs(self):
        return np.ctypeslib.as_array(self.l.orc_costs(self.h), shape=(self.rows(),)).copy()

    def basis(self):
        return np.ctypeslib.as_array(self.l.orc_basis(self.h), shape=(self.m,)).copy()

    def trace(self):
        k = min(self.l.orc_trace_len(self.h), self._trace.shape[0])
        return self._trace[:k].copy()

    def hash(self):
        return int(self.l.orc_hash(self.h))
