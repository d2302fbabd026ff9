"""Python handle over the C-ABI solver (include/b2s.h).  Thin: numpy in, numpy out."""
import ctypes as C

import numpy as np

from . import _lib as L


class B2SError(RuntimeError):
    pass


def seed_triplet(seed, flavour=L.RAND_GLIBC):
    """srand(seed); rand() x3 of the reference's generateRandomProblem (src/problem.cu:63-67)."""
    out = (C.c_uint * 3)()
    L.load().b2s_seed_triplet(C.c_uint(seed & 0xFFFFFFFF), int(flavour), out)
    return tuple(int(v) for v in out)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class Solver:
    """One device-resident two-phase simplex solver (opaque b2s_solver handle)."""

    def __init__(self, device=0, dtype=L.F64, pivot_rule=L.RULE_REFERENCE, fold_artificials=True,
                 skip_zero_rows=True, use_graph=True, batch=0, max_pivots=0, trace_capacity=0,
                 update_variant=8, persistent="auto", relative_infeasibility=False, lookahead="auto", fp64_polish=True, drive_out_artificials=False):
        self.lib = L.load()
        opt = L.Options()
        self.lib.b2s_default_options(C.byref(opt))
        opt.device = device
        opt.dtype = dtype
        opt.pivot_rule = pivot_rule
        opt.fold_artificials = int(bool(fold_artificials))
        opt.skip_zero_rows = int(bool(skip_zero_rows))
        opt.use_graph = int(bool(use_graph))
        opt.batch = batch
        opt.max_pivots = max_pivots
        opt.trace_capacity = trace_capacity
        opt.update_variant = update_variant
        opt.persistent = 2 if persistent in ("auto", None) else int(bool(persistent))
        opt.relative_infeasibility = int(bool(relative_infeasibility))
        opt.lookahead = 2 if lookahead in ("auto", None) else int(bool(lookahead))
        opt.fp64_polish = int(bool(fp64_polish))
        opt.drive_out_artificials = int(bool(drive_out_artificials))
        self.h = C.c_void_p()
        rc = self.lib.b2s_create(C.byref(opt), C.byref(self.h))
        if rc != L.OK:
            msg = self.lib.b2s_last_error(None)
            raise B2SError(f"b2s_create failed ({rc}): {msg.decode() if msg else ''}")
        self.n = self.m = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.b2s_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc != L.OK:
            msg = self.lib.b2s_last_error(self.h)
            raise B2SError(f"b2s error {rc}: {msg.decode() if msg else ''}")

    # ---- problem input ----------------------------------------------------------------------
    def load(self, A_varmajor, b, c):
        """A_varmajor[j, i] = coefficient of variable j in constraint i (problem_t layout)."""
        A = np.ascontiguousarray(A_varmajor, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        n, m = A.shape
        assert b.shape == (m,) and c.shape == (n,)
        self._ck(self.lib.b2s_load_problem_host(self.h, n, m, _dp(A), _dp(b), _dp(c)))
        self.n, self.m = n, m

    def generate(self, n, m, seeds, lo, hi):
        arr = (C.c_uint * 3)(*[int(s) & 0xFFFFFFFF for s in seeds])
        self._ck(self.lib.b2s_generate_problem_device(self.h, n, m, arr, float(lo), float(hi)))
        self.n, self.m = n, m

    def copy_problem(self):
        A = np.empty((self.n, self.m)); b = np.empty(self.m); c = np.empty(self.n)
        self._ck(self.lib.b2s_copy_problem(self.h, _dp(A), _dp(b), _dp(c)))
        return A, b, c

    # ---- whole solve ---------------------------------------------------------------------------
    def solve(self):
        """twoPhaseMethod: returns dict(status, x, objective, basis, stats)."""
        x = np.zeros(self.n); obj = C.c_double(0.0); basis = np.zeros(self.m, dtype=np.int32)
        status = C.c_int(0); stats = L.Stats()
        self._ck(self.lib.b2s_solve_two_phase(self.h, C.byref(status), _dp(x), C.byref(obj), _ip(basis), C.byref(stats)))
        return {"status": status.value, "x": x, "objective": obj.value, "basis": basis, "stats": stats}

    # ---- stepping ------------------------------------------------------------------------------
    def build_phase1(self):
        self._ck(self.lib.b2s_build_phase1(self.h))

    def price_out(self):
        self._ck(self.lib.b2s_price_out(self.h))

    def select_entering(self):
        self._ck(self.lib.b2s_select_entering(self.h))

    def iterate(self, max_pivots=-1):
        st = C.c_int(0); done = C.c_longlong(0)
        self._ck(self.lib.b2s_iterate(self.h, int(max_pivots), C.byref(st), C.byref(done)))
        return st.value, done.value

    def phase1_verdict(self):
        st = C.c_int(0)
        self._ck(self.lib.b2s_phase1_verdict(self.h, C.byref(st)))
        return st.value

    def switch_phase2(self):
        self._ck(self.lib.b2s_switch_phase2(self.h))

    def extract(self):
        x = np.zeros(self.n); obj = C.c_double(0.0)
        self._ck(self.lib.b2s_extract_solution(self.h, _dp(x), C.byref(obj)))
        return x, obj.value

    # ---- introspection -------------------------------------------------------------------------
    def dims(self):
        n = C.c_int(); m = C.c_int(); ra = C.c_longlong(); rs = C.c_longlong(); ld = C.c_longlong()
        self._ck(self.lib.b2s_get_dims(self.h, C.byref(n), C.byref(m), C.byref(ra), C.byref(rs), C.byref(ld)))
        return {"n": n.value, "m": m.value, "rows_active": ra.value, "rows_stored": rs.value, "ld": ld.value}

    def tableau(self):
        d = self.dims()
        out = np.empty((d["rows_active"], d["m"]))
        self._ck(self.lib.b2s_copy_tableau(self.h, _dp(out)))
        return out

    def costs(self):
        out = np.empty(self.dims()["rows_active"])
        self._ck(self.lib.b2s_copy_costs(self.h, _dp(out)))
        return out

    def basis(self):
        out = np.empty(self.m, dtype=np.int32)
        self._ck(self.lib.b2s_copy_basis(self.h, _ip(out)))
        return out

    def trace(self, capacity=1 << 20):
        n = C.c_longlong(); h = C.c_ulonglong()
        self._ck(self.lib.b2s_copy_trace(self.h, None, 0, C.byref(n), C.byref(h)))
        cnt = min(n.value, capacity)
        qp = np.zeros((max(cnt, 1), 2), dtype=np.int32)
        if cnt:
            self._ck(self.lib.b2s_copy_trace(self.h, _ip(qp), cnt, C.byref(n), C.byref(h)))
        return qp[:cnt], n.value, h.value

    def stats(self):
        st = L.Stats()
        self._ck(self.lib.b2s_get_stats(self.h, C.byref(st)))
        return st

    # ---- kernel-level hooks ----------------------------------------------------------------------
    def tournament(self, vec):
        v = np.ascontiguousarray(vec, dtype=np.float64)
        val = C.c_double(); idx = C.c_int()
        self._ck(self.lib.b2s_tournament(self.h, _dp(v), v.size, C.byref(val), C.byref(idx)))
        return val.value, idx.value

    def bench_update(self, launches=20, flush_l2=False):
        ms = (C.c_float * launches)(); nbytes = C.c_double()
        self._ck(self.lib.b2s_bench_update(self.h, launches, int(flush_l2), ms, C.byref(nbytes)))
        return np.array(list(ms), dtype=np.float64), nbytes.value

    def profile_pivots(self, count):
        """`count` real pivots with per-kernel CUDA-event times: dict of ms arrays + pivots made."""
        a = (C.c_float * count)(); b = (C.c_float * count)(); u = (C.c_float * count)(); done = C.c_longlong()
        self._ck(self.lib.b2s_profile_pivots(self.h, count, a, b, u, C.byref(done)))
        k = done.value
        return {"ratio_ms": np.array(a[:k]), "gather_ms": np.array(b[:k]), "update_ms": np.array(u[:k]), "pivots": k}

    def profile_lookahead(self, count, stages=True):
        """`count` pivots of the look-ahead kernel, one launch each: kernel ms + the chain's stage times (us from kernel start)."""
        ms = (C.c_float * count)(); us = (C.c_double * (6 * count))(); done = C.c_longlong()
        self._ck(self.lib.b2s_profile_lookahead(self.h, count, ms, us if stages else None, C.byref(done)))
        k = done.value
        out = {"kernel_ms": np.array(ms[:k]), "pivots": k}
        if stages:
            a = np.array(us[:6 * k]).reshape(k, 6)
            for j, name in enumerate(("rhs_row_us", "entering_known_us", "leaving_known_us", "pivot_row_complete_us",
                                      "proposal_ready_us", "committed_us")):
                out[name] = a[:, j]
        return out

    def loop_info(self):
        a = C.c_int(); b = C.c_int(); c = C.c_int()
        self._ck(self.lib.b2s_get_loop_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"launches_per_pivot": a.value, "lookahead": bool(b.value), "persistent": bool(c.value)}

    def loop_mode(self):
        i = self.loop_info()
        if i["persistent"]:
            return "persistent cooperative loop kernel (1 launch per batch of pivots)"
        if i["lookahead"]:
            return "look-ahead kernel: 1 launch per pivot (streaming update + next pivot's selection under it), CUDA graph"
        return f"{i['launches_per_pivot']} launches per pivot replayed as a CUDA graph"

    def launches_per_pivot(self):
        return self.loop_info()["launches_per_pivot"]

    def update_kernel_name(self):
        if self.loop_info()["lookahead"]:
            return "update_la_kernel (rank-1 update streamed over the row list + cost update + complete selection of the next pivot)"
        return "update_kernel (fused rank-1 update + cost update + entering tournament)"

    # ---- sharding --------------------------------------------------------------------------------
    def dist_init_host(self, rank, world, allgather):
        """Bootstrap the sharded solver through a host all-gather instead of NCCL.  `allgather(send_ptr, recv_ptr, nbytes)`
        works on raw host addresses and returns 0; it is called collectively by the library (b2s_allgather_fn)."""
        def _cb(user, send, recv, nbytes):
            try:
                return int(allgather(send, recv, nbytes))
            except Exception:   # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        self._host_ag = L.ALLGATHER_FN(_cb)   # keep the trampoline alive as long as the handle
        self._ck(self.lib.b2s_dist_init_host(self.h, rank, world, C.cast(self._host_ag, C.c_void_p), None))

    def dist_init(self, rank, world, unique_id):
        assert len(unique_id) == L.NCCL_ID_BYTES
        self._ck(self.lib.b2s_dist_init(self.h, rank, world, unique_id))


def dist_unique_id():
    buf = C.create_string_buffer(L.NCCL_ID_BYTES)
    rc = L.load().b2s_dist_unique_id(buf)
    if rc != L.OK:
        raise B2SError(f"b2s_dist_unique_id failed ({rc})")
    return buf.raw
