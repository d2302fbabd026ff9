"""ctypes binding of the C ABI declared in include/b2s.h (simplexoncuda_b200/lib/libb2s.so).

There is no CPU fallback: if the shared library is missing or no CUDA device is visible the
calls fail loudly.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``bash simplexoncuda_b200/csrc/build.sh``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb2s.so")

# error codes / statuses / options (mirror include/b2s.h)
OK, ERR_ARG, ERR_STATE, ERR_CUDA, ERR_NOMEM, ERR_NCCL, ERR_NOGPU, ERR_PEER = range(8)
FEASIBLE, INFEASIBLE, UNBOUNDED, DEGENERATE, ITER_LIMIT, RUNNING = 0, -1, -2, -3, -4, -10
F64, F32 = 0, 1
RULE_REFERENCE, RULE_LOWEST, RULE_BLAND = 0, 1, 2
RAND_GLIBC, RAND_MSVC = 0, 1
NCCL_ID_BYTES = 128

STATUS_NAMES = {FEASIBLE: "FEASIBLE", INFEASIBLE: "INFEASIBLE", UNBOUNDED: "UNBOUNDED",
                DEGENERATE: "DEGENERATE", ITER_LIMIT: "ITER_LIMIT", RUNNING: "RUNNING"}


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("dtype", C.c_int), ("pivot_rule", C.c_int),
                ("fold_artificials", C.c_int), ("skip_zero_rows", C.c_int), ("use_graph", C.c_int),
                ("batch", C.c_int), ("max_pivots", C.c_longlong), ("trace_capacity", C.c_longlong),
                ("update_variant", C.c_int), ("persistent", C.c_int), ("relative_infeasibility", C.c_int), ("lookahead", C.c_int),
                ("fp64_polish", C.c_int), ("drive_out_artificials", C.c_int), ("reserved", C.c_int * 2)]


class Stats(C.Structure):
    _fields_ = [("pivots_phase1", C.c_longlong), ("pivots_phase2", C.c_longlong),
                ("trace_hash", C.c_ulonglong), ("seconds_total", C.c_double),
                ("seconds_load", C.c_double), ("seconds_phase1", C.c_double),
                ("seconds_phase2", C.c_double), ("rows_streamed", C.c_longlong),
                ("rows_total", C.c_longlong), ("reserved", C.c_longlong * 4)]


# every symbol include/b2s.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
_LL = C.POINTER(C.c_longlong)
SIGNATURES = {
    "b2s_default_options": (None, [C.POINTER(Options)]),
    "b2s_create": (C.c_int, [C.POINTER(Options), C.POINTER(_P)]),
    "b2s_destroy": (None, [_P]),
    "b2s_last_error": (C.c_char_p, [_P]),
    "b2s_device_count": (C.c_int, []),
    "b2s_load_problem_host": (C.c_int, [_P, C.c_int, C.c_int, _D, _D, _D]),
    "b2s_generate_problem_device": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_uint), C.c_double, C.c_double]),
    "b2s_seed_triplet": (None, [C.c_uint, C.c_int, C.POINTER(C.c_uint)]),
    "b2s_copy_problem": (C.c_int, [_P, _D, _D, _D]),
    "b2s_solve_two_phase": (C.c_int, [_P, _I, _D, _D, _I, C.POINTER(Stats)]),
    "b2s_build_phase1": (C.c_int, [_P]),
    "b2s_price_out": (C.c_int, [_P]),
    "b2s_select_entering": (C.c_int, [_P]),
    "b2s_iterate": (C.c_int, [_P, C.c_longlong, _I, _LL]),
    "b2s_phase1_verdict": (C.c_int, [_P, _I]),
    "b2s_switch_phase2": (C.c_int, [_P]),
    "b2s_extract_solution": (C.c_int, [_P, _D, _D]),
    "b2s_get_dims": (C.c_int, [_P, _I, _I, _LL, _LL, _LL]),
    "b2s_copy_tableau": (C.c_int, [_P, _D]),
    "b2s_copy_costs": (C.c_int, [_P, _D]),
    "b2s_copy_basis": (C.c_int, [_P, _I]),
    "b2s_copy_trace": (C.c_int, [_P, _I, C.c_longlong, _LL, C.POINTER(C.c_ulonglong)]),
    "b2s_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "b2s_tournament": (C.c_int, [_P, _D, C.c_longlong, _D, _I]),
    "b2s_bench_update": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_float), _D]),
    "b2s_attach_tableau_device": (C.c_int, [_P, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2s_set_basis": (C.c_int, [_P, _I]),
    "b2s_min_element_device": (C.c_int, [_P, C.c_void_p, C.c_longlong, _D, C.POINTER(C.c_uint)]),
    "b2s_ratio_min_device": (C.c_int, [_P, C.c_void_p, C.c_void_p, C.c_longlong, _D, C.POINTER(C.c_uint)]),
    "b2s_max_le_zero_device": (C.c_int, [_P, C.c_void_p, C.c_longlong, _I]),
    "b2s_profile_pivots": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), _LL]),
    "b2s_get_loop_info": (C.c_int, [_P, _I, _I, _I]),
    "b2s_profile_lookahead": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _D, _LL]),
    "b2s_dist_init_host": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "b2s_dist_unique_id": (C.c_int, [C.c_char_p]),
    "b2s_dist_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p]),
}

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)  # b2s_allgather_fn

_lib = None


def load():
    """Load libb2s.so (once) and attach the prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension is not built (run __graft_entry__.build()); "
            "simplexoncuda_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
