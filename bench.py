#!/usr/bin/env python
"""bench.py -- pivots/s and rank-1-update HBM GB/s of the two-phase dense-tableau simplex on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (C ABI, libb2s.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference arm (see below)

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): the reference's
published benchmark instance random_8192_8192 -- generateRandomProblem(8192, 8192, seed 827392,
1, 100) with the MSVC seed derivation of the published runs -- whose tableau is the "8192x16384
fp64" shape (16385 stored rows x 8192 constraints, 1.07 GB; with the artificial rows folded onto
the slack rows that shape holds in both phases).  A "step" is --pivots-per-step consecutive simplex
pivots of the real solve (ratio test, gather/normalise, fused rank-1 update + next entering
column), tableau resident in HBM.  `value` = pivots/s over the K timed steps (device time from
CUDA events on the solver's stream, max over ranks).  `e2e` = the same metric for one complete
two-phase solve through the host-buffer C-ABI call path (b2s_load_problem_host from pinned host
memory + b2s_solve_two_phase + results back on the host), copies inside the timed region.

--gpus N > 1 (torchrun, one rank per GPU): the same LP, constraint-sharded over the ranks (strong
scaling) with one all-gather + one all-reduce over NCCL per pivot.

Reference arm: `--impl reference` runs the UNMODIFIED reference (oracle/_ref/libsimplex_ref.so,
built from /root/reference by oracle/build_ref.sh; its own CUDA kernels, its own host loop) on the
same LP in a fresh subprocess on rank 0's GPU; if that build is absent it falls back to the serial
oracle port on the host cores.  The reference has no CPU implementation of this path; its
"cpu_baseline" entry says which of the two ran.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pivots/s (two-phase dense-tableau simplex, 8192x16384 fp64 tableau) and rank-1 update HBM GB/s"
UNIT = "pivots/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2s", choices=["b2s", "reference"])
    ap.add_argument("--vars", type=int, default=8192)
    ap.add_argument("--constraints", type=int, default=8192)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--pivots-per-step", type=int, default=100)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-pivots", type=int, default=0, help="pivot budget of the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-skip-zero-rows", dest="skip_zero_rows", action="store_false",
                    help="stream every stored row (the nominal 2*R*m*8 bytes per pivot); default: rows whose pivot-row "
                         "entry is exactly 0 are not streamed (value-exact, the library default)")
    ap.add_argument("--skip-zero-rows", dest="skip_zero_rows", action="store_true")
    ap.set_defaults(skip_zero_rows=True)
    ap.add_argument("--no-latency-config", dest="latency_config", action="store_false",
                    help="skip the supplementary 1024x1024 complete solve (BASELINE.json configs[2])")
    ap.add_argument("--no-compat-e2e", dest="compat_e2e", action="store_false",
                    help="skip timing the drop-in twoPhaseMethod() entry of libb2s_compat.so")
    ap.add_argument("--no-large-config", dest="large_config", action="store_false",
                    help="skip the supplementary 65536x65536 measurement")
    ap.add_argument("--update-variant", type=int, default=8)
    ap.add_argument("--loop", default="auto", choices=["auto", "persistent", "launches"],
                    help="persistent cooperative loop kernel, three launches per pivot (CUDA graph), or the library's choice")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


def default_seed(a):
    return a.seed if a.seed is not None else a.vars * 100 + a.constraints  # main.cu:63


def workload_name(a):
    return (f"random_{a.vars}_{a.constraints} (generateRandomProblem seed {default_seed(a)}, range [1,100], MSVC seed "
            f"derivation = the reference's published instance); tableau {a.constraints}x{a.vars + a.constraints} fp64")


# ------------------------------------------------------------------------------------------------
# cpu_baseline: the serial oracle (port) on the box's host cores, bounded pivot budget
# ------------------------------------------------------------------------------------------------
def cpu_baseline(a, threads=None, budget=None):
    exe = os.path.join(ROOT, "oracle", "serial_tableau")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    threads = threads or (os.cpu_count() or 1)
    if budget is None:
        # ~2*R*m*8 bytes per pivot at O(10 GB/s/..): aim for roughly 10-20 s
        per_pivot = 2.0 * (1 + a.vars + 2 * a.constraints) * a.constraints * 8 / 25e9
        budget = a.cpu_pivots or int(max(20, min(2000, 15.0 / max(per_pivot, 1e-6))))
    out = subprocess.run([exe, str(a.vars), str(a.constraints), str(default_seed(a)), "1", "100", "1", "0",
                          str(threads), str(budget)], capture_output=True, text=True, check=True).stdout.split()
    pivots = int(out[1]) + int(out[2])
    secs = float(out[6]) if len(out) > 6 else float(out[5])
    return {"value": pivots / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {pivots} pivots of the same LP (serial oracle restatement, OpenMP over tableau rows, "
                      f"{secs:.2f} s in the pivot loop)"}


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libsimplex_ref.so")
    have_gpu_ref = os.path.exists(ref_lib) and not os.environ.get("B2S_BENCH_FORCE_PORT")
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "pivots_per_step": a.pivots_per_step}}
    if have_gpu_ref:
        import numpy as np
        import oracle_py as O
        A, b, c = O.generate(a.vars, a.constraints, O.seed_triplet(default_seed(a), 1), 1, 100)
        tmp = tempfile.mkdtemp(prefix="b2s_ref_")
        prob = os.path.join(tmp, "prob.npz")
        np.savez(prob, A=A, b=b, c=c)
        del A, b, c
        sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
        sampler.start()
        t0 = time.time()
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), prob, os.path.join(tmp, "ref")],
                       check=True, capture_output=True, timeout=3000)
        wall = time.time() - t0
        clocks = sampler.stop()
        res = json.load(open(os.path.join(tmp, "ref.json")))
        pivots = res["pivots_phase1"] + res["pivots_phase2"]
        loop = res["seconds_loop_phase1"] + res["seconds_loop_phase2"]
        value = pivots / loop
        e2e = pivots / res["seconds_total"]
        sample = (f"one complete two-phase solve by the unmodified reference CUDA build (oracle/_ref, sm_100) on this box's "
                  f"GPU 0: {res['pivots_phase1']}+{res['pivots_phase2']} pivots, status {res['status']}, {loop:.2f} s in its "
                  f"pivot loops, {res['seconds_total']:.2f} s in twoPhaseMethod, {wall:.1f} s process wall")
        line.update(value=value, ms_per_step=1e3 * a.pivots_per_step / value, clocks=clocks, gpu_launches=0,
                    cpu_baseline={"value": value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample,
                                  "device": "NVIDIA B200 (the reference has no CPU path; its host loop is one thread)"},
                    e2e={"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    per_phase={"phase1_pivots_per_s": res["pivots_phase1"] / max(res["seconds_loop_phase1"], 1e-9),
                               "phase2_pivots_per_s": res["pivots_phase2"] / max(res["seconds_loop_phase2"], 1e-9),
                               "phase1_tableau_rows": 1 + a.vars + 2 * a.constraints,
                               "phase2_tableau_rows": 1 + a.vars + a.constraints})
        # supplementary, same box, same process model: BASELINE.json configs[2] (1024 x 2048 tableau, latency-bound regime)
        if a.latency_config and (a.vars, a.constraints) == (8192, 8192):
            try:
                ln = lm = 1024
                A, b, c = O.generate(ln, lm, O.seed_triplet(ln * 100 + lm, 1), 1, 100)
                prob2 = os.path.join(tmp, "prob_lat.npz")
                np.savez(prob2, A=A, b=b, c=c)
                subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), prob2, os.path.join(tmp, "lat")],
                               check=True, capture_output=True, timeout=600)
                rl = json.load(open(os.path.join(tmp, "lat.json")))
                pl = rl["pivots_phase1"] + rl["pivots_phase2"]
                ls = rl["seconds_loop_phase1"] + rl["seconds_loop_phase2"]
                line["latency_config"] = {"workload": f"random_{ln}_{lm} (seed {ln * 100 + lm}, [1,100], MSVC seeds); tableau "
                                                      f"{lm}x{ln + lm} fp64", "pivots_per_s": pl / ls, "us_per_pivot": 1e6 * ls / pl,
                                          "pivots": pl, "pivots_per_s_wall": pl / rl["seconds_total"], "status": rl["status"],
                                          "timed": "the unmodified reference's two pivot loops (host-driven, 6 device syncs per pivot)"}
            except Exception as exc:
                line["latency_config"] = {"skipped": str(exc)[:200]}
    else:
        cb = cpu_baseline(a)
        cb["sample"] += "; oracle/_ref absent, so the oracle port stands in for the reference"
        line.update(value=cb["value"], ms_per_step=1e3 * a.pivots_per_step / cb["value"], cpu_baseline=cb, gpu_launches=0,
                    clocks={"sm_mhz": None, "sm_max_mhz": None, "reasons": ["cpu run"]},
                    e2e={"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def fixture(n, m, seed):
    path = os.path.join(ROOT, "tests", "golden", "oracle_results.json")
    try:
        return json.load(open(path)).get(f"{n}_{m}_{seed}")
    except Exception:
        return None


def parity_of_solve(r, fx):
    """Compare one complete solve with the committed oracle fixture (pivot counts, (q,p) hash, objective)."""
    if fx is None:
        return {"status": "unpinned", "why": "no committed oracle fixture for this instance"}
    got = {"pivots": [int(r["stats"].pivots_phase1), int(r["stats"].pivots_phase2)],
           "trace_hash": str(int(r["stats"].trace_hash)), "objective": r["objective"], "solver_status": int(r["status"])}
    want = {"pivots": [fx["pivots_phase1"], fx["pivots_phase2"]], "trace_hash": fx["trace_hash"],
            "objective": fx["objective"], "solver_status": fx["status"]}
    ok = got == want
    out = {"status": "ok" if ok else "MISMATCH", "trace_hash": got["trace_hash"], "pivots": got["pivots"],
           "objective_bit_exact": got["objective"] == want["objective"], "against": "tests/golden/oracle_results.json"}
    if not ok:
        out["expected"] = want
        out["got"] = got
    return out


def compat_two_phase(A, b, c):
    """The reference's own entry point, `int twoPhaseMethod(problem_t*, TYPE*, TYPE*)` (include/twoPhaseMethod.h:10-19),
    as exported by libb2s_compat.so: pageable problem_t in, solution/optimum out.  Its progress lines go to /dev/null."""
    import numpy as np
    lib = ctypes.CDLL(os.path.join(ROOT, "simplexoncuda_b200", "lib", "libb2s_compat.so"), mode=ctypes.RTLD_GLOBAL)
    fn = getattr(lib, "_Z14twoPhaseMethodP9problem_tPdS1_")

    class ProblemT(ctypes.Structure):   # include/problem.h:10-26
        _fields_ = [("constraintsMatrix", ctypes.c_void_p), ("knownTermsVector", ctypes.c_void_p),
                    ("objectiveFunction", ctypes.c_void_p), ("vars", ctypes.c_int), ("constraints", ctypes.c_int)]
    n, m = A.shape
    prob = ProblemT(A.ctypes.data, b.ctypes.data, c.ctypes.data, n, m)
    x = np.zeros(n)
    opt = ctypes.c_double(0.0)
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.POINTER(ProblemT), ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        t0 = time.time()
        status = fn(ctypes.byref(prob), x.ctypes.data, ctypes.byref(opt))
        secs = time.time() - t0
    finally:
        ctypes.CDLL(None).fflush(None)
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    return status, x, opt.value, secs


def run_b2s(a):
    import numpy as np
    import torch
    import simplexoncuda_b200 as S
    from simplexoncuda_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, m, P = a.vars, a.constraints, a.pivots_per_step
    seeds = S.seed_triplet(default_seed(a), S.RAND_MSVC)
    loop_opt = {"auto": "auto", "persistent": True, "launches": False}[a.loop]
    mismatches = []

    def new_solver(skip):
        h = S.Solver(device=local, skip_zero_rows=skip, update_variant=a.update_variant, persistent=loop_opt)
        if world > 1:
            sharding.init_sharded_solver(h, dist)
        return h

    s = new_solver(a.skip_zero_rows)
    s.generate(n, m, seeds, 1, 100)
    s.build_phase1(); s.price_out(); s.select_entering()
    dims = s.dims()
    elem = 8
    slab_bytes = dims["rows_stored"] * (m // world) * elem
    loop_mode = s.loop_mode() if hasattr(s, "loop_mode") else "launches"
    bytes_per_pivot = 2.0 * dims["rows_stored"] * (m // world) * elem  # per rank: read + write of the stored slab

    for _ in range(a.warmup):
        st, done = s.iterate(P)
        assert done == P, f"phase ended during warm-up ({st}, {done})"
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.time()
    st0 = s.stats()
    for _ in range(a.steps):
        st, done = s.iterate(P)
        assert done == P, f"phase ended inside the timed region ({st}, {done})"
    barrier()
    wall = time.time() - t0
    st1 = s.stats()
    clocks = sampler.stop()
    dev_ms = max_over_ranks((st1.seconds_phase1 - st0.seconds_phase1) * 1e3)
    pivots = a.steps * P
    value = pivots / (dev_ms * 1e-3)
    rows_frac_timed = None
    if a.skip_zero_rows and st1.rows_total > st0.rows_total:
        rows_frac_timed = (st1.rows_streamed - st0.rows_streamed) / (st1.rows_total - st0.rows_total)

    # ---- roofline of the dominant kernel, live: up to 100 more real pivots with per-kernel CUDA events -------------
    def live_roofline(h, label):
        b0 = h.stats()
        prof = h.profile_pivots(min(P, 100))
        b1 = h.stats()
        upd = float(np.mean(prof["update_ms"]))
        tot = float(np.mean(prof["update_ms"] + prof["ratio_ms"] + prof["gather_ms"]))
        frac_rows = 1.0
        if b1.rows_streamed > b0.rows_streamed and b1.rows_total > b0.rows_total:
            frac_rows = (b1.rows_streamed - b0.rows_streamed) / (b1.rows_total - b0.rows_total)
        peak, peak_src = peaks()
        moved = bytes_per_pivot * frac_rows
        achieved = moved / (upd * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": label, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "frac_of_nominal_8TBps": achieved / 8000.0,
                "bytes_moved_per_launch": moved, "rows_streamed_fraction": frac_rows,
                "algorithmic_bytes_per_launch": bytes_per_pivot,
                "effective_GBps_on_algorithmic_bytes": bytes_per_pivot / (upd * 1e-3) / 1e9,
                "launch_ms": upd, "kernel_share_of_pivot": upd / tot,
                "other_kernels_ms": {"ratio": float(np.mean(prof["ratio_ms"])), "gather": float(np.mean(prof["gather_ms"]))}}

    roofline = None
    if world == 1:
        roofline = live_roofline(s, s.update_kernel_name() if hasattr(s, "update_kernel_name") else
                                 "update_kernel (fused rank-1 update + cost update + entering tournament)")
        roofline["whole_pivot_GBps_on_algorithmic_bytes"] = bytes_per_pivot * value / 1e9
        roofline["traffic"] = traffic_from_profile(n, m, a.skip_zero_rows)
    launches_per_pivot = s.launches_per_pivot() if hasattr(s, "launches_per_pivot") else (3 if world == 1 else 4)
    if "persistent" in loop_mode:
        batch = int(min(256.0, max(4.0, 3e-3 / max(12e-6, 2.0 * slab_bytes / 5.0e12))))
        launches = a.steps * ((P + batch - 1) // batch + 1)
    else:
        launches = launches_per_pivot * pivots
        if launches_per_pivot == 1:   # look-ahead kernel: + prologue (2) and column write-back (1) per iterate() call
            launches += 3 * a.steps
    s.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "pivots_per_step": P, "phase": 1,
                       "tableau_rows_stored": dims["rows_stored"], "tableau_bytes": dims["rows_stored"] * m * elem,
                       "l2": (f"inputs larger than L2 (tableau {dims['rows_stored'] * m * elem / 1e9:.2f} GB vs 126 MB), no flush"
                              if dims["rows_stored"] * m * elem > 4 * 126e6 else
                              f"tableau {dims['rows_stored'] * m * elem / 1e6:.1f} MB is L2-resident: latency-bound regime, no flush"),
                       "parallelism": f"constraint slabs x{world}" if world > 1 else "single GPU",
                       "skip_zero_rows": bool(a.skip_zero_rows), "update_variant": a.update_variant,
                       "loop": loop_mode},
            "clocks": clocks, "gpu_launches": launches, "wall_s_timed_region": wall}
    if rows_frac_timed is not None:
        line["config"]["rows_streamed_fraction"] = rows_frac_timed
    if roofline:
        line["roofline"] = roofline
    parity = {}

    # ---- nominal-bytes line: the same timed protocol with every stored row streamed (skip_zero_rows off) ----------
    if world == 1 and a.skip_zero_rows:
        z = new_solver(False)
        z.generate(n, m, seeds, 1, 100)
        z.build_phase1(); z.price_out(); z.select_entering()
        for _ in range(a.warmup):
            z.iterate(P)
        torch.cuda.synchronize()
        z0 = z.stats().seconds_phase1
        nsteps = max(3, min(a.steps, 10))
        for _ in range(nsteps):
            z.iterate(P)
        zms = (z.stats().seconds_phase1 - z0) * 1e3
        zr = live_roofline(z, "the same kernel with skip_zero_rows = 0 (every stored row read and written)")
        zv = nsteps * P / (zms * 1e-3)
        zr["whole_pivot_GBps"] = bytes_per_pivot * zv / 1e9
        zr["whole_pivot_frac_of_nominal_8TBps"] = bytes_per_pivot * zv / 8e12
        line["no_skip_zero_rows"] = {"value": zv, "unit": UNIT, "steps": nsteps, "roofline": zr}
        z.close()

    # ---- e2e: one complete solve through the host-buffer API ----------------------------------------
    Ap = bp = cp = None
    if not a.no_e2e and n * m * 8 <= 8e9:
        # the host-buffer path a caller of twoPhaseMethod() takes: arrays in (pinned) host memory -> tableau
        # (every rank copies its own constraint slab) -> complete two-phase solve -> x, objective, basis on the host
        with S.Solver(device=local) as g:
            g.generate(n, m, seeds, 1, 100)
            A, b, c = g.copy_problem()
        Ap = torch.from_numpy(A).pin_memory(); bp = torch.from_numpy(b).pin_memory(); cp = torch.from_numpy(c).pin_memory()
        e = new_solver(a.skip_zero_rows)
        # warm the handle like a caller that solves more than one LP: buffers, (sharded) peer-memory arenas and
        # their IPC mappings are created by the first load and reused; the timed region starts from host arrays
        e.load(Ap.numpy(), bp.numpy(), cp.numpy())
        barrier()
        t0 = time.time()
        e.load(Ap.numpy(), bp.numpy(), cp.numpy())
        r = e.solve()
        t1 = time.time()
        e.close()
        secs = max_over_ranks(t1 - t0)
        piv = r["stats"].pivots_phase1 + r["stats"].pivots_phase2
        line["e2e"] = {"value": piv / secs, "unit": UNIT, "h2d_bytes_per_step": int((n * m + n + m) * 8),
                       "d2h_bytes_per_step": int(n * 8 + 8 + m * 4) * world,
                       "step": f"one complete two-phase solve from pinned host arrays: status {r['status']}, "
                               f"{r['stats'].pivots_phase1}+{r['stats'].pivots_phase2} pivots in {secs:.3f} s "
                               f"(max over ranks; load {r['stats'].seconds_load:.3f} s)", "objective": r["objective"],
                       "rows_streamed_fraction": r["stats"].rows_streamed / max(1, r["stats"].rows_total) if a.skip_zero_rows else 1.0,
                       "per_phase": {"phase1_pivots_per_s": r["stats"].pivots_phase1 / max(r["stats"].seconds_phase1, 1e-9),
                                     "phase2_pivots_per_s": r["stats"].pivots_phase2 / max(r["stats"].seconds_phase2, 1e-9),
                                     "tableau_rows_stored_both_phases": dims["rows_stored"]}}
        parity["e2e"] = parity_of_solve(r, fixture(n, m, default_seed(a)))
        # every rank must hold the same replicated result
        if dist is not None:
            box = [None] * world
            dist.all_gather_object(box, (int(r["status"]), str(int(r["stats"].trace_hash)), r["objective"]))
            parity["e2e"]["ranks_agree"] = all(x == box[0] for x in box)
            if not parity["e2e"]["ranks_agree"]:
                parity["e2e"]["status"] = "MISMATCH"
    elif not a.no_e2e:
        line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "step": "skipped: the constraint matrix of this configuration does not fit in host memory; "
                               "the instance only ever exists on the devices"}

    # ---- e2e through the reference's own entry point (drop-in path): pageable problem_t -> twoPhaseMethod() --------
    if world == 1 and a.compat_e2e and Ap is not None:
        try:
            A2 = np.array(Ap.numpy(), copy=True); b2 = np.array(bp.numpy(), copy=True); c2 = np.array(cp.numpy(), copy=True)
            st_c, x_c, opt_c, secs_c = compat_two_phase(A2, b2, c2)       # first call creates the process-wide handle
            st_c, x_c, opt_c, secs_c = compat_two_phase(A2, b2, c2)
            fx = fixture(n, m, default_seed(a))
            piv_c = (fx["pivots_phase1"] + fx["pivots_phase2"]) if fx else piv
            line["e2e_compat"] = {"value": piv_c / secs_c, "unit": UNIT, "seconds": secs_c, "status": int(st_c),
                                  "entry": "twoPhaseMethod(problem_t*, TYPE*, TYPE*) of libb2s_compat.so, pageable malloc'd "
                                           "arrays, H2D + solve + solution inside the timed region (include/twoPhaseMethod.h:10-19)",
                                  "objective": opt_c}
            if fx:
                ok = (st_c == fx["status"]) and (opt_c == fx["objective"])
                parity["e2e_compat"] = {"status": "ok" if ok else "MISMATCH", "objective_bit_exact": opt_c == fx["objective"]}
            del A2
        except Exception as exc:
            line["e2e_compat"] = {"skipped": str(exc)[:200]}
    if not a.no_cpu_baseline and rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline(a)

    # ---- supplementary: BASELINE.json configs[2], the launch/latency-bound regime (1024 x 2048 tableau) -----------
    if a.latency_config and world == 1 and (n, m) == (8192, 8192):
        ln = lm = 1024
        lseed = ln * 100 + lm
        with S.Solver(device=local) as g:
            g.generate(ln, lm, S.seed_triplet(lseed, S.RAND_MSVC), 1, 100)
            g.solve()                                    # warm: graphs / cooperative kernel, clocks
            g.generate(ln, lm, S.seed_triplet(lseed, S.RAND_MSVC), 1, 100)
            torch.cuda.synchronize()
            t0 = time.time()
            rl = g.solve()
            t1 = time.time()
            lmode = g.loop_mode() if hasattr(g, "loop_mode") else "auto"
        pl = rl["stats"].pivots_phase1 + rl["stats"].pivots_phase2
        dev_s = rl["stats"].seconds_phase1 + rl["stats"].seconds_phase2
        line["latency_config"] = {"workload": f"random_{ln}_{lm} (seed {lseed}, [1,100], MSVC seeds); tableau {lm}x{ln + lm} fp64, "
                                              "16.8 MB stored: L2-resident, launch/latency-bound",
                                  "pivots_per_s": pl / dev_s, "us_per_pivot": 1e6 * dev_s / pl, "pivots": pl,
                                  "pivots_per_s_wall": pl / (t1 - t0), "loop": lmode,
                                  "timed": "complete two-phase solve, device time of the two pivot loops incl. price-out"}
        parity["latency_config"] = parity_of_solve(rl, fixture(ln, lm, lseed))

    # ---- supplementary: BASELINE.json configs[4], the sharded shape (65536 x 131072 fp64, 68.7 GB tableau) ----
    # Same protocol as `value` on a short pivot budget; generated on the devices (its constraint matrix alone is
    # 34 GB).  Not part of `value`: it documents the north star's "near-linear scaling at 65536 x 131072".
    if a.large_config and (n, m) == (8192, 8192):
        try:
            ln = lm = 65536
            g = S.Solver(device=local, update_variant=a.update_variant)
            if world > 1:
                sharding.init_sharded_solver(g, dist)
            g.generate(ln, lm, S.seed_triplet(ln * 100 + lm, S.RAND_MSVC), 1, 100)
            g.build_phase1(); g.price_out(); g.select_entering()
            ldims = g.dims()
            g.iterate(5)
            barrier()
            b0 = g.stats()
            st_l, done_l = g.iterate(20)
            barrier()
            b1 = g.stats()
            ms_l = max_over_ranks((b1.seconds_phase1 - b0.seconds_phase1) * 1e3)
            qp, cnt, h_l = g.trace()
            g.close()
            pps = done_l / (ms_l * 1e-3)
            slab = 2.0 * ldims["rows_stored"] * (lm // world) * elem
            frac = (b1.rows_streamed - b0.rows_streamed) / max(1, b1.rows_total - b0.rows_total) if b1.rows_streamed else 1.0
            line["large_config"] = {"workload": "random_65536_65536 (seed 6619136, [1,100]); tableau 65536x131072 fp64, "
                                                f"{ldims['rows_stored'] * lm * elem / 1e9:.1f} GB, constraint slabs x{world}",
                                    "pivots_per_s": pps, "pivots_timed": done_l, "rows_streamed_fraction": frac,
                                    "per_gpu_stream_GBps_moved": slab * frac * pps / 1e9,
                                    "frac_of_measured_hbm_peak_moved": slab * frac * pps / 1e9 / peaks()[0],
                                    "effective_GBps_per_gpu_on_algorithmic_bytes": slab * pps / 1e9}
            lfx = None
            lpath = os.path.join(ROOT, "tests", "golden", "large_config_trace.json")
            if os.path.exists(lpath):
                lfx = json.load(open(lpath))
            got = {"pivots": int(cnt), "trace_hash": str(int(h_l))}
            if lfx is None:
                parity["large_config"] = dict(got, status="unpinned", why="tests/golden/large_config_trace.json absent")
            else:
                ok = got["pivots"] == lfx["pivots"] and got["trace_hash"] == lfx["trace_hash"]
                parity["large_config"] = dict(got, status="ok" if ok else "MISMATCH", against="tests/golden/large_config_trace.json "
                                              "(first 25 (q,p) pairs of the single-GPU solve)")
                if not ok:
                    parity["large_config"]["expected"] = lfx
        except Exception as exc:  # e.g. not enough free HBM on a shared box
            line["large_config"] = {"skipped": str(exc)[:200]}
    bad = [k for k, v in parity.items() if v.get("status") == "MISMATCH"]
    parity["overall"] = "MISMATCH" if bad else ("ok" if parity else "not run")
    line["parity"] = parity
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 1 if bad else 0


def traffic_from_profile(n, m, skip):
    """dram bytes per launch of the update kernel from the committed ncu capture -- only if that capture was taken from the
    kernel source now in the tree (its sha256 is stamped into the JSON by tools/stamp_traffic.py) and of this workload."""
    import hashlib
    tpath = os.path.join(ROOT, "profiles", "update_kernel_traffic.json")
    if not os.path.exists(tpath):
        return None
    t = json.load(open(tpath))
    if (n, m) != tuple(t.get("vars_constraints", (8192, 8192))) or bool(t.get("skip_zero_rows", False)) != bool(skip):
        return None
    h = hashlib.sha256()
    for f in t.get("kernel_sources", []):
        fp = os.path.join(ROOT, f)
        if not os.path.exists(fp):
            return None
        h.update(open(fp, "rb").read())
    if not t.get("kernel_sources") or h.hexdigest() != t.get("kernel_sources_sha256"):
        return None
    return t.get("dram_bytes_per_launch")


if __name__ == "__main__":
    args = parse()
    sys.exit(run_reference(args) if args.impl == "reference" else run_b2s(args))
