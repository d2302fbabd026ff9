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
    ap.add_argument("--skip-zero-rows", action="store_true")
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

    n, m, P = a.vars, a.constraints, a.pivots_per_step
    seeds = S.seed_triplet(default_seed(a), S.RAND_MSVC)
    s = S.Solver(device=local, skip_zero_rows=a.skip_zero_rows, update_variant=a.update_variant,
                 persistent={"auto": "auto", "persistent": True, "launches": False}[a.loop])
    if world > 1:
        sharding.init_sharded_solver(s, dist)
    s.generate(n, m, seeds, 1, 100)
    s.build_phase1(); s.price_out(); s.select_entering()
    dims = s.dims()
    elem = 8
    slab_bytes = dims["rows_stored"] * (m // world) * elem
    persistent_on = a.loop == "persistent" or (a.loop == "auto" and world == 1 and slab_bytes < 32e6)
    loop_mode = ("persistent cooperative loop kernel (1 launch per batch of pivots)" if persistent_on
                 else "3 launches per pivot replayed as a CUDA graph")
    bytes_per_pivot = 2.0 * dims["rows_stored"] * (m // world) * elem  # per rank: read + write of the stored slab

    for _ in range(a.warmup):
        st, done = s.iterate(P)
        assert done == P, f"phase ended during warm-up ({st}, {done})"
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    dev_ms = 0.0
    t0 = time.time()
    before = s.stats().seconds_phase1
    for _ in range(a.steps):
        st, done = s.iterate(P)
        assert done == P, f"phase ended inside the timed region ({st}, {done})"
    barrier()
    wall = time.time() - t0
    dev_ms = (s.stats().seconds_phase1 - before) * 1e3
    clocks = sampler.stop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    pivots = a.steps * P
    value = pivots / (dev_ms * 1e-3)

    # ---- roofline of the dominant kernel, live: P more pivots with per-kernel events -------------------
    roofline = None
    if world == 1:
        prof = s.profile_pivots(min(P, 100))
        upd = float(np.mean(prof["update_ms"]))
        tot = float(np.mean(prof["update_ms"] + prof["ratio_ms"] + prof["gather_ms"]))
        peak, peak_src = peaks()
        achieved = bytes_per_pivot / (upd * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "update_kernel_traffic.json")
        if os.path.exists(tpath) and (n, m) == (8192, 8192):   # the ncu capture under profiles/ is of this workload
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        roofline = {"bound": "hbm", "kernel": "update_kernel (fused rank-1 update + cost update + entering tournament)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "frac_of_nominal_8TBps": achieved / 8000.0, "traffic": traffic,
                    "algorithmic_bytes_per_launch": bytes_per_pivot, "launch_ms": upd,
                    "kernel_share_of_pivot": upd / tot,
                    "other_kernels_ms": {"ratio": float(np.mean(prof["ratio_ms"])), "gather": float(np.mean(prof["gather_ms"]))},
                    "whole_pivot_GBps": bytes_per_pivot * value / 1e9}
    # kernels launched inside the timed region: per-pivot kernels, or one cooperative loop kernel per batch
    # (the library's batch: ~3 ms of pivots, 4..256 -- b2s_solver.cu pick_batch)
    if "persistent" in loop_mode:
        batch = int(min(256.0, max(4.0, 3e-3 / max(12e-6, 2.0 * slab_bytes / 5.0e12))))
        launches = a.steps * ((P + batch - 1) // batch + 1)
    else:
        launches = (3 if world == 1 else 4) * pivots
    stats = s.stats()
    rows_note = None
    if a.skip_zero_rows and stats.rows_total:
        rows_note = stats.rows_streamed / stats.rows_total
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
    if rows_note is not None:
        line["config"]["rows_streamed_fraction"] = rows_note
    if roofline:
        line["roofline"] = roofline

    # ---- e2e: one complete solve through the host-buffer API ----------------------------------------
    if not a.no_e2e and n * m * 8 <= 8e9:
        # the host-buffer path a caller of twoPhaseMethod() takes: arrays in (pinned) host memory -> tableau
        # (every rank copies its own constraint slab) -> complete two-phase solve -> x, objective, basis on the host
        with S.Solver(device=local) as g:
            g.generate(n, m, seeds, 1, 100)
            A, b, c = g.copy_problem()
        Ap = torch.from_numpy(A).pin_memory(); bp = torch.from_numpy(b).pin_memory(); cp = torch.from_numpy(c).pin_memory()
        del A
        e = S.Solver(device=local, skip_zero_rows=a.skip_zero_rows, update_variant=a.update_variant,
                     persistent={"auto": "auto", "persistent": True, "launches": False}[a.loop])
        if world > 1:
            sharding.init_sharded_solver(e, dist)
        # warm the handle like a caller that solves more than one LP: buffers, (sharded) peer-memory arenas and
        # their IPC mappings are created by the first load and reused; the timed region starts from host arrays
        e.load(Ap.numpy(), bp.numpy(), cp.numpy())
        barrier()
        t0 = time.time()
        e.load(Ap.numpy(), bp.numpy(), cp.numpy())
        r = e.solve()
        t1 = time.time()
        e.close()
        tt = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        secs = float(tt.item())
        piv = r["stats"].pivots_phase1 + r["stats"].pivots_phase2
        line["e2e"] = {"value": piv / secs, "unit": UNIT, "h2d_bytes_per_step": int((n * m + n + m) * 8),
                       "d2h_bytes_per_step": int(n * 8 + 8 + m * 4) * world,
                       "step": f"one complete two-phase solve from pinned host arrays: status {r['status']}, "
                               f"{r['stats'].pivots_phase1}+{r['stats'].pivots_phase2} pivots in {secs:.3f} s "
                               f"(max over ranks; load {r['stats'].seconds_load:.3f} s)", "objective": r["objective"],
                       "per_phase": {"phase1_pivots_per_s": r["stats"].pivots_phase1 / max(r["stats"].seconds_phase1, 1e-9),
                                     "phase2_pivots_per_s": r["stats"].pivots_phase2 / max(r["stats"].seconds_phase2, 1e-9),
                                     "tableau_rows_streamed_both_phases": dims["rows_stored"]}}
    elif not a.no_e2e:
        line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "step": "skipped: the constraint matrix of this configuration does not fit in host memory; "
                               "the instance only ever exists on the devices"}
    if not a.no_cpu_baseline and rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline(a)

    # ---- supplementary: the same complete solve with skip_zero_rows (opt-in, value-exact): rows whose pivot-
    # constraint entry is exactly 0 are not streamed.  Reported apart so that `value`, `e2e` and `roofline` above
    # keep the full 2*R*m*8 bytes per pivot.
    if world == 1 and not a.no_e2e and not a.skip_zero_rows and n * m * 8 <= 8e9:
        with S.Solver(device=local, skip_zero_rows=True, update_variant=a.update_variant,
                      persistent={"auto": "auto", "persistent": True, "launches": False}[a.loop]) as z:
            torch.cuda.synchronize()
            t0 = time.time()
            z.load(Ap.numpy(), bp.numpy(), cp.numpy())
            rz = z.solve()
            t1 = time.time()
        pz = rz["stats"].pivots_phase1 + rz["stats"].pivots_phase2
        line["skip_zero_rows"] = {"e2e_pivots_per_s": pz / (t1 - t0), "seconds": t1 - t0,
                                  "rows_streamed_fraction": rz["stats"].rows_streamed / max(1, rz["stats"].rows_total),
                                  "same_pivot_sequence": int(rz["stats"].trace_hash) == int(r["stats"].trace_hash),
                                  "same_objective": rz["objective"] == r["objective"]}

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
            before = g.stats().seconds_phase1
            st_l, done_l = g.iterate(20)
            barrier()
            ms_l = torch.tensor([(g.stats().seconds_phase1 - before) * 1e3], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(ms_l, op=dist.ReduceOp.MAX)
            g.close()
            pps = done_l / (float(ms_l.item()) * 1e-3)
            slab = 2.0 * ldims["rows_stored"] * (lm // world) * elem
            line["large_config"] = {"workload": "random_65536_65536 (seed 6619136, [1,100]); tableau 65536x131072 fp64, "
                                                f"{ldims['rows_stored'] * lm * elem / 1e9:.1f} GB, constraint slabs x{world}",
                                    "pivots_per_s": pps, "pivots_timed": done_l,
                                    "per_gpu_stream_GBps": slab * pps / 1e9,
                                    "frac_of_measured_hbm_peak": slab * pps / 1e9 / peaks()[0]}
        except Exception as exc:  # e.g. not enough free HBM on a shared box
            line["large_config"] = {"skipped": str(exc)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse()
    sys.exit(run_reference(args) if args.impl == "reference" else run_b2s(args))
