"""CPU test of bench.py's reference-arm fallback (the serial oracle port on the host cores) and of the JSON
contract keys the driver reads.  The GPU arms run on the B200 box only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_port_fallback_prints_contract_line():
    env = dict(os.environ, B2S_BENCH_FORCE_PORT="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--vars", "96",
                          "--constraints", "64", "--cpu-pivots", "40", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "pivots/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["vs_baseline"] is None


def test_non_zero_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_roofline_traffic_only_while_the_stamp_matches_the_kernel_sources(tmp_path, monkeypatch):
    """roofline.traffic comes from the committed ncu capture, but only for the workload it was taken on and only while the
    sha256 of the kernel sources stamped into profiles/update_kernel_traffic.json matches the tree."""
    b = _bench()
    stamp = json.load(open(os.path.join(ROOT, "profiles", "update_kernel_traffic.json")))
    got = b.traffic_from_profile(8192, 8192, stamp["skip_zero_rows"])
    assert got in (None, stamp["dram_bytes_per_launch"])              # None once the kernels changed after the capture
    assert b.traffic_from_profile(4096, 4096, stamp["skip_zero_rows"]) is None      # other workload
    assert b.traffic_from_profile(8192, 8192, not stamp["skip_zero_rows"]) is None  # other kernel mode
    # a tree whose kernel source differs from the stamped one: no traffic figure
    fake = tmp_path / "repo"
    (fake / "profiles").mkdir(parents=True)
    for f in stamp["kernel_sources"]:
        dst = fake / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        dst.write_bytes(open(os.path.join(ROOT, f), "rb").read())
    (fake / "profiles" / "update_kernel_traffic.json").write_text(json.dumps(stamp))
    monkeypatch.setattr(b, "ROOT", str(fake))
    assert b.traffic_from_profile(8192, 8192, stamp["skip_zero_rows"]) == stamp["dram_bytes_per_launch"]
    with open(fake / stamp["kernel_sources"][0], "ab") as fh:
        fh.write(b"\n// edited after the capture\n")
    assert b.traffic_from_profile(8192, 8192, stamp["skip_zero_rows"]) is None


def test_parity_object_flags_any_difference_from_the_fixture():
    b = _bench()

    class St:
        pivots_phase1, pivots_phase2, trace_hash = 1777, 76, 8360765708768715880
    fx = b.fixture(1024, 1024, 103424)
    r = {"status": 0, "objective": fx["objective"], "stats": St()}
    assert b.parity_of_solve(r, fx)["status"] == "ok"
    assert b.parity_of_solve(dict(r, objective=fx["objective"] * (1 + 1e-15)), fx)["status"] == "MISMATCH"
    St2 = type("St2", (), dict(pivots_phase1=1777, pivots_phase2=76, trace_hash=1))
    assert b.parity_of_solve(dict(r, stats=St2()), fx)["status"] == "MISMATCH"
    assert b.parity_of_solve(r, None)["status"] == "unpinned"
