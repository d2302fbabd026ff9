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
