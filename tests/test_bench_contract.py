"""bench.py's reference arm runs on the host cores, so its side of the JSON-line contract can be checked
without a GPU: exactly one line on stdout, the keys the driver reads, `e2e` repeating the line's own
value with zero copy bytes, and a `cpu_baseline` that says what was timed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample-clusters", "2", "--views", "4"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "clusters_per_second" and d["unit"] == "clusters/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 1
    assert d["value"] > 0 and "workload" in d["config"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["unit"] == d["unit"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120,
                       cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
