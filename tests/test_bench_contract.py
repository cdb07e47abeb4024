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
                        "--warmup", "0", "--ref-sample-clusters", "2", "--views", "4"],
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
    # with the reference tree mounted (or staged under oracle/_ref by oracle/make_ref.py) the arm must run
    # the unmodified reference loop, not the port
    from oracle import ref_harness as rh
    if rh.available():
        assert cb["kind"] == "reference" and "UNMODIFIED reference" in cb["sample"]
    assert d["config"]["sample_clusters_per_step"] == 2 and d["gpu_launches"] == 0


def test_reference_staging_recipe_copies_only_what_the_path_imports(tmp_path, monkeypatch):
    """oracle/make_ref.py: byte-for-byte copies of the reference's src/ and clip packages into a git-ignored
    directory (never into the history), enough for ref_harness to import the hot path from the copy alone."""
    import filecmp
    from oracle import make_ref, ref_harness as rh
    if not os.path.isdir(make_ref.SOURCE_ROOT):
        import pytest
        pytest.skip("reference tree not mounted")
    monkeypatch.setattr(make_ref, "DEST", str(tmp_path / "_ref"))
    n = make_ref.stage()
    assert n >= 20
    for rel in ("src/utils/mv_utils.py", "src/utils/clip_utils.py", "src/vilgod/lidar_frame.py",
                "third_party/CLIP/clip/model.py", "third_party/CLIP/clip/bpe_simple_vocab_16e6.txt.gz"):
        assert filecmp.cmp(os.path.join(make_ref.SOURCE_ROOT, rel), str(tmp_path / "_ref" / rel), shallow=False)
    ignored = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in ignored
    gpurunignore = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(gpurunignore) or "oracle/_ref" not in open(gpurunignore).read()


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120,
                       cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
