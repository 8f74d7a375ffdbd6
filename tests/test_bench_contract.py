"""bench.py's host-side contract, checked without a GPU: the committed ncu capture belongs to the kernels in the tree (a capture
of other kernels would be reported as stale, not as traffic), the film digests the timed legs are checked against exist, the
CPU sample is a whole number of tiles per worker, and the reference arm prints the line the driver expects."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_ncu_capture_is_reported_only_for_the_kernels_it_was_taken_from(monkeypatch):
    """A capture whose stamp equals the hash of today's kernel sources is reported; any other capture yields None and says
    'stale' — so the line can never carry the traffic of kernels that are not the ones running."""
    with open(os.path.join(ROOT, "profiles", "r02", "traffic.json")) as f:
        stored = json.load(f)
    for key in ("c2", "c5"):
        monkeypatch.setattr(bench, "kernel_source_hash", lambda key=key: stored[key]["kernel_hash"])
        t = bench.ncu_traffic(key)
        assert t["traffic_state"] == "current" and t["traffic_commit"] == stored[key]["commit"]
        assert t["dram_bytes_per_launch"] > 0 and t["l2_bytes_per_launch"] > t["dram_bytes_per_launch"]
        monkeypatch.setattr(bench, "kernel_source_hash", lambda: "0" * 16)
        t = bench.ncu_traffic(key)
        assert t["traffic_state"].startswith("stale") and t["dram_bytes_per_launch"] is None and t["l2_bytes_per_launch"] is None


def test_kernel_source_hash_ignores_host_only_sources(tmp_path, monkeypatch):
    before = bench.kernel_source_hash()
    src = os.path.join(ROOT, "yuki_b200", "csrc")
    names = [n for n in os.listdir(src) if n.endswith(".cuh") or n in ("render.cu", "yk_libm.h", "yk_fastdiv.h")]
    assert "wf_trace.cuh" in names and "render.cu" in names and "multi.inl" not in names and "host_pbrt.cpp" not in names
    assert before == bench.kernel_source_hash() and len(before) == 16


def test_stored_film_digests_cover_the_timed_legs():
    with open(bench.DIGESTS) as f:
        stored = json.load(f)
    assert "c2_1024spp" in stored and bench.C5_DIGEST in stored
    assert bench.check_digest("c2_1024spp", stored["c2_1024spp"], False) == "ok"
    assert bench.check_digest("no_such_workload", "0" * 32, False) == "absent"
    try:
        bench.check_digest("c2_1024spp", "0" * 32, False)
    except SystemExit as e:
        assert "film digest" in str(e)
    else:
        raise AssertionError("a wrong film must stop the bench")


def test_reference_arm_prints_the_drivers_line():
    """`bench.py --impl reference` at 4 spp (a fraction of a second of CPU work): one JSON line on stdout with impl / cpu_baseline / e2e."""
    env = dict(os.environ)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--spp-side", "2"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "Msamples/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert np.isfinite(line["ms_per_step"])
