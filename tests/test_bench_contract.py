"""bench.py contract checks that need no GPU: the reference arm (the CPU oracle port timed on the host cores) runs here
at a reduced size and prints ONE JSON line with the keys the driver reads; the roofline helpers read the committed
profile summary."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--refs", "1000", "--sketch-size", "500", "--reads", "64", "--lineages", "2", "--cpu-sample", "3"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "reads/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    assert cb["all_cores"]["cores"] >= 1 and cb["all_cores"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("predict 64 reads vs 1000 x s=500")
    # the same `config` object as the GPU arm prints for this command line (the driver compares the two arms' dicts):
    # both come from bench.config_of, which holds nothing a run measures
    sys.path.insert(0, ROOT)
    import bench
    a = bench.parse(["--refs", "1000", "--sketch-size", "500", "--reads", "64", "--lineages", "2"])
    assert d["config"] == bench.config_of(a.config, 1000, 500, 64, a.top, a, 1)
    assert set(d["config"]) == {"workload", "refs", "sketch_size", "reads", "read_len", "k", "top", "lineages", "rows", "l2", "parallelism"}


def test_reference_arm_is_silent_on_other_ranks():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_roofline_helpers():
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench.ncu_traffic_bytes()
    assert t is not None and 3.2e9 < t < 3.3e9 and src.startswith("profiles/")   # one pass over the 3.2 GB matrix, no re-reads
    assert bench.workload_name("c3", 40000, 10000, 100000, 10).startswith("C3: predict 100,000 synthetic ONT reads")
    assert bench.workload_name("c4", 40000, 1000, 1000000, 5).startswith("C4: streaming predict + genotype consensus")


def test_consensus_calls_match_the_host_mirror():
    """bench.py's vectorised per-read consensus (C4) == api.consensus_value on every read, ties included."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from sketchy_b200.api import consensus_value
    rng = np.random.default_rng(3)
    table = rng.integers(0, 3, size=(50, 4)).astype(np.int32)
    idx = rng.integers(0, 50, size=(200, 5)).astype(np.uint32)
    calls = bench.consensus_calls(idx, table)
    for r in range(200):
        for c in range(4):
            assert str(calls[r, c]) == consensus_value([str(x) for x in table[idx[r].astype(np.int64), c]])
    peak, how = bench.measured_peak_gbs()
    assert peak > 1000 and how in ("measured", "fallback")
