"""bench.py's CPU-runnable behaviour: the reference arm really runs (plain-C port on a bounded sample of the named workload)
and prints the contract's line; the roofline's algorithmic bytes are SURVEY 8(d)-(ii)'s formula; the multi-GPU sharding of
the C4 configuration (chunk -> rank, then rebalance) is a balanced partition."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_runs_and_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--small", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] in baseline["metric"] and d["unit"] == "candidates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "NOT the reference's Python" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_algorithmic_bytes_follow_survey_8d():
    sys.path.insert(0, ROOT)
    import bench
    L, flank, w = 1000, 50, 10
    exp = (L + (L + 2 * flank)) * 3 / 8 + 8 * (2 * L / (w + 1)) + 16 + -(-(L + 100) // 4)
    assert abs(float(bench.algo_bytes_pair(L)) - exp) < 1e-9
    assert 2400 < exp < 2600                                   # SURVEY: "L = 1 kbp => ~2.5 kB / pair"
    v = bench.algo_bytes_pair(np.array([200, 3000]))
    assert v.shape == (2,) and v[1] > v[0]


def test_chunk_sharding_is_a_balanced_partition():
    sys.path.insert(0, ROOT)
    from hite_b200 import dist as hd
    rng = np.random.default_rng(3)
    n = 4000
    lengths = [int(x) for x in rng.integers(100, 3000, n)]
    chunk = [int(x) for x in rng.choice(4, n, p=[0.55, 0.25, 0.15, 0.05])]
    for world in (1, 2, 4, 8):
        parts = [hd.shard_by_chunk(chunk, lengths, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(n))
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) <= 1.01 * sum(loads) / world
        # candidates stay on their chunk's rank unless the rebalance moved them: the lightest chunk's rank only receives
        if world == 4:
            assert {chunk[i] for i in parts[3]} >= {3}
