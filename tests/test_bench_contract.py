"""The driver's bench contract on the committed bench lines of this round (profiles/r1_i_*.json, produced by bench.py on a
B200): every key the contract names is present and consistent.  CPU only -- guards the JSON shape, not the numbers."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_own_arm_line_has_the_contract_keys():
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for name, n in (("r1_i_bench_n1.json", 1), ("r1_i_bench_n2.json", 2)):
        d = _line(name)
        assert d["metric"] in baseline["metric"] and d["unit"] == "candidates/s"
        assert d["n_gpus"] == n and d["steps"] >= 1 and d["warmup"] >= 3
        assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert d["dtype"] == "int32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
        assert "l2" in d["config"]                                   # how L2 is defeated between timed iterations
        assert abs(d["value"] - d["config"]["candidates"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
        e = d["e2e"]
        assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        assert d["gpu_launches"] > 0
        c = d["clocks"]
        assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert "traffic" in r and r["kernel"]
        assert set(d["stages_ms"]) == {"sketch_cands", "seed_sort_cluster", "extend", "copy_rules_msa", "clean_judge", "compact"}
    d = _line("r1_i_bench_n1.json")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["unit"] == d["unit"] and cb["sample"]
    assert d["roofline"]["traffic"] and d["roofline"]["traffic"] > 0
    for k in ("per_type_candidates_per_s", "coarse_boundary", "tir_prefilter", "nonltr_prefilter", "known_te_masking", "homology_copy_search"):
        assert k in d, k
    assert set(d["per_type_candidates_per_s"]) >= {"tir", "helitron", "non_ltr"}


def test_reference_arm_line_has_the_contract_keys():
    d = _line("r1_i_bench_reference_arm.json")
    own = _line("r1_i_bench_n1.json")
    assert d["impl"] == "reference" and d["metric"] == own["metric"] and d["unit"] == own["unit"]
    assert d["higher_is_better"] is True and d["config"]["workload"] == own["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
