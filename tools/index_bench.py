#!/usr/bin/env python
"""Kernel (i) alone: pack + (k, w)-minimizer index build on random genomes of the named sizes, second (warm) build reported.
Algorithmic bytes per SURVEY 8(d)-(i): 0.375 G + 24 M (M = minimizers); peak = MEASURED_PEAKS.json hbm_gbs.

    python tools/index_bench.py --mbp 100 800 2300 --out gpurun_out/index_bench.json"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hite_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mbp", type=int, nargs="+", default=[100, 800])
ap.add_argument("--out", default=None)
a = ap.parse_args()
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = []
B = np.frombuffer(b"ACGT", dtype=np.uint8)
for mbp in a.mbp:
    rng = np.random.default_rng(mbp)
    n_contigs = 10
    contigs = [B[rng.integers(0, 4, mbp * 1_000_000 // n_contigs, dtype=np.uint8)] for _ in range(n_contigs)]
    with hite_b200.Context(0) as ctx:
        best = None
        for it in range(3):
            tl = ctx.load_genome(contigs)
            ti = ctx.build_index(15, 10)
            info = ctx.genome_info()
            rec = dict(mbp=mbp, pack_ms=tl["kernel_ms"], h2d_ms=tl["h2d_ms"], sketch_ms=ti["k_ms"][0], sort_ms=ti["k_ms"][1], bucket_ms=ti["k_ms"][2],
                       index_ms=ti["kernel_ms"], n_minimizers=int(info["n_minimizers"]))
            if it and (best is None or rec["index_ms"] + rec["pack_ms"] < best["index_ms"] + best["pack_ms"]):
                best = rec
        G, M = info["total_padded"], info["n_minimizers"]
        algo = 1.0 * G + 0.375 * G + 0.375 * G + 24.0 * M            # ASCII in + packed out (pack), packed in + keys + one binning pass (index)
        best["algorithmic_GB"] = algo / 1e9
        best["GBps"] = algo / 1e9 / ((best["index_ms"] + best["pack_ms"]) * 1e-3)
        best["frac_of_measured_hbm"] = best["GBps"] / peak
        best["index_only_GBps"] = (0.375 * G + 24.0 * M) / 1e9 / (best["index_ms"] * 1e-3)
        best["Gbp_per_s"] = mbp / 1e3 / ((best["index_ms"] + best["pack_ms"]) * 1e-3)
        rows.append(best)
        print(json.dumps(best), flush=True)
    del contigs
if a.out:
    json.dump(dict(peak_gbs=peak, rows=rows), open(a.out, "w"), indent=1)
