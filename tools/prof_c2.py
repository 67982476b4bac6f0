#!/usr/bin/env python
"""One index build + `--rounds` adjust rounds on a named workload, for ncu (`-k regex:...` selects the kernels):
    ncu --set full --clock-control none -k regex:'pack_kernel|minimizer' -c 8 -o gpurun_out/x python tools/prof_c2.py --config C2"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import hite_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--rounds", type=int, default=1)
ap.add_argument("--small", action="store_true")
ap.add_argument("--coarse", action="store_true")
a = ap.parse_args()
wl, name = bench.workload_for(a.config, "weak", 1, a.small)
with hite_b200.Context(0) as ctx:
    ctx.load_genome(wl.contig_seqs, wl.contig_names)
    ctx.build_index(15, 10)
    flat = np.frombuffer(b"".join(wl.cand_seqs) + b"\0", dtype=np.uint8)
    offs = np.zeros(len(wl.cand_seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(c) for c in wl.cand_seqs])
    for _ in range(a.rounds):
        calls, _ = ctx.adjust_raw((flat, offs), "tir", 1)
        tm = ctx.timing()
    if a.coarse:
        ctx.coarse_boundary(threads=8)
    print(name, "is_te", int((calls["is_te"] != 0).sum()), "stages", [round(x, 3) for x in tm["k_ms"]], "kernel_ms", tm["kernel_ms"])
