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
ap.add_argument("--shard-of", type=int, default=1, help="process only rank 0's shard of this many (what one GPU of N sees)")
a = ap.parse_args()
wl, name = bench.workload_for(a.config, "weak", 1, a.small)
with hite_b200.Context(0) as ctx:
    ctx.load_genome(wl.contig_seqs, wl.contig_names)
    ctx.build_index(15, 10)
    from hite_b200 import dist as hbdist
    mine = hbdist.shard_indices(len(wl.cand_seqs), 0, a.shard_of, [len(c) for c in wl.cand_seqs])
    for kind in sorted({bench.te_type_of(t) for t in wl.truth}):
        cands = [wl.cand_seqs[i] for i in mine if bench.te_type_of(wl.truth[i]) == kind]
        if not cands:
            continue
        flat = np.frombuffer(b"".join(cands) + b"\0", dtype=np.uint8)
        offs = np.zeros(len(cands) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(c) for c in cands])
        for _ in range(a.rounds):
            calls, _ = ctx.adjust_raw((flat, offs), kind, 1)
            tm = ctx.timing()
        print(kind, len(cands), "stages", [round(x, 3) for x in tm["k_ms"]], "kernel_ms", tm["kernel_ms"])
    if a.coarse:
        ctx.coarse_boundary(threads=8)
    print(name, "is_te", int((calls["is_te"] != 0).sum()), "stages", [round(x, 3) for x in tm["k_ms"]], "kernel_ms", tm["kernel_ms"])
