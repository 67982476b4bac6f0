#!/usr/bin/env python
"""Device time of one hb_adjust step on the C2 workload for several hb_params.overlap settings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hite_b200
from workloads import synth

wl = synth.make_workload(genome_bp=100_000_000, n_contigs=10, n_families=1000, cands_per_family=10, copies=(10, 40),
                         length=(200, 2000), seed=1)
ctx = hite_b200.Context(0)
ctx.load_genome(wl.contig_seqs); ctx.build_index(15, 10)
flat = np.frombuffer(b"".join(wl.cand_seqs) + b"\0", dtype=np.uint8)
offs = np.zeros(len(wl.cand_seqs) + 1, dtype=np.uint64); offs[1:] = np.cumsum([len(c) for c in wl.cand_seqs])
for ov in [int(x) for x in (sys.argv[1:] or ["0", "2", "3", "4", "8"])]:
    for _ in range(3):
        ctx.adjust_raw((flat, offs), "tir", 1, overlap=ov)
    dev, wall = [], []
    for _ in range(8):
        t0 = time.perf_counter(); ctx.adjust_raw((flat, offs), "tir", 1, overlap=ov); wall.append(time.perf_counter() - t0)
        tm = ctx.timing(); dev.append(tm["kernel_ms"])
    print("overlap=%d  device %.2f ms  wall %.2f ms  stages %s launches %d" % (ov, np.mean(dev), np.mean(wall) * 1e3,
          " ".join("%.2f" % x for x in tm["k_ms"][:6]), tm["launches"]), flush=True)
