#!/usr/bin/env python
"""One Helitron / non-LTR adjust workload (the per-type leg of bench.py) for profiling: python tools/type_probe.py helitron"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hite_b200
from workloads import synth

kind = sys.argv[1] if len(sys.argv) > 1 else "helitron"
wk = synth.make_workload(genome_bp=30_000_000, n_contigs=3, n_families=300, cands_per_family=10, copies=(10, 40),
                         length=(200, 2000), seed=2, kind=kind)
ctx = hite_b200.Context(0)
ctx.load_genome(wk.contig_seqs); ctx.build_index(15, 10)
fk = np.frombuffer(b"".join(wk.cand_seqs) + b"\0", dtype=np.uint8)
ok_ = np.zeros(len(wk.cand_seqs) + 1, dtype=np.uint64); ok_[1:] = np.cumsum([len(c) for c in wk.cand_seqs])
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    ctx.adjust_raw((fk, ok_), kind, 1)
    tm = ctx.timing()
    print(kind, "device %.2f ms" % tm["kernel_ms"], " ".join("%.2f" % x for x in tm["k_ms"][:6]), flush=True)
