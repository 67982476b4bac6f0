#!/usr/bin/env python
"""The kernels of the "next" rows on C2-sized inputs, for timing / ncu:
tsd_kernel, polya_tsd_kernel, qhsp_kernel (library vs genome), mask_fill_kernel, export_kernel.
    python tools/next_probe.py [genome_Mbp] [iterations]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hite_b200
from hite_b200 import other, masking
from workloads import synth

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 30
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = synth.make_workload(genome_bp=mbp * 1_000_000, n_contigs=3, n_families=10 * mbp, cands_per_family=10, copies=(10, 40),
                         length=(200, 2000), seed=1)
ctx = hite_b200.Context(0)
cands = [c for c in wl.cand_seqs if len(c) > 100]
fams = {}
for t in wl.truth:
    fams.setdefault("fam%d#DNA/TIR" % t["family"], t["consensus"].decode())
names = list(fams)
for it in range(iters):
    ctx.tsd_candidates(cands, [51] * len(cands), [len(c) - 50 for c in cands], 50, 1)
    t_tsd = ctx.timing()["kernel_ms"]
    ctx.nonltr_candidates(cands, 50)
    t_pa = ctx.timing()["kernel_ms"]
    ctx.load_genome(wl.contig_seqs); ctx.build_index(15, 10); ctx.contig_names = list(wl.contig_names)
    rec = other.query_hsp_records(ctx, names, fams)
    tq = ctx.timing()
    iv = masking.intervals_from_hsp_records(rec, names, fams)
    fresh = ctx.mask_genome([i[1] for i in iv], [i[2] for i in iv], [i[3] for i in iv])
    t_fill = ctx.timing()["kernel_ms"]
    ctx.export_contig(0, len(wl.contig_seqs[0]))
    t_exp = ctx.timing()["kernel_ms"]
    print("n=%d  tsd %.3f ms  polya %.3f ms | library %d seqs: hsps %d  sketch %.2f seed %.2f extend %.2f qhsp %.3f ms | "
          "mask_fill %d intervals %.3f ms (%d bp)  export %.1f Mbp %.3f ms" % (
              len(cands), t_tsd, t_pa, len(names), len(rec), tq["k_ms"][0], tq["k_ms"][1], tq["k_ms"][2], tq["k_ms"][3],
              len(iv), t_fill, fresh, len(wl.contig_seqs[0]) / 1e6, t_exp), flush=True)
