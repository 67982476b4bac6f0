"""GPU: coarse-boundary all-vs-all (self-join of the minimizer index -> HSPs) + FMEA through the C-ABI vs the
plain-C restatement (oracle/hbo_hsp.c + hbo_fmea.c); planted copies must be found."""
import numpy as np
import pytest

from hite_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_coarse_boundary_bit_exact_and_finds_planted_copies(ctx):
    wl = synth.make_workload(genome_bp=3_000_000, n_contigs=2, n_families=10, cands_per_family=1, copies=(6, 14),
                             length=(300, 1500), seed=31, sub=(0.0, 0.04), indel=0.003)
    lens = [len(c) for c in wl.contig_seqs]
    G = oracle.Genome(wl.contig_seqs)
    ctx.load_genome(wl.contig_seqs)
    ctx.build_index(15, 10)
    for seg_len, occ_cap in ((1_000_000, 64), (400_000, 8)):
        exp_recs, exp_segs, exp_iv = G.coarse_boundary(lens, seg_len=seg_len, skip_gap=2000, max_len=30000, occ_cap=occ_cap)
        iv = ctx.coarse_boundary(seg_len=seg_len, skip_gap=2000, max_len=30000, occ_cap=occ_cap, threads=4)
        hsps, q_off, segs = ctx.last_hsps()
        assert [(int(s["chr"]), int(s["start"])) for s in segs] == exp_segs
        # GPU HSPs are grouped by query segment (stable); the oracle list is in sorted-key order
        got = []
        for g in range(len(segs)):
            for h in hsps[int(q_off[g]):int(q_off[g + 1])]:
                got.append((g, int(h["subj"]), int(h["qs"]), int(h["qe"]), int(h["ss"]), int(h["se"])))
        order = np.argsort(exp_recs[:, 0], kind="stable")
        exp = [tuple(int(x) for x in exp_recs[i]) for i in order]
        assert got == exp
        assert [(int(o["chr"]), int(o["start"]), int(o["end"])) for o in iv] == exp_iv
    # planted truth (coarse: within 30 bp at both ends) for the default segmentation
    iv = ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64)
    found = {(int(o["chr"]), int(o["start"]) // 30, int(o["end"]) // 30) for o in iv}
    starts = {}
    for o in iv:
        starts.setdefault(int(o["chr"]), []).append((int(o["start"]), int(o["end"])))
    hit = tot = 0
    for name in wl.cand_names:           # candidate names are exact or near-exact copy loci
        c, rng = name.split(":")
        a, b = (int(x) for x in rng.split("-"))
        ci = wl.contig_names.index(c)
        tot += 1
        hit += any(abs(s - a) <= 40 and abs(e - b) <= 40 for s, e in starts.get(ci, []))
    assert hit >= 0.6 * tot, (hit, tot)
