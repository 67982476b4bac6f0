#!/usr/bin/env python
"""Recall against planted truth over divergence / indel / nesting settings: for every setting a
12 Mbp genome with 60 families (copies U[10,30]) is generated (workloads/synth.py), every candidate goes through hb_adjust
and the genome through hb_coarse_boundary.  Reported per setting:

  cand_te      fraction of candidates judged a TE
  ends_ok      fraction whose emitted consensus starts and ends like the planted family consensus (<= 2 mismatches in 20 bp)
  copy_recall  of the planted INTACT copies of a candidate's family that the reference's own coverage rule admits (a copy
               more than 1 / 0.95 times as long as the candidate fails (M+I)/(M+D) >= 0.95, module/Util.py:8013-8022, whatever the
               aligner), the fraction among the kept copies (>= 90 % reciprocal overlap)
  coarse_fam   fraction of families with a coarse interval within 40 bp of a planted copy at both ends

Writes a markdown table (stdout and --out).  Runs on the GPU box: python tools/recall_sweep.py --out gpurun_out/recall.md"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import hite_b200  # noqa: E402
from workloads import synth  # noqa: E402

COMP = bytes.maketrans(b"ACGT", b"TGCA")


def mism(a, b):
    n = min(len(a), len(b))
    return sum(1 for x, y in zip(a[:n], b[:n]) if x != y) + abs(len(a) - len(b))


def run_setting(ctx, name, **kw):
    args = dict(genome_bp=12_000_000, n_contigs=3, n_candidates=600, copies=("uniform", 10, 30), cands_per_family=10, seed=11,
                lengths={"tir": (300, 2500)})
    args.update(kw)
    wl = synth.make_mixed_workload(**args)
    ctx.load_genome(wl.contig_seqs, wl.contig_names)
    ctx.build_index(15, 10)
    kinds = sorted({t["kind"] for t in wl.truth})
    te = {"tir": "tir", "ltr": "tir", "helitron": "helitron"}
    n_te = n_ends = 0
    cr_num = cr_den = 0
    by_fam = {}
    for (f, c, a, b, s) in wl.copies_truth:
        by_fam.setdefault(f, []).append((c, a, b))
    for kind in kinds:
        idx = [i for i, t in enumerate(wl.truth) if t["kind"] == kind]
        res = ctx.adjust([wl.cand_seqs[i] for i in idx], te[kind], 1)
        copies = ctx.last_copies()
        kept = copies[copies["keep"] != 0]
        for j, i in enumerate(idx):
            t, r = wl.truth[i], res[j]
            n_te += r["is_te"]
            if r["is_te"]:
                cons = r["cons"].encode()
                fam = t["consensus"]
                rc = fam.translate(COMP)[::-1]
                ok = any(mism(cons[:20], x[:20]) <= 2 and mism(cons[-20:], x[-20:]) <= 2 for x in (fam, rc))
                n_ends += ok
            mine = kept[kept["cand"] == j]
            spans = {}
            for m in mine:
                spans.setdefault(int(m["contig"]), []).append((int(m["start1"]) - 1, int(m["end1"])))
            for (c, a, b) in by_fam[t["family"]]:
                if (b - a) * 0.95 > len(wl.cand_seqs[i]):
                    continue
                cr_den += 1
                for (x, y) in spans.get(c, []):
                    ov = min(b, y) - max(a, x)
                    if ov >= 0.9 * (b - a) and ov >= 0.9 * (y - x):
                        cr_num += 1
                        break
    iv = ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64, threads=8)
    starts = {}
    for o in iv:
        starts.setdefault(int(o["chr"]), []).append((int(o["start"]), int(o["end"])))
    fam_hit = 0
    for f, lst in by_fam.items():
        fam_hit += any(abs(s - a) <= 40 and abs(e - b) <= 40 for (c, a, b) in lst for (s, e) in starts.get(c, []))
    n = len(wl.cand_seqs)
    return "| %s | %d | %.3f | %.3f | %.3f | %.3f |" % (name, n, n_te / n, n_ends / max(1, n), cr_num / max(1, cr_den), fam_hit / max(1, len(by_fam)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = ["| setting | candidates | cand_te | ends_ok | copy_recall | coarse_fam |", "|---|---|---|---|---|---|"]
    with hite_b200.Context(0) as ctx:
        for sub in (0.0, 0.05, 0.10, 0.15, 0.20):
            rows.append(run_setting(ctx, "sub %.0f %%, indel 0.5 %%" % (100 * sub), sub=(sub, sub), indel=(0.005, 0.005)))
        for ind in (0.01, 0.02, 0.04):
            rows.append(run_setting(ctx, "sub 5 %%, indel %.0f %%" % (100 * ind), sub=(0.05, 0.05), indel=(ind, ind)))
        for bl in (20, 50, 100, 300):
            rows.append(run_setting(ctx, "sub 5 %%, one %d bp block indel per copy" % bl, sub=(0.05, 0.05), indel=(0.002, 0.002),
                                    block=(1, bl, bl)))
        rows.append(run_setting(ctx, "sub 5 %, 30 % of copies nested into other copies", sub=(0.05, 0.05), indel=(0.005, 0.005),
                                nested_frac=0.3))
        for drift in (0.01, 0.03):
            rows.append(run_setting(ctx, "LTR-like 4-10 kbp, sub 3 %%, indel %.0f %% (net drift)" % (100 * drift), mix=(("ltr", 1.0),),
                                    genome_bp=24_000_000, n_candidates=200, copies=("uniform", 8, 16), sub=(0.03, 0.03), indel=(drift, drift)))
    text = "\n".join(rows)
    print(text)
    if a.out:
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
