#!/usr/bin/env python
"""Baseline (A) of SURVEY.md section 8(d) / BASELINE.md section 3: the REFERENCE's own Python code for the part of the path that
lives in its repository -- remove_sparse_col_in_align_file (module/Util.py:10344) + judge_boundary_v5 (:9145) per candidate,
and get_longest_repeats_v4 (:4122, FMEA) per BLAST result file -- run under a ProcessPoolExecutor exactly as
flank_region_align_v5 (:8141) and process_blast_alignments (:4773) do, fed the aligned matrices / HSP lists the CPU
restatement (oracle/) produces for a sample of the C2 workload.  The reference's aligners (minimap2, mafft, blastn) are not
installed, so their time is NOT in these numbers: this is a lower bound of the reference's cost per candidate for the
post-alignment half only.  Needs /root/reference (build container); writes one JSON object.

    python tools/baseline_a.py --candidates 512 --workers 4 8 --out profiles/r2_baseline_A.json"""
import argparse
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ProcessPoolExecutor, as_completed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

U = None


def _init():
    global U
    U = ref_import.load_util()


def judge_one(task):
    """what the tail of is_TE_from_align_file does after mafft (module/Util.py:10419-10437)"""
    cur_seq, maf = task
    if U is None:
        _init()
    align_file = U.remove_sparse_col_in_align_file(maf)
    is_te, info, cons, copy_num = U.judge_boundary_v5(cur_seq, align_file, 0, "tir", 1, "cons")
    return bool(is_te), len(cons) if is_te else 0, int(copy_num) if is_te else 0


def fmea_one(task):
    path, skip_gap, max_len = task
    if U is None:
        _init()
    return len(U.get_longest_repeats_v4(path, skip_gap, max_len, 0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=512)
    ap.add_argument("--workers", type=int, nargs="+", default=[4, os.cpu_count() or 8])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from workloads import synth
    import bench
    # a 1/10 slice of C2 (same family / copy-number / divergence distributions) keeps the CPU restatement's part short
    wl = synth.make_mixed_workload(10_000_000, 10, max(a.candidates, 64), seed=1)
    G = O.Genome(wl.contig_seqs)
    G.set_tie_rule(1, wl.contig_names)
    work = tempfile.mkdtemp(prefix="baselineA_")
    tasks, rows_total, cells = [], 0, 0
    t0 = time.perf_counter()
    for ci, seq in enumerate(wl.cand_seqs[:a.candidates]):
        cps, rows = G.candidate_msa(seq)
        if not rows:
            continue
        p = os.path.join(work, "c%d.blast.bed.fa.maf.fa" % ci)
        kept = [c for c in cps if c["keep"]]
        with open(p, "w") as f:
            for c, r in zip(kept, rows):
                f.write(">%s:%d-%d(%s)\n%s\n" % (wl.contig_names[c["contig"]], c["start1"], c["end1"], "-" if c["rel"] else "+", r))
        tasks.append((seq.decode(), p))
        rows_total += len(rows); cells += len(rows) * len(rows[0])
    t_align = time.perf_counter() - t0
    res = {"what": "reference Python (imported module/Util.py) on the post-alignment half of the path; aligner time excluded",
           "host_cpus": os.cpu_count(), "candidates": len(tasks), "mean_rows": rows_total / max(1, len(tasks)),
           "mean_cells": cells / max(1, len(tasks)), "workload": "10 Mbp slice with C2's distributions (workloads/synth.py seed 1)",
           "cpu_restatement_align_s_1thread": t_align, "judge": {}, "fmea": {}}
    for w in a.workers:
        t0 = time.perf_counter()
        n_te = 0
        with ProcessPoolExecutor(max_workers=w, initializer=_init) as ex:
            futs = [ex.submit(judge_one, t) for t in tasks]
            for fu in as_completed(futs):
                n_te += fu.result()[0]
        dt = time.perf_counter() - t0
        res["judge"]["workers_%d" % w] = {"seconds": dt, "candidates_per_s": len(tasks) / dt, "is_te": n_te}
    # FMEA: the oracle's HSPs of the same genome as outfmt-6 files, one per query segment (:4763)
    lens = [len(c) for c in wl.contig_seqs]
    hsps, segs, iv = G.coarse_boundary(lens, seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64)
    seg_name = ["%s$%d" % (wl.contig_names[c], o) for c, o in segs]
    files = {}
    for q, s, qs, qe, ss, se in hsps.tolist():
        files.setdefault(q, []).append("%s\t%s\t100.0\t%d\t0\t0\t%d\t%d\t%d\t%d\t0.0\t100\n" % (seg_name[q], seg_name[s], qe - qs + 1, qs, qe, ss, se))
    ftasks = []
    for q, lines in files.items():
        p = os.path.join(work, "q%d.final.out" % q)
        open(p, "w").writelines(lines)
        ftasks.append((p, 2000, 30000))
    for w in a.workers:
        t0 = time.perf_counter()
        n_iv = 0
        with ProcessPoolExecutor(max_workers=w, initializer=_init) as ex:
            for fu in as_completed([ex.submit(fmea_one, t) for t in ftasks]):
                n_iv += fu.result()
        dt = time.perf_counter() - t0
        res["fmea"]["workers_%d" % w] = {"seconds": dt, "Mbp_per_s": wl.genome_bp / 1e6 / dt, "hsps": int(len(hsps)), "intervals": n_iv}
    G.close()
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
