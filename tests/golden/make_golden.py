#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference functions (imported read-only from
/root/reference/module/Util.py through oracle/ref_import.py) on seeded synthetic inputs.

Run in the build container only:   python tests/golden/make_golden.py
Outputs (committed):  tests/golden/judge_v5.json.gz
"""
import gzip
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_import  # noqa: E402
import msa_synth  # noqa: E402


def run_reference_v5(U, case, tmpdir):
    """remove_sparse_col_in_align_file (Util.py:10344) + judge_boundary_v5 (Util.py:9145)."""
    path = os.path.join(tmpdir, "a.maf.fa")
    with open(path, "w") as f:
        for i, row in enumerate(case["rows"]):
            f.write(">r%d\n%s\n" % (i, row))
    clean = U.remove_sparse_col_in_align_file(path)
    names, contigs = U.read_fasta(clean)
    clean_rows = [contigs[n] for n in names]
    trace = []
    orig = U.search_boundary_homo_v3

    def spy(*a, **k):
        r = orig(*a, **k)
        trace.append(r)
        return r

    U.search_boundary_homo_v3 = spy
    try:
        is_te, info, cons, copy_num = U.judge_boundary_v5(case["cand"], clean, 0, "tir", case["plant"], "cons")
    finally:
        U.search_boundary_homo_v3 = orig
    hs = trace[0] if len(trace) > 0 else -1
    he = trace[1] if len(trace) > 1 else -1
    return dict(clean_rows=clean_rows, is_te=bool(is_te), info=info, cons=cons, copy_num=copy_num,
                homo_start=hs, homo_end=he)


def run_reference_v6(U, case, tmpdir):
    """remove_sparse_col_in_align_file (Util.py:10344) + judge_boundary_v6 (Util.py:9821)."""
    path = os.path.join(tmpdir, "h.maf.fa")
    with open(path, "w") as f:
        for i, row in enumerate(case["rows"]):
            f.write(">r%d\n%s\n" % (i, row))
    clean = U.remove_sparse_col_in_align_file(path)
    names, contigs = U.read_fasta(clean)
    trace = []
    orig = U.search_boundary_homo_v4

    def spy(*a, **k):
        r = orig(*a, **k)
        trace.append(r)
        return r

    U.search_boundary_homo_v4 = spy
    try:
        is_te, info, cons, copy_num = U.judge_boundary_v6(case["cand"], clean, 0, "helitron", case["plant"], "cons")
    finally:
        U.search_boundary_homo_v4 = orig
    return dict(clean_rows=[contigs[n] for n in names], is_te=bool(is_te), info=info, cons=cons, copy_num=copy_num,
                v4=[[bool(a), int(b)] for a, b in trace])


def run_reference_v9(U, case, tmpdir):
    """remove_sparse_col_in_align_file (Util.py:10344) + judge_boundary_v9 (Util.py:9483)."""
    path = os.path.join(tmpdir, "n.maf.fa")
    with open(path, "w") as f:
        for i, row in enumerate(case["rows"]):
            f.write(">r%d\n%s\n" % (i, row))
    clean = U.remove_sparse_col_in_align_file(path)
    names, contigs = U.read_fasta(clean)
    is_te, info, cons, copy_num = U.judge_boundary_v9(case["cand"], clean, 0, "non_ltr", case["plant"], "cons")
    return dict(clean_rows=[contigs[n] for n in names], is_te=bool(is_te), info=info, cons=cons, copy_num=copy_num)


def run_reference_fmea(U, case, tmpdir):
    """get_longest_repeats_v4 (Util.py:4122) on a synthetic outfmt-6 file of one query segment."""
    path = os.path.join(tmpdir, "q.final.out")
    qn = "%s$%d" % case["query"]
    with open(path, "w") as f:
        for (si, qs, qe, ss, se) in case["hsps"]:
            sn = "%s$%d" % tuple(case["subjects"][si])
            f.write("\t".join([qn, sn, "95.0", str(abs(qe - qs) + 1), "0", "0", str(qs), str(qe), str(ss), str(se), "1e-50", "200"]) + "\n")
    out = U.get_longest_repeats_v4(path, case["skip_gap"], case["max_len"], 0)
    d = U.load_from_file(out)
    return list(d.keys())


def run_reference_copies(U, case, tmpdir):
    """get_query_copies (Util.py:6828) for one query named 'q' against subjects s0..sk."""
    subject_dict = {}
    for (si, qs, qe, ss, se), idn in zip(case["hsps"], case["idents"]):
        subject_dict.setdefault("s%d" % si, []).append((qs, qe, ss, se, idn))
    subj_path = os.path.join(tmpdir, "subj.fa")
    with open(subj_path, "w") as f:
        for i, ln in enumerate(case["subj_lens"]):
            f.write(">s%d\n%s\n" % (i, "A" * ln))
    res = U.get_query_copies([("q", subject_dict)], {"q": "A" * case["qlen"]}, subj_path, case["query_coverage"],
                             case["subject_coverage"], max_copy_num=case["max_copy_num"])
    return [[int(c[0][1:]), int(c[1]), int(c[2]), int(c[3]), c[4]] for c in res["q"]]


def write(path, out):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, len(out), "cases;", sum(r["expect"]["is_te"] for r in out), "TE;", os.path.getsize(path), "bytes")
    from collections import Counter
    print(Counter((r["scenario"], r["expect"]["is_te"], r["expect"]["info"]) for r in out))


def main():
    U = ref_import.load_util()
    with tempfile.TemporaryDirectory() as td:
        out = []
        for seed in range(1, 121):
            case = msa_synth.synth_case_helitron(seed)
            exp = run_reference_v6(U, case, td)
            out.append(dict(seed=seed, scenario=case["scenario"], cand=case["cand"], plant=case["plant"],
                            rows=case["rows"], truth=case["truth"], expect=exp))
        write(os.path.join(ROOT, "tests", "golden", "judge_v6.json.gz"), out)
        out = []
        for seed in range(1, 121):
            case = msa_synth.synth_case_nonltr(seed)
            exp = run_reference_v9(U, case, td)
            out.append(dict(seed=seed, scenario=case["scenario"], cand=case["cand"], plant=case["plant"],
                            rows=case["rows"], truth=case["truth"], expect=exp))
        write(os.path.join(ROOT, "tests", "golden", "judge_v9.json.gz"), out)
    with tempfile.TemporaryDirectory() as td:
        recs = []
        for seed in range(1, 201):
            case = msa_synth.synth_hsps(seed)
            case["expect"] = run_reference_fmea(U, case, td)
            case["seed"] = seed
            recs.append(case)
        path = os.path.join(ROOT, "tests", "golden", "fmea.json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(recs, separators=(",", ":")).encode())
        print("wrote", path, len(recs), "cases;", sum(len(r["expect"]) for r in recs), "intervals;", os.path.getsize(path), "bytes")
    with tempfile.TemporaryDirectory() as td:
        recs = []
        for seed in range(1, 201):
            case = msa_synth.synth_copy_hsps(seed)
            case["expect"] = run_reference_copies(U, case, td)
            case["seed"] = seed
            recs.append(case)
        path = os.path.join(ROOT, "tests", "golden", "query_copies.json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(recs, separators=(",", ":")).encode())
        print("wrote", path, len(recs), "cases;", sum(len(r["expect"]) for r in recs), "copies;", os.path.getsize(path), "bytes")
    out = []
    with tempfile.TemporaryDirectory() as td:
        seeds = list(range(1, 161))
        for seed in seeds:
            case = msa_synth.synth_case(seed)
            exp = run_reference_v5(U, case, td)
            rec = dict(seed=seed, scenario=case["scenario"], cand=case["cand"], plant=case["plant"],
                       rows=case["rows"], truth=case["truth"], expect=exp)
            out.append(rec)
    path = os.path.join(ROOT, "tests", "golden", "judge_v5.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    n_te = sum(r["expect"]["is_te"] for r in out)
    print("wrote", path, len(out), "cases;", n_te, "TE;", os.path.getsize(path), "bytes")
    from collections import Counter
    print(Counter((r["scenario"], r["expect"]["is_te"], r["expect"]["info"]) for r in out))


if __name__ == "__main__":
    main()
