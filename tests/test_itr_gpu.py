"""GPU: hb_itr_search (HB-itr) through the C-ABI == oracle/hbo_tsd.c::hbo_itr_search on the records of the binary's golden
file and on long / odd records; run_itrsearch writes the files the reference's callers parse."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("found", "len5", "len3", "score", "n_match", "n_mismatch", "n_gap")


def test_itr_search_bit_exact_vs_oracle(ctx):
    with gzip.open(os.path.join(ROOT, "tests", "golden", "itrsearch.json.gz"), "rt") as f:
        seqs = [c["seq"] for c in json.load(f)["cases"]]
    rng = np.random.default_rng(9)
    B = np.frombuffer(b"ACGTN", dtype=np.uint8)
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    for L, k in ((3000, 700), (1001, 300), (2, 1), (1, 0), (15, 7), (80, 40)):       # long windows, odd lengths, tiny records
        t = B[rng.integers(0, 4, k)].tobytes()
        body = B[rng.integers(0, 5, max(0, L - 2 * k))].tobytes()
        seqs.append((t + body + t.translate(comp)[::-1]).decode())
    seqs.append("acgtgacgtacg" + "T" * 50 + "cgtacgtcacgt")
    got = ctx.itr_search(seqs)
    for s, g in zip(seqs, got):
        exp = O.itr_search(s)
        assert {f: int(g[f]) for f in FIELDS} == exp, (len(s), exp)
    assert int(got["found"].sum()) > 500
    assert len(ctx.itr_search([])) == 0


def test_run_itrsearch_writes_the_references_files(ctx, tmp_path, monkeypatch):
    from hite_b200 import boundary, tir_prefilter as TP
    monkeypatch.setattr(boundary, "get_context", lambda device=None, role="reference": ctx)
    comp = str.maketrans("ACGT", "TGCA")
    tir = "GTAGCCAGAAGGCTGCAACTC"
    recs = {"withTIR-C_0-tsd_TA-distance_0": tir + "ACGTTGCAAGGTCCAT" * 5 + tir.translate(comp)[::-1],
            "noTIR-C_1-tsd_TAA-distance_3": "ACGGTCATTGACCGTAAGCTAGGCTTAGCATCGGATCAGCTAGGATCGACGGATTTACGACT"}
    inp = str(tmp_path / "0.fa")
    with open(inp, "w") as f:
        for n, s in recs.items():
            f.write(">%s\n%s\n" % (n, s))
    out, log = TP.run_itrsearch("unused", inp, str(tmp_path))
    assert out == inp + ".itr" and log == inp + ".log"
    lines = open(out).read().splitlines()
    L = len(recs["withTIR-C_0-tsd_TA-distance_0"])
    e = O.itr_search(recs["withTIR-C_0-tsd_TA-distance_0"])
    assert e["found"] and e["len5"] >= 21
    assert lines[0] == ">withTIR-C_0-tsd_TA-distance_0 ITR(1,%d)..(%d,%d) Length itr=%d" % (e["len5"], L, L - e["len3"] + 1, e["len5"] - 1)
    assert "".join(lines[1:]) == recs["withTIR-C_0-tsd_TA-distance_0"] and all(len(x) <= 60 for x in lines[1:])
    assert "noTIR" not in open(out).read() and "no alignment!" in open(log).read()
    assert int(lines[0].split("Length itr=")[1]) == e["len5"] - 1                # what module/Util.py:6586-6593 parses
