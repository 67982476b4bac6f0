"""GPU parity: library-vs-genome HSPs through the C-ABI (hb_query_hsps) vs the C oracle (bit-exact records), the host
mirror of multi_process_align_and_get_copies (reference module/Util.py:8382) vs oracle HSPs chained by the pinned
get_query_copies restatement, and planted truth."""
import os

import numpy as np
import pytest

import other_synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FIELDS = ("rank", "contig", "qs", "qe", "ss", "se", "n_match", "aln_len", "score")


@pytest.fixture(scope="module")
def wl():
    return other_synth.make(seed=5)


@pytest.fixture(scope="module")
def gctx(wl):
    import hite_b200
    c = hite_b200.Context(0)
    c.load_genome(wl[1])
    c.build_index(15, 10)
    c.contig_names = wl[0]
    yield c
    c.close()


def test_query_hsps_match_oracle(gctx, wl):
    names, contigs, lib_names, lib, truth = wl
    recs = gctx.query_hsps(lib)
    assert len(recs) > len(truth) // 2
    assert np.all(np.diff(recs["query"].astype(np.int64)) >= 0)
    G = O.Genome(contigs)
    n = 0
    for qi, q in enumerate(lib):
        mine = [tuple(int(r[f]) for f in FIELDS) for r in recs[recs["query"] == qi]]
        exp = [tuple(h[f] for f in FIELDS) for h in G.query_hsps(q)]
        assert mine == exp, (qi, mine[:3], exp[:3])
        n += len(exp)
    assert n == len(recs)
    tm = gctx.timing()
    assert tm["launches"] > 10 and tm["k_ms"][6] >= len(recs)
    G.close()


def test_query_hsps_edge_inputs(gctx, wl):
    assert len(gctx.query_hsps([])) == 0
    # queries without a single seed, shortest legal query, N runs, lower case
    lib = wl[3]
    odd = [b"ACG", b"N" * 500, b"A" * 300, lib[0].lower(), lib[1][:40], lib[2][:700] + b"N" * 30 + lib[2][700:]]
    recs = gctx.query_hsps(odd)
    G = O.Genome(wl[1])
    for qi, q in enumerate(odd):
        mine = [tuple(int(r[f]) for f in FIELDS) for r in recs[recs["query"] == qi]]
        assert mine == [tuple(h[f] for f in FIELDS) for h in G.query_hsps(q.upper())], qi
    assert len(recs[recs["query"] == 3]) > 0                      # lower case is read like upper case (read_fasta upper-cases)
    G.close()
    with pytest.raises(Exception):
        gctx.query_hsps([b"AC"])
    with pytest.raises(Exception):
        gctx.query_hsps([b"A" * 32001])


def test_multi_process_align_and_get_copies_mirror(gctx, wl, tmp_path, monkeypatch):
    """Same return shape as the reference; equals oracle HSPs chained by the (pinned) get_query_copies restatement; finds the
    planted full-length and internally deleted copies; 5'-truncated copies below 95 % query coverage stay out."""
    from hite_b200 import boundary, other
    names, contigs, lib_names, lib, truth = wl
    gpath, lpath = str(tmp_path / "genome.fa"), str(tmp_path / "non_LTR.lib")
    with open(gpath, "w") as f:
        for n, s in zip(names, contigs):
            f.write(">%s\n%s\n" % (n, s.decode()))
    with open(lpath, "w") as f:
        for n, s in zip(lib_names, lib):
            f.write(">%s some description\n" % n)
            for o in range(0, len(s), 70):                         # multi-line records, like the shipped library
                f.write(s[o:o + 70].decode() + "\n")
    monkeypatch.setattr(boundary, "_CTX", {})
    monkeypatch.setattr(boundary, "_GENOME", {})
    blast_dir = str(tmp_path / "other_TE")
    all_copies = other.multi_process_align_and_get_copies(lpath, gpath, blast_dir, "non_ltr", 4, is_removed_dir=True,
                                                         query_coverage=0.95, subject_coverage=0)
    assert os.path.isdir(blast_dir)
    G = O.Genome(contigs)
    for qi, qn in enumerate(lib_names):
        hsps = G.query_hsps(lib[qi])
        if not hsps:
            assert qn not in all_copies
            continue
        rows = [(h["contig"], h["qs"], h["qe"], h["ss"], h["se"]) for h in hsps]
        ident = [round(100.0 * h["n_match"] / h["aln_len"], 3) for h in hsps]
        # the oracle numbers subjects by contig index; the mirror by first appearance -- compare by name
        exp = [(names[c[0]], c[1], c[2], c[3], c[4]) for c in O.query_copies(rows, ident, len(lib[qi]), [1] * len(names))]
        assert sorted(all_copies[qn]) == sorted(exp), qn
    G.close()
    want = found = 0
    for t in truth:
        cps = all_copies.get(lib_names[t["query"]], [])
        hit = [c for c in cps if c[0] == names[t["contig"]] and abs(c[1] - t["start1"]) <= 12 and abs(c[2] - t["end1"]) <= 12
               and c[4] == t["strand"]]
        if t["kind"] == "trunc5":
            assert not hit or (t["end1"] - t["start1"] + 1) >= 0.95 * t["qlen"] - 12
        else:
            want += 1
            found += 1 if hit else 0
    assert want > 60 and found >= 0.95 * want, (found, want)
    # what run_Other_detection does with the result (judge_Other_transposons.py:50-77): the longest copy per query
    for qn, cps in all_copies.items():
        for c in cps:
            assert 1 <= c[1] <= c[2] <= len(contigs[names.index(c[0])])


def test_long_query_is_tiled(gctx, wl):
    """A library entry longer than the 32 000 bp kernel limit (two exist in the shipped non_LTR.lib) is tiled; its HSPs chain
    back into one copy."""
    from hite_b200 import boundary, other
    rng = np.random.default_rng(77)
    big = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 40_000)].tobytes()
    import hite_b200
    c2 = hite_b200.Context(0)
    contig = wl[1][0][:50_000] + big + wl[1][0][50_000:100_000]
    c2.load_genome([contig])
    c2.build_index(15, 10)
    c2.contig_names = ["chrX"]
    rec = other.query_hsps(c2, ["big#LINE/L1"], {"big#LINE/L1": big.decode()})
    hs = rec["big#LINE/L1"]["chrX"]
    assert len(hs) == 2 and hs[0][:4] == (1, 32000, 50_001, 82_000) and hs[1][:4] == (32001, 40000, 82_001, 90_000)
    cps = boundary.get_query_copies(list(rec.items()), {"big#LINE/L1": big.decode()}, "", 0.95, 0)
    assert len(cps["big#LINE/L1"]) == 1 and cps["big#LINE/L1"][0][:3] == ("chrX", 50_001, 90_000) and cps["big#LINE/L1"][0][4] == "+"
    c2.close()
