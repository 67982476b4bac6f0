"""Oracle for the library-vs-genome HSPs (oracle/hbo_align.c::hbo_query_hsps) against planted truth, and the chain
HSPs -> get_query_copies (oracle/hbo_fmea.c, pinned on reference goldens) -> copies.  CPU only."""
import other_synth
from oracle import oracle as O


def test_oracle_hsps_recover_planted_copies():
    names, contigs, lib_names, lib, truth = other_synth.make(seed=5, n_lib=12, n_contig=2, contig_bp=150_000, sub=(0.0, 0.0))
    G = O.Genome(contigs)
    by_q = {}
    for t in truth:
        by_q.setdefault(t["query"], []).append(t)
    n_full = 0
    for qi, ts in by_q.items():
        hsps = G.query_hsps(lib[qi])
        assert [h["rank"] for h in hsps] == sorted(h["rank"] for h in hsps)
        for h in hsps:
            assert 1 <= h["qs"] <= h["qe"] <= len(lib[qi]) and h["n_match"] <= h["aln_len"] and h["score"] > 0
        for t in ts:
            if t["kind"] != "full":
                continue
            # an undiverged full-length copy is one HSP over the whole query, all matches, exact subject span
            lo, hi = (t["start1"], t["end1"]) if t["strand"] == "+" else (t["end1"], t["start1"])
            hit = [h for h in hsps if h["contig"] == t["contig"] and (h["ss"], h["se"]) == (lo, hi)]
            assert len(hit) == 1, (qi, t, hsps)
            h = hit[0]
            assert (h["qs"], h["qe"]) == (1, t["qlen"]) and h["n_match"] == h["aln_len"] == t["qlen"] and h["score"] == 2 * t["qlen"]
            n_full += 1
    assert n_full >= 10
    G.close()


def test_oracle_hsps_chain_into_copies():
    names, contigs, lib_names, lib, truth = other_synth.make(seed=9, n_lib=10, n_contig=2, contig_bp=150_000, sub=(0.0, 0.05))
    G = O.Genome(contigs)
    found = want = 0
    for qi in range(len(lib)):
        hsps = G.query_hsps(lib[qi])
        if not hsps:
            continue
        rows = [(h["contig"], h["qs"], h["qe"], h["ss"], h["se"]) for h in hsps]
        ident = [round(100.0 * h["n_match"] / h["aln_len"], 3) for h in hsps]
        copies = O.query_copies(rows, ident, len(lib[qi]), [1] * len(contigs))
        for t in truth:
            if t["query"] != qi or t["kind"] == "trunc5":
                continue
            want += 1
            # full-length and internally deleted copies (two HSPs, 150 bp apart) come back as one copy
            ok = [c for c in copies if c[0] == t["contig"] and abs(c[1] - t["start1"]) <= 12 and abs(c[2] - t["end1"]) <= 12
                  and c[4] == t["strand"]]
            found += 1 if ok else 0
    assert want >= 15 and found >= 0.9 * want, (found, want)
    G.close()
