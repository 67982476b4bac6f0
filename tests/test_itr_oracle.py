"""HB-itr (the stand-in for the closed binary tools/itrsearch) against what the BINARY itself returned on 1 500 seeded records
(tests/golden/itrsearch.json.gz, made by tests/golden/make_golden_itr.py in the build container).  The binary's search is a
heuristic whose source is not available, so this is a measured agreement, not bit-parity ("parity unpinned"): the presence
decision must agree on >= 96 % of the records and, whenever both report a terminal inverted repeat, the two arm lengths --
the `ITR(1,a)..(L,b)` and `Length itr=` the reference parses (module/Util.py:6586-6593) -- must be identical on >= 97 %."""
import gzip
import json
import os

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "itrsearch.json.gz"), "rt") as f:
        return json.load(f)["cases"]


def test_hb_itr_agrees_with_the_itrsearch_binary():
    cases = _cases()
    agree = both = same = 0
    for c in cases:
        o = O.itr_search(c["seq"])
        agree += bool(o["found"]) == bool(c["saved"])
        if o["found"] and c["saved"]:
            both += 1
            same += (o["len5"], o["len3"]) == (c["len5"], c["len3"])
            if (o["len5"], o["len3"]) == (c["len5"], c["len3"]):
                L = len(c["seq"])                                 # the header the reference parses
                assert c["header"] == "%s ITR(1,%d)..(%d,%d) Length itr=%d" % (c["name"], o["len5"], L, L - o["len3"] + 1, o["len5"] - 1)
    print("HB-itr vs itrsearch binary: presence agrees on %d / %d records; both report a TIR on %d, same arms on %d"
          % (agree, len(cases), both, same))
    assert agree >= 0.96 * len(cases) and both > 400 and same >= 0.97 * both


def test_hb_itr_basics():
    comp = str.maketrans("ACGT", "TGCA")
    tir = "GTAGCCAGAAGGCTGCAACTCATCGACTCT"
    s = tir + "ATGTAGTGACCGCGTCGATGTCAAACCCCGGGGGGAGCTCAGATATCCGATACAGGGATG" + tir.translate(comp)[::-1]
    o = O.itr_search(s)
    assert o["found"] == 1 and (o["len5"], o["len3"]) == (30, 30) and o["score"] == 300
    assert O.itr_search(s.lower())["found"] == 1                                  # case-insensitive
    assert O.itr_search("ACGT")["found"] == 0 and O.itr_search("")["found"] == 0   # too short for the minimum arm
    o = O.itr_search("N" * 40 + "ACGT" * 5 + "N" * 40)
    assert o["found"] == 0                                                         # N never matches
    long_tir = (tir * 30)[:700]
    o = O.itr_search(long_tir + "ACGTTGCA" * 100 + long_tir.translate(comp)[::-1])
    assert o["found"] == 1 and o["len5"] == 500                                   # the window is 500 bases per end
