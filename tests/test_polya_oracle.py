"""Oracle (oracle/hbo_tsd.c::hbo_search_polyA_TSD) vs the reference's own search_polyA_TSD / get_candidate_non_LTR outputs
(tests/golden/polya_tsd.json.gz, made by tests/golden/make_golden_polya.py from the imported reference)."""
import gzip
import json
import os
import zlib

import pytest

from oracle import oracle as O
from polya_synth import polya_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_polya():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "polya_tsd.json.gz")) as f:
        return json.loads(f.read())


def test_oracle_matches_reference_goldens(golden_polya):
    n_found = 0
    for g in golden_polya["cases"]:
        c = polya_case(g["seed"])
        found, tsd, body, rec = O.search_polyA_TSD(c["seq"], c["flank"])
        assert [found, tsd, len(body), zlib.crc32(body.encode()) & 0xFFFFFFFF] == g["expect"], (g["seed"], c["scenario"], rec)
        n_found += found
    assert n_found > 100


def candidates_from(search, cases, flank=50):
    """get_candidate_non_LTR (module/Util.py:11008-11045) on top of a search function -> (sine, line) lists"""
    sine, line = [], []
    for c in cases:
        seq = c["seq"]
        name = "chr1:%d-%d" % (c["seed"] * 10000, c["seed"] * 10000 + len(seq))
        seq_len = len(seq) - 2 * flank
        if not (100 <= seq_len <= 8000):
            continue
        found, tsd, body = search(seq, flank)[:3]
        rec = [name + "\tTSD:" + tsd, len(body), zlib.crc32(body.encode()) & 0xFFFFFFFF]
        if seq_len <= 700:
            if found and 100 <= len(body) <= 700:
                sine.append(rec)
        elif found and 700 < len(body) <= 8000:
            line.append(rec)
    return sine, line


def test_oracle_candidate_classes_match_reference(golden_polya):
    cases = [polya_case(g["seed"]) for g in golden_polya["cases"]]
    sine, line = candidates_from(O.search_polyA_TSD, [c for c in cases if c["flank"] == 50])
    assert sine == golden_polya["sine"] and line == golden_polya["line"]
    assert len(sine) > 10 and len(line) > 30
