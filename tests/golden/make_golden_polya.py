#!/usr/bin/env python
"""Golden fixtures for the non-LTR candidate generator: the UNMODIFIED reference functions
search_polyA_TSD (module/Util.py:10915) and get_candidate_non_LTR (:11008), imported read-only through oracle/ref_import.py
(fuzzysearch stubbed by oracle/fnm.py -- only the existence of a match is consumed), on seeded synthetic flanked candidates.

Run in the build container only:   python tests/golden/make_golden_polya.py
Output (committed): tests/golden/polya_tsd.json.gz -- per case (found_TSD, TSD_seq, len and crc32 of non_ltr_seq), and for
the whole set the names get_candidate_non_LTR keeps as SINE / LINE candidates.
"""
import gzip
import json
import os
import sys
import tempfile
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_import  # noqa: E402
from polya_synth import polya_case  # noqa: E402


def main():
    U = ref_import.load_util()
    cases = []
    for seed in range(1, 401):
        c = polya_case(seed)
        found, tsd, body = U.search_polyA_TSD(c["seq"], c["flank"], 25, list(range(8, 21)))
        c["expect"] = [bool(found), tsd, len(body), zlib.crc32(body.encode()) & 0xFFFFFFFF]
        cases.append(c)
    # get_candidate_non_LTR on the flank-50 cases as one FASTA
    sub = [c for c in cases if c["flank"] == 50]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "longest_repeats.flanked.fa")
        with open(path, "w") as f:
            for c in sub:
                f.write(">chr1:%d-%d\n%s\n" % (c["seed"] * 10000, c["seed"] * 10000 + len(c["seq"]), c["seq"]))
        sine, line = U.get_candidate_non_LTR(path, 50)
    out = dict(cases=[dict(seed=c["seed"], expect=c["expect"]) for c in cases],
               sine=[[n, len(s), zlib.crc32(s.encode()) & 0xFFFFFFFF] for n, s in sine.items()],
               line=[[n, len(s), zlib.crc32(s.encode()) & 0xFFFFFFFF] for n, s in line.items()])
    path = os.path.join(ROOT, "tests", "golden", "polya_tsd.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    from collections import Counter
    print("wrote", path, os.path.getsize(path), "bytes;", sum(1 for c in cases if c["expect"][0]), "of", len(cases), "with a TSD;",
          len(sine), "SINE,", len(line), "LINE candidates")
    print(Counter((c["scenario"], c["expect"][0], c["expect"][2] > 0) for c in cases))


if __name__ == "__main__":
    main()
