#!/usr/bin/env python
"""Golden fixtures for the TSD-bounded variant generation: the UNMODIFIED
reference functions search_confident_tir_v4 (module/Util.py:7734) and get_short_tir_contigs (:7297), imported
read-only through oracle/ref_import.py, run on seeded synthetic flanked candidates.

Run in the build container only:   python tests/golden/make_golden_tsd.py
Output (committed):  tests/golden/tsd_v4.json.gz -- per case the inputs and, per returned record, its index i in the
reference's distance-sorted list, the TSD k-mer, the distance, and (length, crc32) of the emitted sequence, plus
whether get_short_tir_contigs keeps it.
"""
import gzip
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from tsd_synth import synth_tsd_case  # noqa: E402


def main():
    U = ref_import.load_util()
    out = []
    for seed in range(1, 241):
        c = synth_tsd_case(seed)
        res = U.search_confident_tir_v4(c["seq"], c["raw_start"], c["raw_end"], c["d"], "q", c["plant"])
        short = U.get_short_tir_contigs(res, c["plant"])
        recs = []
        for name, s in res.items():
            # q-C_<i>-tsd_<kmer>-distance_<dist>
            i = int(name.split("-C_")[1].split("-")[0])
            kmer = name.split("-tsd_")[1].split("-")[0]
            dist = int(name.split("-distance_")[1])
            recs.append([i, kmer, dist, len(s), zlib.crc32(s.encode()) & 0xFFFFFFFF, 1 if name in short else 0])
        c["expect"] = recs
        out.append(c)
    path = os.path.join(ROOT, "tests", "golden", "tsd_v4.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    from collections import Counter
    print("wrote", path, len(out), "cases;", sum(len(c["expect"]) for c in out), "records;",
          sum(r[5] for c in out for r in c["expect"]), "short-TIR;", os.path.getsize(path), "bytes")
    print(Counter((c["scenario"], min(3, len(c["expect"]))) for c in out))


if __name__ == "__main__":
    main()
