"""Oracle (oracle/hbo_fmea.c::hbo_full_length_copies_v1) vs the reference's get_full_length_copies_from_blastn_v1 outputs
(tests/golden/full_length_copies.json.gz), and the oracle-side mask_genome_intactTE restatement vs the reference's masked
genomes (tests/golden/mask_genome.json.gz).  Goldens made by tests/golden/make_golden_mask.py from the imported reference."""
import gzip
import json
import os
import zlib

import pytest

import mask_synth
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    with gzip.open(os.path.join(ROOT, "tests", "golden", name)) as f:
        return json.loads(f.read())


def paf_to_hsps(paf, lib_names, names):
    """convert_paf_to_blast6_format (Util.py:6335-6374) + the parse of get_full_length_copies_from_blastn_v1 (:5918-5938):
    {query(without '#...'): [(subj, qs, qe, ss, se)]} -- strand is dropped, subject spans ascend."""
    known = {n.split("#")[0] for n in lib_names}
    out = {}
    for p in paf:
        q = p[0].split("#")[0]
        if q not in known:
            continue
        out.setdefault(q, []).append((names.index(p[5]), p[2] + 1, p[3], p[7] + 1, p[8]))
    return out


def oracle_mask(case):
    """mask_genome_intactTE (Util.py:6389-6431) on the oracle: chain per query, N-fill [start, end) per copy."""
    contigs = [bytearray(c.encode()) for c in case["contigs"]]
    qlen = {n.split("#")[0]: len(s) for n, s in zip(case["lib_names"], case["lib"])}
    for q, hs in paf_to_hsps(case["paf"], case["lib_names"], case["names"]).items():
        for subj, start, end, _ in O.full_length_copies_v1(hs, qlen[q], 0.95):
            c = contigs[subj]
            start, end = max(0, start), min(len(c), end)
            if end > start:
                c[start:end] = b"N" * (end - start)
    return contigs


def test_full_length_copies_oracle_matches_reference_goldens():
    n = 0
    for case in load("full_length_copies.json.gz"):
        got = O.full_length_copies_v1([tuple(h) for h in case["hsps"]], case["qlen"], case["threshold"])
        assert [[g[0], g[1], g[2]] for g in got] == case["expect"], case["seed"]
        n += len(got)
    assert n > 1000


def test_mask_genome_oracle_matches_reference_goldens():
    masked = 0
    for g in load("mask_genome.json.gz"):
        case = mask_synth.mask_case(g["seed"])
        assert case["paf"] == g["paf"]                      # the seeded inputs are the ones the golden was made from
        contigs = oracle_mask(case)
        for (name, crc, n_count), c in zip(g["expect"], contigs):
            assert (zlib.crc32(bytes(c)) & 0xFFFFFFFF, bytes(c).count(b"N")) == (crc, n_count), (g["seed"], name)
            masked += n_count
    assert masked > 100_000


# ---- product host code (no GPU needed: hb_full_length_copies is host C++ inside libhite_b200.so) -------------------------
def test_product_full_length_copies_matches_reference_goldens():
    import numpy as np
    from hite_b200 import _lib
    n = 0
    for case in load("full_length_copies.json.gz"):
        hs = np.array([tuple(h) for h in case["hsps"]], dtype=_lib.HSP_DTYPE)
        got = _lib.full_length_copies(hs, case["qlen"], case["threshold"])
        assert [[int(g["subj"]), int(g["start"]), int(g["end"])] for g in got] == case["expect"], case["seed"]
        n += len(got)
    assert n > 1000
    assert len(_lib.full_length_copies(np.zeros(0, dtype=_lib.HSP_DTYPE), 100, 0.95)) == 0


def test_product_mask_text_matches_reference_goldens():
    """hite_b200.masking.records_to_intervals + fill_text on the goldens' PAF records == the reference's masked genome
    (module/Util.py:6389-6431), IUPAC letters kept, '#' cut from query names, strand dropped, unknown queries skipped."""
    from hite_b200 import masking
    for g in load("mask_genome.json.gz"):
        case = mask_synth.mask_case(g["seed"])
        records = {}
        for p in case["paf"]:                                   # convert_paf_to_blast6_format (:6349-6356)
            records.setdefault(p[0], {}).setdefault(p[5], []).append((p[2] + 1, p[3], p[7] + 1, p[8], 97.0))
        qlens = {n: len(s) for n, s in zip(case["lib_names"], case["lib"])}
        iv = masking.records_to_intervals(records, qlens, case["names"])
        masked = masking.fill_text(dict(zip(case["names"], case["contigs"])), case["names"], iv)
        for name, crc, n_count in g["expect"]:
            s = masked[name]
            assert (zlib.crc32(s.encode()) & 0xFFFFFFFF, s.count("N")) == (crc, n_count), (g["seed"], name)


def test_product_batched_chaining_matches_reference_goldens():
    """hb_full_length_copies_batch / hb_query_copies_batch (threads over queries) == the per-query goldens."""
    import numpy as np
    from hite_b200 import _lib
    cases = load("full_length_copies.json.gz")
    for thr in (0.95, 0.8):
        sub = [c for c in cases if c["threshold"] == thr]
        hs = np.array([tuple(h) for c in sub for h in c["hsps"]], dtype=_lib.HSP_DTYPE)
        q_off = np.cumsum([0] + [len(c["hsps"]) for c in sub])
        out, oq = _lib.full_length_copies_batch(hs, q_off, [c["qlen"] for c in sub], thr, threads=4)
        for g, c in enumerate(sub):
            got = [[int(o["subj"]), int(o["start"]), int(o["end"])] for o in out[oq == g]]
            assert got == c["expect"], c["seed"]
    cases = [c for c in load("query_copies.json.gz") if c["subject_coverage"] == 0]
    assert len(cases) > 50
    for cov in (0.95, 0.8):
        for mc in (100, 3):
            sub = [c for c in cases if c["query_coverage"] == cov and c["max_copy_num"] == mc]
            if not sub:
                continue
            hs = np.array([tuple(h) for c in sub for h in c["hsps"]], dtype=_lib.HSP_DTYPE)
            idn = [i for c in sub for i in c["idents"]]
            q_off = np.cumsum([0] + [len(c["hsps"]) for c in sub])
            out, oq = _lib.query_copies_batch(hs, idn, q_off, [c["qlen"] for c in sub], query_coverage=cov, max_copy_num=mc, threads=3)
            for g, c in enumerate(sub):
                got = [[int(o["subj"]), int(o["start"]), int(o["end"]), int(o["qlen"]), "-" if o["strand"] else "+"] for o in out[oq == g]]
                assert got == c["expect"], c["seed"]


def test_product_full_length_copies_fuzz_vs_oracle():
    """Adversarial HSP lists (tiny coordinate ranges, duplicates, reversed query spans, single-base subjects, mixed strands):
    product host code == oracle.  The same generator was run against the imported reference in the build container
    (1 500 cases, thresholds 0.95 / 0.8 / 0.5 / 0.3: reference == oracle == product)."""
    import random
    import numpy as np
    import msa_synth
    from hite_b200 import _lib
    for seed in range(600):
        rng = random.Random(seed)
        if seed % 2 == 0:
            c = msa_synth.synth_copy_hsps(70000 + seed)
            hs, qlen = [tuple(h) for h in c["hsps"]], c["qlen"]
        else:
            qlen, nsubj, hs = rng.randint(20, 200), rng.randint(1, 3), []
            for _ in range(rng.randint(1, 25)):
                qs, qe = rng.randint(1, qlen), rng.randint(1, qlen)
                if rng.random() < 0.8 and qs > qe:
                    qs, qe = qe, qs
                hs.append((rng.randrange(nsubj), qs, qe, rng.randint(1, 400), rng.randint(1, 400)))
                if rng.random() < 0.2:
                    hs.append(hs[-1])
        thr = rng.choice([0.95, 0.8, 0.5, 0.3])
        exp = [(g[0], g[1], g[2]) for g in O.full_length_copies_v1(hs, qlen, thr)]
        got = _lib.full_length_copies(np.array(hs, dtype=_lib.HSP_DTYPE), qlen, thr)
        assert [(int(g["subj"]), int(g["start"]), int(g["end"])) for g in got] == exp, seed
