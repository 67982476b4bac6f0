"""Seeded inputs for the known-TE genome masking tests (reference mask_genome_intactTE, module/Util.py:6389): a small
genome, a TE library (prev_TE) and the PAF records a mapper could report for it -- full-length hits on both strands,
hits split into two records, partial hits, a record touching a contig end, IUPAC letters in the genome."""
import numpy as np

import other_synth


def mask_case(seed):
    rng = np.random.default_rng(1000 + seed)
    names, contigs, lib_names, lib, truth = other_synth.make(seed=300 + seed, n_lib=8, n_contig=2, contig_bp=60_000,
                                                             copies=(1, 5), length=(300, 2500), sub=(0.0, 0.05))
    contigs = [bytearray(c) for c in contigs]
    for c in contigs:                                      # a few IUPAC / lower-case-free oddities the mask must keep
        for p in rng.integers(0, len(c), 5):
            c[int(p)] = ord("R")
    lib_names = [n.replace("LINE/L1", "DNA/TIR") for n in lib_names]
    paf = []
    for t in truth:
        qn, qlen = lib_names[t["query"]], t["qlen"]
        tlen = len(contigs[t["contig"]])
        ts, te = t["start1"] - 1, t["end1"]                # PAF: 0-based start, end exclusive
        qs, qe = t["qstart1"] - 1, qlen
        n_match = int((te - ts) * 0.97)
        rec = lambda a, b, c, d: [qn, qlen, a, b, t["strand"], names[t["contig"]], tlen, c, d, int((d - c) * 0.97), d - c, 60]
        if t["kind"] == "deletion":                        # two records 150 bp apart on the query
            cut = (te - ts) // 2
            if t["strand"] == "+":
                paf.append(rec(0, cut, ts, ts + cut)); paf.append(rec(cut + 150, qlen, ts + cut, te))
            else:
                paf.append(rec(cut + 150, qlen, ts, te - cut)); paf.append(rec(0, cut, te - cut, te))
        else:
            paf.append(rec(qs, qe, ts, te))
        if rng.random() < 0.3:                             # a secondary, partial record of the same query elsewhere
            p = int(rng.integers(100, tlen - 700))
            paf.append(rec(10, 10 + 500, p, p + 500))
    paf.append([lib_names[0], len(lib[0]), 0, len(lib[0]), "+", names[0], len(contigs[0]), len(contigs[0]) - len(lib[0]), len(contigs[0]),
                len(lib[0]), len(lib[0]), 60])              # ends exactly at the contig end
    paf.append(["not_in_library", 500, 0, 500, "+", names[0], len(contigs[0]), 1000, 1500, 500, 500, 60])   # :5944 skip
    order = rng.permutation(len(paf))
    paf = [paf[int(i)] for i in order]
    return dict(seed=seed, names=names, contigs=[bytes(c).decode() for c in contigs], lib_names=lib_names,
                lib=[x.decode() for x in lib], paf=paf, truth=truth)
