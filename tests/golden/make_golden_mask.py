#!/usr/bin/env python
"""Golden fixtures for the known-TE genome masking, made with the UNMODIFIED reference
functions imported read-only through oracle/ref_import.py.  Run in the build container only:

    python tests/golden/make_golden_mask.py

1. tests/golden/full_length_copies.json.gz -- get_full_length_copies_from_blastn_v1 (module/Util.py:5907) on seeded HSP
   lists of one query against several subjects (the lists of tests/msa_synth.synth_copy_hsps; half of them with the
   subject span made ascending, which is what convert_paf_to_blast6_format :6335-6374 hands it inside
   mask_genome_intactTE): the 'chr:start-end' keys it returns.
2. tests/golden/mask_genome.json.gz -- mask_genome_intactTE itself (:6389-6431) on small genomes, its one external call
   (te_genome_align_minimap2, :6318) replaced by a writer of a fixed PAF: the masked contigs (crc32) and the N intervals.
"""
import gzip
import json
import os
import random
import sys
import tempfile
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_import  # noqa: E402
import msa_synth  # noqa: E402
import mask_synth  # noqa: E402


def run_full_length(U, case, tmpdir):
    lib = os.path.join(tmpdir, "lib.fa")
    ref = os.path.join(tmpdir, "ref.fa")
    out = os.path.join(tmpdir, "blast.out")
    with open(lib, "w") as f:
        f.write(">q#DNA/TIR\n%s\n" % ("A" * case["qlen"]))
    with open(ref, "w") as f:
        for i, ln in enumerate(case["subj_lens"]):
            f.write(">s%d\n%s\n" % (i, "A" * ln))
    with open(out, "w") as f:
        f.write("#qseqid\tsseqid\tpident\tlength\tmismatch\tgapopen\tqstart\tqend\tsstart\tsend\tevalue\tbitscore\n")
        for (si, qs, qe, ss, se) in case["hsps"]:
            f.write("q#DNA/TIR\ts%d\t90.0\t%d\t0\t0\t%d\t%d\t%d\t%d\t0.0\t0.0\n" % (si, abs(qe - qs) + 1, qs, qe, ss, se))
    full, _ = U.get_full_length_copies_from_blastn_v1(lib, ref, out, tmpdir, 1, 20, case["threshold"], False, "")
    res = []
    for key in full.get("q", {}).keys():
        name, span = key.split(":")
        a, b = span.split("-")
        res.append([int(name[1:]), int(a), int(b)])
    return res


def run_mask(U, case, tmpdir):
    lib = os.path.join(tmpdir, "prev_TE.fa")
    gen = os.path.join(tmpdir, "genome.fa")
    with open(lib, "w") as f:
        for n, s in zip(case["lib_names"], case["lib"]):
            f.write(">%s\n%s\n" % (n, s))
    with open(gen, "w") as f:
        for n, s in zip(case["names"], case["contigs"]):
            f.write(">%s\n%s\n" % (n, s))

    def fake_minimap2(query_fasta, target_fasta, output_paf, threads, multi_records=100):
        with open(output_paf, "w") as f:
            for p in case["paf"]:
                f.write("\t".join(str(x) for x in p) + "\n")
    saved = U.te_genome_align_minimap2
    U.te_genome_align_minimap2 = fake_minimap2
    try:
        U.mask_genome_intactTE(lib, gen, tmpdir, 1, 0, debug=0)
    finally:
        U.te_genome_align_minimap2 = saved
    names, contigs = U.read_fasta(gen + ".masked")
    return [[n, zlib.crc32(contigs[n].encode()) & 0xFFFFFFFF, contigs[n].count("N")] for n in names]


def write(path, out):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, len(out), "cases,", os.path.getsize(path), "bytes")


def main():
    U = ref_import.load_util()
    gold = os.path.join(ROOT, "tests", "golden")
    out = []
    for seed in range(1, 241):
        case = msa_synth.synth_copy_hsps(20_000 + seed)
        rng = random.Random(seed)
        case = dict(qlen=case["qlen"], subj_lens=case["subj_lens"], hsps=[list(h) for h in case["hsps"]], seed=seed,
                    threshold=rng.choice([0.95, 0.95, 0.8]), paf_like=seed % 2)
        if case["paf_like"]:
            case["hsps"] = [[si, qs, qe, min(ss, se), max(ss, se)] for (si, qs, qe, ss, se) in case["hsps"]]
        with tempfile.TemporaryDirectory() as td:
            case["expect"] = run_full_length(U, case, td)
        out.append(case)
    print(sum(len(c["expect"]) for c in out), "copies;", sum(1 for c in out if not c["expect"]), "cases without a copy")
    write(os.path.join(gold, "full_length_copies.json.gz"), out)
    out = []
    for seed in range(1, 13):
        case = mask_synth.mask_case(seed)
        with tempfile.TemporaryDirectory() as td:
            case["expect"] = run_mask(U, case, td)
        # the fixture keeps what is needed to rebuild the inputs from the seed, not the sequences themselves
        out.append(dict(seed=seed, paf=case["paf"], expect=case["expect"]))
    print(sum(e[2] for c in out for e in c["expect"]), "masked bases")
    write(os.path.join(gold, "mask_genome.json.gz"), out)


if __name__ == "__main__":
    main()
