"""Known-TE masking end to end on the CPU: the mirror hite_b200.masking.mask_genome_intactTE (file in -> genome.masked out) with
an oracle-backed stand-in for the GPU context, against the REFERENCE's own mask_genome_intactTE (module/Util.py:6389, imported
unmodified) whose single external call (minimap2, :6318) is replaced by a writer of the SAME alignment records as PAF lines.
Given the same records the two masked genomes must be byte-identical.  Needs the reference checkout (skipped on the GPU box)."""
import os

import numpy as np
import pytest

import other_synth
from oracle import oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")


def test_mirror_masked_genome_equals_reference_on_the_same_records(tmp_path, monkeypatch):
    from hite_b200 import _lib, boundary, fasta, masking
    U = ref_import.load_util()
    names, contigs, lib_names, lib, truth = other_synth.make(seed=44, n_lib=25, n_contig=3, contig_bp=200_000, sub=(0.0, 0.06))
    contigs = [bytearray(c) for c in contigs]
    contigs[1][50] = ord("Y")                                    # IUPAC letters survive masking in both (no copy is planted below 200)
    contigs = [bytes(c) for c in contigs]
    lib_names = [n.replace("LINE/L1", "DNA/TIR") for n in lib_names]
    gpath, lpath = str(tmp_path / "genome.fa"), str(tmp_path / "prev_TE.fa")
    fasta.store_fasta({n: c.decode() for n, c in zip(names, contigs)}, gpath)
    fasta.store_fasta({n: s.decode() for n, s in zip(lib_names, lib)}, lpath)
    G = O.Genome(contigs)
    per_query = [G.query_hsps(q) for q in lib]
    G.close()

    class OracleBackedCtx:
        device = 0
        contig_names = names
        masked = None

        def query_hsps(self, seqs, occ_cap=512, max_loci=128):
            assert [bytes(s) for s in seqs] == lib
            rows = [(qi, h["contig"], h["qs"], h["qe"], h["ss"], h["se"], h["n_match"], h["aln_len"], h["score"], h["rank"])
                    for qi, hs in enumerate(per_query) for h in hs]
            return np.array(rows, dtype=_lib.QHSP_DTYPE)

        def timing(self):
            return dict(kernel_ms=0.0, k_ms=[0.0] * 8)

        def mask_genome(self, contig, start, end):               # the device copy: recorded, checked below
            self.masked = sorted(zip(contig, start, end))
            return 0

        def build_index(self, k, w):
            return dict(kernel_ms=0.0)

    ctx = OracleBackedCtx()
    monkeypatch.setattr(boundary, "get_context", lambda device=None, role="reference": ctx)
    monkeypatch.setattr(boundary, "ensure_genome", lambda c, ref, k=15, w=10, _parsed=None: c)
    monkeypatch.setattr(boundary, "_GENOME", {})
    out = masking.mask_genome_with_TE(lpath, gpath, str(tmp_path), 4, 0, debug=0)
    mine = open(out).read()
    os.rename(out, out + ".mine")

    def fake_minimap2(query_fasta, target_fasta, output_paf, threads, multi_records=100):
        with open(output_paf, "w") as f:
            for qi, hs in enumerate(per_query):
                for h in hs:
                    lo, hi = min(h["ss"], h["se"]), max(h["ss"], h["se"])
                    f.write("\t".join(str(x) for x in (lib_names[qi], len(lib[qi]), h["qs"] - 1, h["qe"], "+" if h["ss"] <= h["se"] else "-",
                                                       names[h["contig"]], len(contigs[h["contig"]]), lo - 1, hi, h["n_match"],
                                                       h["aln_len"], 60)) + "\n")
    monkeypatch.setattr(U, "te_genome_align_minimap2", fake_minimap2)
    U.mask_genome_intactTE(lpath, gpath, str(tmp_path), 1, 0, debug=0)
    ref = open(gpath + ".masked").read()
    assert mine == ref
    assert mine.count("N") > 50_000 and "Y" in mine
    # the intervals handed to the device are the ones the text was filled with
    n_text = sum(b - a for _, a, b in {(c, a, b) for c, a, b in ctx.masked})
    assert n_text >= mine.count("N")
