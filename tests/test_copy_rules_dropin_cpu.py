"""The copy rules of the reference's get_copies_minimap2 (module/Util.py:7977-8030, imported unmodified) on the oracle's own
alignments: every candidate x locus alignment of the oracle is written as a SAM record (aggregated CIGAR: the rule only sums
operations), read back through a minimal stand-in for pysam.AlignmentFile, and the reference's selection -- both coverages
>= 0.95, 1-based inclusive coordinates, strand -- must pick exactly the alignments the oracle (and with it copy_filter_kernel)
marks as passing the coverage rules.  Needs the reference checkout (skipped on the GPU box)."""
import re
import types

import pytest

from workloads import synth
from oracle import oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")


class _Read:
    def __init__(self, line):
        f = line.rstrip("\n").split("\t")
        self.query_name, flag, self.reference_name, pos, cigar = f[0], int(f[1]), f[2], int(f[3]), f[5]
        self.is_unmapped = bool(flag & 4)
        self.is_reverse = bool(flag & 16)
        self.cigartuples = [("MIDNSHP=X".index(op), int(n)) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]
        self.query_alignment_length = sum(n for op, n in self.cigartuples if op in (0, 1, 7, 8))
        self.reference_start = pos - 1
        self.reference_end = self.reference_start + sum(n for op, n in self.cigartuples if op in (0, 2, 3, 7, 8))


class _AlignmentFile:
    def __init__(self, path, mode="r", check_sq=True):
        self._lines = [l for l in open(path) if not l.startswith("@")]

    def __iter__(self):
        return (_Read(l) for l in self._lines)

    def close(self):
        pass


def test_reference_copy_rules_select_the_oracles_copies(tmp_path, monkeypatch):
    U = ref_import.load_util()
    wl = synth.make_workload(genome_bp=1_500_000, n_contigs=2, n_families=12, cands_per_family=3, copies=(4, 25),
                             length=(150, 1500), seed=41)
    ref, cand = str(tmp_path / "genome.fa"), str(tmp_path / "cand.fa")
    wl.write_fasta(ref, cand)
    G = O.Genome(wl.contig_seqs)
    tot = G.total
    goff = [G._L.hbo_genome_contig_off(G.g, c) for c in range(len(wl.contig_seqs))] if hasattr(G._L, "hbo_genome_contig_off") else None
    sam = str(tmp_path / "query.sam")
    expect = {}
    n_partial = 0
    with open(sam, "w") as f:
        f.write("@HD\tVN:1.6\n")
        for name, seq in zip(wl.cand_names, wl.cand_seqs):
            cps, _ = G.candidate_msa(seq)
            for c in cps:
                if c["score"] <= 0:
                    continue
                q_aln = c["i_e"] - c["i_s"] + 1                      # M + I
                t_aln = c["j_e"] - c["j_s"] + 1                      # M + D
                m = min(q_aln, t_aln)
                cigar = "%dM" % m + ("%dI" % (q_aln - m) if q_aln > m else "") + ("%dD" % (t_aln - m) if t_aln > m else "")
                if c["contig"] >= 0:                                  # passed both coverage rules inside one contig
                    start1, end1, contig = c["start1"], c["end1"], c["contig"]
                    expect.setdefault(name, []).append((wl.contig_names[contig], start1, end1, q_aln, "-" if c["rel"] else "+"))
                else:
                    # a failing alignment: written at a fixed place, the reference must drop it on the numbers alone
                    contig, start1 = 0, 1000
                    if q_aln / len(seq) >= 0.95 and q_aln / t_aln >= 0.95:
                        continue                                      # passes the numbers but lies across two contigs: no mapper reports that
                    n_partial += 1
                f.write("\t".join([name, "16" if c["rel"] else "0", wl.contig_names[contig], str(start1), "60", cigar, "*", "0", "0",
                                   "*", "*"]) + "\n")
        f.write("unmapped_read\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n")
    G.close()
    fake = types.ModuleType("pysam")
    fake.AlignmentFile = _AlignmentFile
    monkeypatch.setattr(U, "pysam", fake)
    got = U.get_copies_minimap2(cand, sam)
    for name in set(got) | set(expect):
        assert got.get(name, []) == expect.get(name, []), name
    assert sum(len(v) for v in expect.values()) > 300 and n_partial > 10
