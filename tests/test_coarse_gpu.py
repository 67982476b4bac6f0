"""GPU: coarse-boundary all-vs-all (self-join of the minimizer index -> HSPs) + FMEA through the C-ABI vs the
plain-C restatement (oracle/hbo_hsp.c + hbo_fmea.c); planted copies must be found."""
import numpy as np
import pytest

from workloads import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_coarse_boundary_bit_exact_and_finds_planted_copies(ctx):
    # every copy locus becomes a (exactly named) candidate record, so truth[] lists all planted loci
    wl = synth.make_workload(genome_bp=3_000_000, n_contigs=2, n_families=10, cands_per_family=20, copies=(6, 14),
                             length=(300, 1500), seed=31, sub=(0.0, 0.04), indel=0.003, inexact_frac=0.0)
    lens = [len(c) for c in wl.contig_seqs]
    G = oracle.Genome(wl.contig_seqs)
    ctx.load_genome(wl.contig_seqs)
    ctx.build_index(15, 10)
    for seg_len, occ_cap in ((1_000_000, 64), (400_000, 8)):
        exp_recs, exp_segs, exp_iv = G.coarse_boundary(lens, seg_len=seg_len, skip_gap=2000, max_len=30000, occ_cap=occ_cap)
        iv = ctx.coarse_boundary(seg_len=seg_len, skip_gap=2000, max_len=30000, occ_cap=occ_cap, threads=4)
        hsps, q_off, segs = ctx.last_hsps()
        assert [(int(s["chr"]), int(s["start"])) for s in segs] == exp_segs
        # GPU HSPs are grouped by query segment (stable); the oracle list is in sorted-key order
        got = []
        for g in range(len(segs)):
            for h in hsps[int(q_off[g]):int(q_off[g + 1])]:
                got.append((g, int(h["subj"]), int(h["qs"]), int(h["qe"]), int(h["ss"]), int(h["se"])))
        order = np.argsort(exp_recs[:, 0], kind="stable")
        exp = [tuple(int(x) for x in exp_recs[i]) for i in order]
        assert got == exp
        assert [(int(o["chr"]), int(o["start"]), int(o["end"])) for o in iv] == exp_iv
    # planted truth (coarse: within 30 bp at both ends) for the default segmentation
    iv = ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64)
    found = {(int(o["chr"]), int(o["start"]) // 30, int(o["end"]) // 30) for o in iv}
    starts = {}
    for o in iv:
        starts.setdefault(int(o["chr"]), []).append((int(o["start"]), int(o["end"])))
    fam_hit = {}
    for name, tr in zip(wl.cand_names, wl.truth):       # FMEA keeps every repeated segment once: per family
        c, rng = name.split(":")
        a, b = (int(x) for x in rng.split("-"))
        ci = wl.contig_names.index(c)
        ok = any(abs(s - a) <= 40 and abs(e - b) <= 40 for s, e in starts.get(ci, []))
        fam_hit[tr["family"]] = fam_hit.get(tr["family"], False) or ok
    assert sum(fam_hit.values()) >= 0.8 * len(fam_hit), fam_hit


def test_sharded_coarse_boundary_parts_concatenate_to_the_whole(ctx, monkeypatch):
    """hb_coarse_boundary_part: the query-segment files dealt in contiguous blocks to 1, 2, 3, 5 parts; every part pairs its
    minimizers against the whole index and chains its own files.  Parts concatenated in part order == the unsharded list
    (which is the oracle's), also when there are more parts than files."""
    wl = synth.make_workload(genome_bp=5_200_000, n_contigs=3, n_families=14, cands_per_family=1, copies=(6, 14),
                             length=(300, 1500), seed=33, sub=(0.0, 0.04), indel=0.003)
    ctx.load_genome(wl.contig_seqs, wl.contig_names)
    ctx.build_index(15, 10)
    whole = ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64, threads=4)
    G = oracle.Genome(wl.contig_seqs)
    _, _, exp_iv = G.coarse_boundary([len(c) for c in wl.contig_seqs], seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64)
    key = lambda iv: [(int(o["chr"]), int(o["start"]), int(o["end"])) for o in iv]
    assert key(whole) == exp_iv and len(exp_iv) > 50
    for n_parts in (2, 3, 5, 9):
        got, sizes = [], []
        for p in range(n_parts):
            part = ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64, threads=4, part=p, n_parts=n_parts)
            sizes.append(len(part))
            got += key(part)
        assert got == exp_iv, n_parts
        assert sum(1 for x in sizes if x) >= min(n_parts, 4)          # the work really is spread over the parts
    # the software pipeline inside one part (GPU block b + 1 while the host chains block b) never changes the result:
    # blocks are whole files.  Real chunks get 4-8 blocks; the hook forces them on this small genome, HSP getter included.
    ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64, threads=4)
    whole_hsps, whole_off, _ = (np.array(x) for x in ctx.last_hsps())
    for n_blocks in ("2", "3", "8"):
        monkeypatch.setenv("HB_COARSE_BLOCKS", n_blocks)
        blk = ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64, threads=4)
        assert key(blk) == exp_iv, n_blocks
        hs, off, _ = ctx.last_hsps()
        assert np.array_equal(np.array(off), whole_off) and np.array_equal(np.array(hs), whole_hsps), n_blocks
    monkeypatch.delenv("HB_COARSE_BLOCKS")


def test_determine_repeat_boundary_v5_fasta_contract(ctx, tmp_path):
    """cut-chunk FASTA in -> longest_repeats FASTA + .flanked.fa out, same names/sequence rules as the
    reference (Util.py:4354-4363, 4614-4634, 4783-4796)."""
    from hite_b200 import coarse, fasta
    wl = synth.make_workload(genome_bp=2_400_000, n_contigs=2, n_families=6, cands_per_family=1, copies=(6, 10),
                             length=(300, 900), seed=41)
    ref = str(tmp_path / "genome.fa")
    cand = str(tmp_path / "cand.fa")
    wl.write_fasta(ref, cand)
    # split_genome_chunks.py:44-52 layout: chr$start records of 1 Mbp
    cut = str(tmp_path / "genome.cut0.fa")
    with open(cut, "w") as f:
        for n, s in zip(wl.contig_names, wl.contig_seqs):
            t = s.tobytes().decode()
            for o in range(0, len(t), 1_000_000):
                f.write(">%s$%d\n%s\n" % (n, o, t[o:o + 1_000_000]))
    out = str(tmp_path / "longest_repeats_0.fa")
    coarse.determine_repeat_boundary_v5(cut, out, None, 2000, 30000, str(tmp_path), 4, 0, ref, 0)
    names, contigs = fasta.read_fasta(out)
    assert len(names) > 10
    _, refc = fasta.read_fasta(ref)
    for n in names:
        c, rng = n.split(":")
        a, b = (int(x) for x in rng.split("-"))
        assert contigs[n] == refc[c][a:b] and 80 <= b - a < 30000
    G = oracle.Genome(wl.contig_seqs)
    _, _, exp_iv = G.coarse_boundary([len(c) for c in wl.contig_seqs], seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64)
    assert sorted(names) == sorted({"%s:%d-%d" % (wl.contig_names[c], a, b) for c, a, b in exp_iv})
    flanked = str(tmp_path / "longest_repeats_0.flanked.fa")
    coarse.flanking_seq(out, flanked, ref, 50)
    fn, fc = fasta.read_fasta(flanked)
    assert len(fn) == len(set(fn)) and len(fn) <= len(names)
    for n in fn:
        c, rng = n.split(":")
        a, b = (int(x) for x in rng.split("-"))
        assert fc[n] == refc[c][a - 1:b]
