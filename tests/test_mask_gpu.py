"""GPU: known-TE genome masking (reference mask_genome_intactTE, module/Util.py:6389-6431) -- the N-fill / export kernels
through the C-ABI against numpy, the whole mirror against the oracle (HSPs -> full-length chaining -> N-fill) and planted
truth, and the masking step inside determine_repeat_boundary_v5 (:4654)."""
import os
import zlib

import numpy as np
import pytest

import mask_synth
import other_synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _to_n(s):
    return bytes(c if c in b"ACGT" else ord("N") for c in s)


def test_genome_mask_and_export_kernels():
    import hite_b200
    ctx = hite_b200.Context(0)
    rng = np.random.default_rng(4)
    contigs = [np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes() for n in (70_001, 33, 150_000)]
    contigs[1] = contigs[1][:10] + b"RYN" + contigs[1][13:]
    ctx.load_genome(contigs)
    for c, s in enumerate(contigs):
        assert ctx.export_contig(c, len(s)) == _to_n(s)
    iv = [(0, 0, 1), (0, 31, 33), (0, 63, 64), (0, 64, 129), (0, 1000, 5000), (0, 4000, 4100), (0, 70_000, 70_001), (0, 69_990, 80_000),
          (1, 5, 20), (2, -5, 3), (2, 149_999, 150_000), (2, 500, 500), (2, 700, 600), (2, 160_000, 170_000)]
    iv += [(2, int(a), int(a) + int(l)) for a, l in zip(rng.integers(1000, 140_000, 200), rng.integers(1, 3000, 200))]
    exp = [np.frombuffer(_to_n(s), dtype=np.uint8).copy() for s in contigs]
    before = sum(int((e == ord("N")).sum()) for e in exp)
    for c, a, b in iv:
        a, b = max(0, a), min(len(exp[c]), b)
        if b > a:
            exp[c][a:b] = ord("N")
    fresh = ctx.mask_genome([i[0] for i in iv], [i[1] for i in iv], [i[2] for i in iv])
    assert fresh == sum(int((e == ord("N")).sum()) for e in exp) - before
    for c, e in enumerate(exp):
        assert ctx.export_contig(c, len(e)) == e.tobytes(), c
    assert ctx.mask_genome([], [], []) == 0
    with pytest.raises(Exception):
        ctx.mask_genome([7], [0], [10])                         # unknown contig
    with pytest.raises(Exception):
        ctx.query_hsps([b"ACGTACGTACGT"])                       # the index is stale after masking: must be rebuilt
    ctx.build_index(15, 10)
    clean = [i for i in range(0, 60_000, 997) if (exp[0][i:i + 600] != ord("N")).all()][0]
    hs = ctx.query_hsps([exp[0][clean:clean + 600].tobytes()])
    assert len(hs) >= 1 and (int(hs[0]["ss"]), int(hs[0]["se"])) == (clean + 1, clean + 600)
    ctx.close()


def test_mask_genome_intactTE_mirror(tmp_path, monkeypatch):
    """File contract of the reference (genome_path + '.masked', one line per record) and content == oracle chain on
    oracle HSPs; every planted full-length copy is N afterwards, 5'-truncated copies (< 95 % of the query) stay."""
    from hite_b200 import boundary, fasta, masking
    names, contigs, lib_names, lib, truth = other_synth.make(seed=21, n_lib=30, n_contig=3, contig_bp=300_000)
    gpath, lpath = str(tmp_path / "genome.fa"), str(tmp_path / "prev_TE.fa")
    contigs = [bytearray(c) for c in contigs]
    contigs[0][1234] = ord("R")                                  # outside every planted copy? keep it wherever it is
    contigs = [bytes(c) for c in contigs]
    with open(gpath, "w") as f:
        for n, s in zip(names, contigs):
            f.write(">%s\n%s\n" % (n, s.decode()))
    with open(lpath, "w") as f:
        for n, s in zip(lib_names, lib):
            f.write(">%s\n%s\n" % (n, s.decode()))
    monkeypatch.setattr(boundary, "_CTX", {})
    monkeypatch.setattr(boundary, "_GENOME", {})
    out = masking.mask_genome_with_TE(lpath, gpath, str(tmp_path), 4, 0, debug=0)
    assert out == gpath + ".masked" and os.path.exists(out)
    mnames, masked = fasta.read_fasta(out)
    assert mnames == names
    # oracle: same HSP procedure on the CPU, strand dropped, chained by the pinned restatement
    G = O.Genome(contigs)
    exp = [bytearray(c) for c in contigs]
    for qi, q in enumerate(lib):
        hs = [(h["contig"], h["qs"], h["qe"], min(h["ss"], h["se"]), max(h["ss"], h["se"])) for h in G.query_hsps(q)]
        for subj, a, b, _ in O.full_length_copies_v1(hs, len(q), 0.95):
            exp[subj][a:b] = b"N" * (b - a)
    G.close()
    for n, e in zip(names, exp):
        assert masked[n] == bytes(e).decode(), n
    assert masked[names[0]][1234] in "RN"
    n_full = 0
    for t in truth:
        seg = masked[names[t["contig"]]][t["start1"] - 1:t["end1"]]
        if t["kind"] == "full":
            assert seg.count("N") >= len(seg) - 24, t            # boundaries may be off by a few bases of chance matches
            n_full += 1
        elif t["kind"] == "trunc5" and (t["end1"] - t["start1"] + 1) < 0.9 * t["qlen"]:
            assert seg.count("N") == 0, t
    assert n_full > 30
    # the device copy was masked too (the next stage reuses it): same N pattern as the file, modulo IUPAC -> N
    ctx = boundary.get_context()
    for c, n in enumerate(names):
        assert ctx.export_contig(c, len(contigs[c])) == _to_n(masked[n].encode())
    # no library: plain copy (:6430-6431)
    masking.mask_genome_intactTE(None, gpath, str(tmp_path), 1, 0)
    assert open(gpath + ".masked").read() == open(gpath).read()


def test_coarse_stage_masks_known_TEs(tmp_path, monkeypatch):
    """determine_repeat_boundary_v5 with prev_TE (module/Util.py:4654): repeats that are full-length copies of known TEs
    drop out of longest_repeats, the others stay."""
    from hite_b200 import boundary, coarse, fasta
    names, contigs, lib_names, lib, truth = other_synth.make(seed=33, n_lib=12, n_contig=2, contig_bp=400_000, copies=(4, 8),
                                                             length=(400, 2500), sub=(0.0, 0.03))
    gpath = str(tmp_path / "genome.fa")
    with open(gpath, "w") as f:
        for n, s in zip(names, contigs):
            f.write(">%s\n%s\n" % (n, s.decode()))
    known = str(tmp_path / "prev_TE.fa")
    with open(known, "w") as f:                                  # half of the families are already known
        for n, s in list(zip(lib_names, lib))[:6]:
            f.write(">%s\n%s\n" % (n, s.decode()))
    monkeypatch.setattr(boundary, "_CTX", {})
    monkeypatch.setattr(boundary, "_GENOME", {})
    out0, out1 = str(tmp_path / "lr_plain.fa"), str(tmp_path / "lr_masked.fa")
    coarse.determine_repeat_boundary_v5(gpath, out0, None, 2000, 30000, str(tmp_path), 4, 0, gpath, 0)
    coarse.determine_repeat_boundary_v5(gpath, out1, known, 2000, 30000, str(tmp_path), 4, 0, gpath, 0)
    n0, _ = fasta.read_fasta(out0)
    n1, _ = fasta.read_fasta(out1)

    def covered(recs, t):
        for r in recs:
            c, span = r.split(":")
            a, b = (int(x) for x in span.split("-"))
            if c == names[t["contig"]] and min(b, t["end1"]) - max(a, t["start1"] - 1) > 0.5 * (t["end1"] - t["start1"]):
                return True
        return False
    # FMEA keeps a repeated segment once per result file (module/Util.py:4328-4390), not once per copy: count families
    full = [t for t in truth if t["kind"] == "full"]

    def families(recs, lo, hi):
        return {t["query"] for t in full if lo <= t["query"] < hi and covered(recs, t)}
    known0, known1 = families(n0, 0, 6), families(n1, 0, 6)
    novel0, novel1 = families(n0, 6, 12), families(n1, 6, 12)
    assert len(known0) >= 5 and len(known1) == 0, (known0, known1)
    assert len(novel1) >= len(novel0) - 1 and len(novel1) >= 3, (novel0, novel1)
