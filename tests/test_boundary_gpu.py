"""GPU: the reference-facing Python entry points (same signatures as module/Util.py:8032, :10407, :10439)
end to end, FASTA in -> FASTA out."""
import os

import numpy as np
import pytest

from hite_b200 import boundary, fasta
from workloads import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


class _Log:
    class logger:
        @staticmethod
        def info(msg):
            pass


def test_flank_region_align_v5_fasta_contract(tmp_path):
    wl = synth.make_workload(genome_bp=1_000_000, n_contigs=2, n_families=10, cands_per_family=3, seed=11)
    ref = str(tmp_path / "genome.fa")
    cand = str(tmp_path / "cand.fa")
    wl.write_fasta(ref, cand)
    out = str(tmp_path / "confident_tir_0.r0.fa")
    low = str(tmp_path / "low.fa")
    res = boundary.flank_region_align_v5(cand, out, 50, ref, None, "tir", str(tmp_path), 4, 0, _Log, None, 1, 0, 0, low, "cons")
    names, contigs = fasta.read_fasta(out)
    assert len(names) >= 0.7 * len(wl.cand_names)
    assert set(names) <= set(wl.cand_names)                      # names unchanged (Util.py:8161, 8282)
    raw = open(out).read().splitlines()
    assert all(l.startswith(">") for l in raw[0::2]) and all(set(l) <= set("ACGTN") for l in raw[1::2])
    # against the oracle on the same inputs (post filter included)
    G = oracle.Genome(wl.contig_seqs)
    for name, seq, r in zip(wl.cand_names, wl.cand_seqs, res):
        cps, rows = G.candidate_msa(seq)
        exp = oracle.judge_with_trunc_gate(seq.decode(), cps, rows) if rows else dict(is_te=False, cons="", copy_num=0)
        assert (r["is_te"], r["cons"]) == (exp["is_te"], exp["cons"])
        keep = exp["is_te"] and not (exp["cons"].startswith("TG") and exp["cons"].endswith("CA")) and exp["copy_num"] > 5
        assert (name in contigs) == keep
        if keep:
            assert contigs[name] == exp["cons"]
    # planted truth
    ok = 0
    for name, tr in zip(wl.cand_names, wl.truth):
        c = tr["consensus"].decode()
        rc = synth.revcomp_arr(np.frombuffer(tr["consensus"], dtype=np.uint8)).tobytes().decode()
        ok += contigs.get(name) in (c, rc)
    assert ok >= 0.7 * len(wl.cand_names)


def test_run_find_members_v8_on_member_files(tmp_path):
    """Member files as flank_region_align_v5 writes them (Util.py:8125-8137): copies +-50 bp, unaligned."""
    wl = synth.make_workload(genome_bp=600_000, n_contigs=1, n_families=3, cands_per_family=1, seed=5,
                             length=(300, 600))
    G = oracle.Genome(wl.contig_seqs)
    seq = wl.cand_seqs[0]
    cps, rows = G.candidate_msa(seq)
    members = {}
    for c in cps:
        if c["keep"]:
            s, e = c["start1"], c["end1"]
            ext = wl.contig_seqs[c["contig"]][s - 1 - 50:e + 50]
            if c["rel"]:
                ext = synth.revcomp_arr(ext)
            members["%s:%d-%d(%s)" % (wl.contig_names[c["contig"]], s, e, "-" if c["rel"] else "+")] = ext.tobytes().decode()
    mf = boundary.member_file_name(str(tmp_path), wl.cand_names[0])
    fasta.store_fasta(members, mf)
    name, cons, info, copy_num, f = boundary.run_find_members_v8((wl.cand_names[0], seq.decode(), None, mf), str(tmp_path),
                                                                 None, 1, "tir", 0, "cons")
    exp = oracle.judge_v5(seq.decode(), oracle.remove_sparse_cols(rows), 1)      # no trunc file -> full pass only
    assert name == wl.cand_names[0] and cons == exp["cons"] and copy_num == exp["copy_num"]
    # an already aligned file is judged directly
    af = str(tmp_path / "aligned.fa")
    fasta.store_fasta({"r%d" % i: r for i, r in enumerate(rows)}, af)
    name2, cons2, _, copy2, _ = boundary.is_TE_from_align_file(af, "q", seq.decode(), str(tmp_path), None, 1, "tir", 0, "cons")
    assert (name2, cons2, copy2) == ("q", exp["cons"], exp["copy_num"])


def test_stage_process_uses_every_listed_gpu(tmp_path, monkeypatch):
    """HITE_GPU_DEVICES=0,1: flank_region_align_v5 spreads its candidates over both GPUs (one context + host thread each)
    and returns exactly what one GPU returns.  Needs a box with >= 2 GPUs (gpurun --gpus 2); skipped on one."""
    from hite_b200 import _lib
    if _lib.load().hb_device_count() < 2:
        pytest.skip("one GPU visible")
    wl = synth.make_workload(genome_bp=2_000_000, n_contigs=2, n_families=30, cands_per_family=4, seed=17)
    ref, cand = str(tmp_path / "genome.fa"), str(tmp_path / "cand.fa")
    wl.write_fasta(ref, cand)
    outs = []
    for devs in ("0", "0,1"):
        monkeypatch.setenv("HITE_GPU_DEVICES", devs)
        monkeypatch.setattr(boundary, "_CTX", {})
        monkeypatch.setattr(boundary, "_GENOME", {})
        out, low = str(tmp_path / ("real_%s.fa" % devs)), str(tmp_path / ("low_%s.fa" % devs))
        res = boundary.flank_region_align_v5(cand, out, 50, ref, None, "tir", str(tmp_path), 4, 0, _Log, None, 1, 0, 0, low, "cons")
        outs.append((res, open(out).read(), open(low).read()))
        assert sorted(boundary._CTX) == [int(x) for x in devs.split(",")]
    assert outs[0] == outs[1]
    assert sum(1 for r in outs[0][0] if r["is_te"]) > 60
