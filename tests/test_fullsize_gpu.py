"""GPU, BASELINE.json full size (configs[1]: 100 Mbp genome, 10 000 TIR candidates): size-independent
properties of the path + bit-exact spot checks against the C port."""
import numpy as np
import pytest

from workloads import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_c2_full_size_properties(ctx):
    wl = synth.make_workload(genome_bp=100_000_000, n_contigs=10, n_families=1000, cands_per_family=10, copies=(10, 40),
                             length=(200, 2000), seed=1)
    assert len(wl.cand_seqs) == 10_000
    ctx.load_genome(wl.contig_seqs)
    ctx.build_index(15, 10)
    calls, cons = ctx.adjust_raw(wl.cand_seqs, "tir", 1, overlap=0)
    copies = ctx.last_copies()
    # 0. stream-overlapped sub-batches (hb_params.overlap) return the very same bytes as the serial batch
    for ov in (2, 5):
        c_ov, k_ov = ctx.adjust_raw(wl.cand_seqs, "tir", 1, overlap=ov)
        assert c_ov.tobytes() == calls.tobytes() and np.array_equal(k_ov, cons), ov
    with pytest.raises(Exception):
        ctx.last_copies()                     # the per-pair getters need a serial run
    calls, cons = ctx.adjust_raw(wl.cand_seqs, "tir", 1, overlap=0)
    text = cons.tobytes().decode("latin-1")
    seqs = [text[int(o):int(o) + int(l)] for o, l in zip(calls["cons_off"], calls["cons_len"])]
    te = calls["is_te"] != 0
    # 1. planted truth: most accepted consensi equal the planted family consensus (either strand)
    n_exact = 0
    for i in np.nonzero(te)[0]:
        c = wl.truth[i]["consensus"]
        n_exact += seqs[i].encode() in (c, synth.revcomp_arr(np.frombuffer(c, dtype=np.uint8)).tobytes())
    assert te.sum() >= 0.9 * len(te) and n_exact >= 0.8 * te.sum(), (int(te.sum()), n_exact)
    # 2. copy records obey the reference's rules (module/Util.py:8003-8025, 8103-8108)
    kept = copies[copies["keep"] != 0]
    lens = np.array([len(s) for s in wl.cand_seqs])
    L = lens[kept["cand"]]
    aligned = kept["i_e"] - kept["i_s"] + 1
    assert np.all(aligned / L >= 0.95) and np.all(aligned / (kept["j_e"] - kept["j_s"] + 1) >= 0.95)
    clen = np.array([len(c) for c in wl.contig_seqs])
    assert np.all(kept["start1"] - 1 - 50 >= 0) and np.all(kept["end1"] + 50 <= clen[kept["contig"]])
    assert np.all(kept["ext_len"] >= 100)
    rows_per = np.bincount(kept["cand"], minlength=len(lens))
    assert rows_per.max() <= 100
    # 3. batch-split invariance: the verdict of a candidate does not depend on what else is in the call
    sub = list(range(0, 10_000, 7))
    calls2, cons2 = ctx.adjust_raw([wl.cand_seqs[i] for i in sub], "tir", 1)
    t2 = cons2.tobytes().decode("latin-1")
    for k, i in enumerate(sub):
        assert (int(calls2["is_te"][k]), int(calls2["copy_num"][k]), int(calls2["homo_start"][k])) == \
               (int(calls["is_te"][i]), int(calls["copy_num"][i]), int(calls["homo_start"][i]))
        assert t2[int(calls2["cons_off"][k]):int(calls2["cons_off"][k]) + int(calls2["cons_len"][k])] == seqs[i]
    # 4. idempotence on accepted elements: re-adjusting an emitted consensus returns the same consensus
    acc = [i for i in np.nonzero(te)[0][:400]]
    calls3, cons3 = ctx.adjust_raw([seqs[i] for i in acc], "tir", 1)
    t3 = cons3.tobytes().decode("latin-1")
    same = sum(t3[int(o):int(o) + int(l)] == seqs[i] for i, o, l in zip(acc, calls3["cons_off"], calls3["cons_len"]))
    assert same >= 0.9 * len(acc), same
    # 5. bit-exact spot check against the C port at full genome size
    G = oracle.Genome(wl.contig_seqs)
    assert np.array_equal(ctx.debug_index(), G.index)
    ctx.adjust_raw([wl.cand_seqs[i] for i in range(0, 10_000, 500)], "tir", 1)
    for k, i in enumerate(range(0, 10_000, 500)):
        cps, rows = G.candidate_msa(wl.cand_seqs[i])
        assert ctx.last_msa(k) == rows, i
        if rows:
            exp = oracle.judge_with_trunc_gate(wl.cand_seqs[i].decode(), cps, rows)
            assert (bool(calls["is_te"][i]), seqs[i] if exp["is_te"] else "") == (exp["is_te"], exp["cons"] if exp["is_te"] else ""), i


def test_c2_at_spec_verdicts_match_the_port_on_a_sample(ctx):
    """BASELINE configs[1] at SURVEY section 8(d) spec (copies U[6,100], 0-10 % divergence, indels <= 2 %, families of more
    than 100 copies included): every 25th candidate (400) goes through the C port with all host threads and through
    hb_adjust -- is_te, info, copy_num and the consensus length agree on every one (the check bench.py repeats per run)."""
    import os
    from oracle import oracle as O
    wl = synth.config_c2()
    assert len(wl.cand_seqs) == 10_000 and 99_000_000 <= wl.genome_bp <= 101_000_000
    ctx.load_genome(wl.contig_seqs, wl.contig_names)
    ctx.build_index(15, 10)
    sample = list(range(0, 10_000, 25))
    G = O.Genome(wl.contig_seqs)
    G.set_tie_rule(1, wl.contig_names)
    assert np.array_equal(ctx.debug_index(), G.index)
    verd = G.run_batch([wl.cand_seqs[i] for i in sample], os.cpu_count() or 1)
    calls, _ = ctx.adjust_raw([wl.cand_seqs[i] for i in sample], "tir", 1, trunc_gate=False)
    mine = np.stack([calls["is_te"], calls["info"], calls["copy_num"], calls["cons_len"]], axis=1).astype(np.int32)
    bad = np.nonzero(~(mine == verd).all(axis=1))[0]
    assert len(bad) == 0, (len(bad), [(sample[b], mine[b].tolist(), verd[b].tolist()) for b in bad[:5]])
    assert int(calls["is_te"].sum()) >= 0.8 * len(sample)
    # the same candidates inside the full batch: a verdict does not depend on what else is in the call
    full, _ = ctx.adjust_raw(wl.cand_seqs, "tir", 1, trunc_gate=False)
    for k in ("is_te", "info", "copy_num", "cons_len"):
        assert np.array_equal(full[k][sample], calls[k]), k
    G.close()
