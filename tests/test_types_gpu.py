"""GPU: hb_adjust end to end for Helitron and non-LTR candidates (judge_boundary_v6 / v9 after the same copy
search + star MSA + truncated-member gate) vs the C port, bit-exact, plus planted-truth recovery."""
import numpy as np
import pytest

from workloads import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,te_type,judge", [("helitron", "helitron", oracle.judge_v6),
                                                ("non_ltr", "non_ltr", oracle.judge_v9)])
def test_adjust_other_te_types_bit_exact(ctx, kind, te_type, judge):
    wl = synth.make_workload(genome_bp=2_000_000, n_contigs=2, n_families=12, cands_per_family=3, copies=(6, 24),
                             length=(250, 1600), seed=17, kind=kind, inexact_frac=0.3)
    G = oracle.Genome(wl.contig_seqs)
    ctx.load_genome(wl.contig_seqs)
    ctx.build_index(15, 10)
    res = ctx.adjust(wl.cand_seqs, te_type, 1)
    n_te = n_exact = 0
    for ci, (seq, tr) in enumerate(zip(wl.cand_seqs, wl.truth)):
        cps, rows = G.candidate_msa(seq)
        assert ctx.last_msa(ci) == rows, ci
        if not rows:
            assert res[ci]["no_copies"]
            continue
        exp = oracle.judge_with_trunc_gate(seq.decode(), cps, rows, judge=judge)
        r = res[ci]
        assert (r["is_te"], r["info"], r["copy_num"]) == (exp["is_te"], exp["info"], exp["copy_num"]), (ci, kind)
        if exp["is_te"]:
            assert r["cons"] == exp["cons"], (ci, kind)
        n_te += r["is_te"]
        n_exact += r["is_te"] and r["cons"] == tr["consensus"].decode()
    assert n_te >= 0.7 * len(wl.cand_seqs)
    assert n_exact >= 0.5 * len(wl.cand_seqs)
